#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 200 > gpurun_out/r2c39_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/r2c39_pytest_multi.log; tail -4 gpurun_out/r2c39_pytest_multi.log
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c39_bench_n2.json 2> gpurun_out/r2c39_bench_n2.err; echo "bench rc=$?"
grep -v "^NCCL\|^$" gpurun_out/r2c39_bench_n2.json | cut -c1-330; tail -2 gpurun_out/r2c39_bench_n2.err
