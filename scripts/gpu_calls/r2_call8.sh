#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c8_gpus.txt
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2c8_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c8_pytest_multi.log
tail -30 gpurun_out/r2c8_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c8_bench_n2.json 2> gpurun_out/r2c8_bench_n2.err
cat gpurun_out/r2c8_bench_n2.json; tail -5 gpurun_out/r2c8_bench_n2.err
