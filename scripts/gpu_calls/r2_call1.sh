#!/bin/bash
# round 2, GPU call 1: state of HEAD (GPU tests, bench) + the L2-traffic ablations of the tcgen05 data pass
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c1_pytest.log
timeout 600 python scripts/tc_ablate.py 0 256 64 128 192 32 96 > gpurun_out/r2c1_ablate.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err
tail -3 gpurun_out/r2c1_pytest.log; cat gpurun_out/r2c1_ablate.log; cat gpurun_out/r2c1_bench.json
