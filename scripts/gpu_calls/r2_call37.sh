#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_TC_CTATIMES=gpurun_out/r2c37_cta.bin timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r2c37_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c37_pytest.log
tail -8 gpurun_out/r2c37_pytest.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c37_bench.json 2> gpurun_out/r2c37_bench.err
cut -c1-400 gpurun_out/r2c37_bench.json; tail -3 gpurun_out/r2c37_bench.err
