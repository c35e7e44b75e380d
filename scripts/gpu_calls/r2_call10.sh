#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2c10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c10_pytest.log
tail -40 gpurun_out/r2c10_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err
cut -c1-900 gpurun_out/r2c10_bench.json; tail -3 gpurun_out/r2c10_bench.err
