#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c17_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c17_pytest.log
tail -30 gpurun_out/r2c17_pytest.log
timeout 1200 python scripts/tc_precision_vs_size.py > gpurun_out/r2c17_precision.jsonl 2> gpurun_out/r2c17_precision.err
cat gpurun_out/r2c17_precision.jsonl; tail -3 gpurun_out/r2c17_precision.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err
cut -c1-200 gpurun_out/r2c17_bench.json; tail -3 gpurun_out/r2c17_bench.err
