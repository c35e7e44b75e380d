#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/wide_check.py > gpurun_out/r2c9_wide_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2c9_wide_check.log
cat gpurun_out/r2c9_wide_check.log
timeout 600 python scripts/wide_time.py 10000x50000x128 0 1 2 3 > gpurun_out/r2c9_wide_flags.log 2>&1
timeout 600 python scripts/wide_time.py 10000x30000x256 0 1 2 3 >> gpurun_out/r2c9_wide_flags.log 2>&1
cat gpurun_out/r2c9_wide_flags.log
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "wide" > gpurun_out/r2c9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c9_pytest.log
tail -15 gpurun_out/r2c9_pytest.log
timeout 600 python scripts/config_times.py C4b C5 --steps 10 > gpurun_out/r2c9_config_times.jsonl 2>gpurun_out/r2c9_config_times.err; cat gpurun_out/r2c9_config_times.jsonl
