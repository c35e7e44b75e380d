#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/wide_check.py > gpurun_out/r2c7_wide_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2c7_wide_check.log
cat gpurun_out/r2c7_wide_check.log
timeout 600 python scripts/wide_time.py 10000x50000x128 0 1 2 3 > gpurun_out/r2c7_wide_flags.log 2>&1
timeout 600 python scripts/wide_time.py 10000x30000x256 0 1 2 3 >> gpurun_out/r2c7_wide_flags.log 2>&1
cat gpurun_out/r2c7_wide_flags.log
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "wide" > gpurun_out/r2c7_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c7_pytest.log
tail -5 gpurun_out/r2c7_pytest.log
