#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/wide_check.py > gpurun_out/r2c2_wide_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2c2_wide_check.log
cat gpurun_out/r2c2_wide_check.log
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -k "wide or refuses" > gpurun_out/r2c2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c2_pytest.log
tail -30 gpurun_out/r2c2_pytest.log
timeout 900 python scripts/config_times.py C4b C5 --steps 10 > gpurun_out/r2c2_config_times.jsonl 2> gpurun_out/r2c2_config_times.err
cat gpurun_out/r2c2_config_times.jsonl; tail -5 gpurun_out/r2c2_config_times.err
