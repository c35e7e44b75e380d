#!/bin/bash
mkdir -p gpurun_out
PMF_GUARD=1 timeout 600 python scripts/sanitize_case.py > gpurun_out/r2c15_guards.log 2>&1; echo "rc=$?" >> gpurun_out/r2c15_guards.log
cat gpurun_out/r2c15_guards.log
timeout 900 python -m pytest tests -m gpu -q -k "guard or network" > gpurun_out/r2c15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c15_pytest.log
tail -5 gpurun_out/r2c15_pytest.log
timeout 900 python scripts/config_times.py C4a --steps 10 > gpurun_out/r2c15_config_times.jsonl 2> gpurun_out/r2c15_config_times.err
cut -c1-300 gpurun_out/r2c15_config_times.jsonl
