#!/bin/bash
mkdir -p gpurun_out
for fl in 0 512 1024; do PMF_TC_FLAGS=$fl timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1; done > gpurun_out/r2c23_flags.log
cat gpurun_out/r2c23_flags.log
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "tc or smoke or fit" --timeout 60 > gpurun_out/r2c23_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c23_pytest.log
tail -5 gpurun_out/r2c23_pytest.log
