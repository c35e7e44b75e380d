#!/bin/bash
# drain warps + two staging buffers + all-BF16 Z contraction: parity of the tcgen05 tests, C2 timing
mkdir -p gpurun_out
timeout -s KILL 90 python scripts/tc_time.py > gpurun_out/r2c19_tc_time.log 2>&1; echo "rc=$?" >> gpurun_out/r2c19_tc_time.log
cat gpurun_out/r2c19_tc_time.log
timeout -s KILL 240 python -m pytest tests -m gpu -q -x -k "tc or smoke or fit" --timeout 60 > gpurun_out/r2c19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c19_pytest.log
tail -15 gpurun_out/r2c19_pytest.log
