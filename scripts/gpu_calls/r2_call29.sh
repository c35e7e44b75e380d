#!/bin/bash
mkdir -p gpurun_out
for b in C2 normal bernoulli; do if [ $b = C2 ]; then unset PMF_BLOCKS; else export PMF_BLOCKS=$b; fi; timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1; done > gpurun_out/r2c29_time.log
unset PMF_BLOCKS
cat gpurun_out/r2c29_time.log
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "tc or smoke or fit" --timeout 60 > gpurun_out/r2c29_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c29_pytest.log
tail -5 gpurun_out/r2c29_pytest.log
