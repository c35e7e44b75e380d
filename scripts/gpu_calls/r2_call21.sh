#!/bin/bash
mkdir -p gpurun_out
for fl in 0 512 1024 2048 4096 5120; do PMF_TC_FLAGS=$fl timeout -s KILL 90 python scripts/tc_time.py 2>&1 | tail -1; done > gpurun_out/r2c21_flags.log
cat gpurun_out/r2c21_flags.log
