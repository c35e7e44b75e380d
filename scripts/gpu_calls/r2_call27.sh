#!/bin/bash
mkdir -p gpurun_out
for ab in 0 16 1 2 3 8 4 31; do PMF_BLOCKS=normal PMF_TC_ABLATE=$ab timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1; done > gpurun_out/r2c27_ablate_normal.log
cat gpurun_out/r2c27_ablate_normal.log
