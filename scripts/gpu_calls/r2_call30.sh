#!/bin/bash
mkdir -p gpurun_out
export PMF_BLOCKS=normal
for ab in 0 64 128 192 32 16 208; do PMF_TC_ABLATE=$ab timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1; done > gpurun_out/r2c30_ablate_normal.log
for fl in 512 1024 4096; do PMF_TC_FLAGS=$fl timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1; done >> gpurun_out/r2c30_ablate_normal.log
cat gpurun_out/r2c30_ablate_normal.log
