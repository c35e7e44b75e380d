#!/bin/bash
mkdir -p gpurun_out
for b in normal bernoulli poisson; do PMF_BLOCKS=$b timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1; done > gpurun_out/r2c24_blocks.log
cat gpurun_out/r2c24_blocks.log
PMF_TC_FLAGS=16 PMF_TC_TRACE=gpurun_out/r2c24_cta.bin timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
