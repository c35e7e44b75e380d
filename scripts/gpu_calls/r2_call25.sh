#!/bin/bash
mkdir -p gpurun_out
PMF_TC_CTATIMES=gpurun_out/r2c25_cta.bin timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_BLOCKS=normal PMF_TC_CTATIMES=gpurun_out/r2c25_cta_normal.bin timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
