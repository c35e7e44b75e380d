#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_BLOCKS=normal timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_TC_CTATIMES=gpurun_out/r2c31_cta.bin timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
