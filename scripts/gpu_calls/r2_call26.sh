#!/bin/bash
mkdir -p gpurun_out
PMF_BLOCKS=normal PMF_TC_FLAGS=64 PMF_TC_TRACE=gpurun_out/r2c26_epi.bin PMF_TC_TRACE_CTA=70 timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_BLOCKS=normal PMF_TC_FLAGS=32 PMF_TC_TRACE=gpurun_out/r2c26_mma.bin PMF_TC_TRACE_CTA=70 timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_BLOCKS=normal PMF_TC_TRACE=gpurun_out/r2c26_all.bin PMF_TC_TRACE_CTA=70 timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
