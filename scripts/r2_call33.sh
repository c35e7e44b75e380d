#!/bin/bash
mkdir -p gpurun_out
export PMF_BLOCKS=normal
PMF_TC_FLAGS=128 PMF_TC_TRACE=gpurun_out/r2c33_loopA.bin PMF_TC_TRACE_CTA=70 timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_TC_ABLATE=256 timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
