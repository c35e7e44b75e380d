#!/bin/bash
mkdir -p gpurun_out
PMF_TC_FLAGS=64 PMF_TC_TRACE=gpurun_out/r2c22_epi_cta0.bin PMF_TC_TRACE_CTA=0 timeout -s KILL 90 python scripts/tc_time.py 2>&1 | tail -1
PMF_TC_FLAGS=64 PMF_TC_TRACE=gpurun_out/r2c22_epi_cta100.bin PMF_TC_TRACE_CTA=100 timeout -s KILL 90 python scripts/tc_time.py 2>&1 | tail -1
PMF_TC_FLAGS=32 PMF_TC_TRACE=gpurun_out/r2c22_mma_cta0.bin PMF_TC_TRACE_CTA=0 timeout -s KILL 90 python scripts/tc_time.py 2>&1 | tail -1
