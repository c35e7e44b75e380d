#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 scripts/experiments/tma_stream_bench > gpurun_out/r2c33_tma_stream.log 2>&1; cat gpurun_out/r2c33_tma_stream.log
export PMF_BLOCKS=normal
PMF_TC_FLAGS=128 PMF_TC_TRACE=gpurun_out/r2c33_loopA.bin PMF_TC_TRACE_CTA=70 timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
