#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 120 scripts/experiments/tma_stream_bench > gpurun_out/r2c34_tma_stream.log 2>&1; cat gpurun_out/r2c34_tma_stream.log
