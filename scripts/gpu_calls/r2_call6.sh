#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/wide_time.py 10000x50000x128 0 4 1 2 3 7 > gpurun_out/r2c6_wide_flags.log 2>&1
timeout 900 python scripts/wide_time.py 10000x30000x256 0 4 1 2 3 7 >> gpurun_out/r2c6_wide_flags.log 2>&1
cat gpurun_out/r2c6_wide_flags.log
