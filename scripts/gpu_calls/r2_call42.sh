#!/bin/bash
mkdir -p gpurun_out
PMF_BLOCKS=normal timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:data_pass_tc -s 5 -c 1 -o gpurun_out/r2c42_normal_ts python scripts/tc_time.py > gpurun_out/r2c42_ncu.log 2>&1
tail -1 gpurun_out/r2c42_ncu.log
PMF_BLOCKS=normal PMF_TC_FLAGS=128 PMF_TC_TRACE=gpurun_out/r2c42_loopA.bin PMF_TC_TRACE_CTA=70 timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
