#!/bin/bash
mkdir -p gpurun_out
PMF_BLOCKS=normal timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:data_pass_tc -s 5 -c 1 -o gpurun_out/r2c28_normal python scripts/tc_time.py > gpurun_out/r2c28_ncu_normal.log 2>&1
tail -2 gpurun_out/r2c28_ncu_normal.log
