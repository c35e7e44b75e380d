#!/bin/bash
mkdir -p gpurun_out
python scripts/config_times.py C5 --steps 1 > gpurun_out/r2c4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:zlink -s 3 -c 1 -o gpurun_out/r2c4_zlink python scripts/config_times.py C5 --steps 1 > gpurun_out/r2c4_ncu.log 2>&1
ls -la gpurun_out/r2c4*; tail -3 gpurun_out/r2c4_ncu.log
