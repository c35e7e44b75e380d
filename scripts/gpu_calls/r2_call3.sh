#!/bin/bash
mkdir -p gpurun_out
python scripts/config_times.py C4b C5 --steps 3 > gpurun_out/r2c3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2c3_launches.csv python scripts/config_times.py C4b C5 --steps 3 > gpurun_out/r2c3_ncu.log 2>&1
grep -E "zlink|grad_gemm|prep_wide|multi_pass|control" gpurun_out/r2c3_launches.csv | awk -F'","' '{print $5, $NF}' | tail -40
