#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/wide_check.py > gpurun_out/r2c5_wide_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2c5_wide_check.log
cat gpurun_out/r2c5_wide_check.log
timeout 900 python scripts/config_times.py C4b C5 --steps 10 > gpurun_out/r2c5_config_times.jsonl 2> gpurun_out/r2c5_config_times.err
cat gpurun_out/r2c5_config_times.jsonl; tail -5 gpurun_out/r2c5_config_times.err
python scripts/config_times.py C4b C5 --steps 2 > gpurun_out/r2c5_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2c5_launches.csv python scripts/config_times.py C4b C5 --steps 2 > gpurun_out/r2c5_ncu.log 2>&1
grep -E "zlink|grad_gemm" gpurun_out/r2c5_launches.csv | awk -F'","' '{n=split($5,a,"("); print a[1], $NF}'
