#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "network or fit_loss_curve or c1_runtests" > gpurun_out/r2c13_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c13_pytest.log
tail -15 gpurun_out/r2c13_pytest.log
timeout 900 python scripts/config_times.py C4a C4b --steps 10 > gpurun_out/r2c13_config_times.jsonl 2> gpurun_out/r2c13_config_times.err
cat gpurun_out/r2c13_config_times.jsonl; tail -3 gpurun_out/r2c13_config_times.err
python scripts/config_times.py C4a --steps 2 > gpurun_out/r2c13_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c13_launches_c4a.csv python scripts/config_times.py C4a --steps 2 > gpurun_out/r2c13_ncu.log 2>&1
awk -F'","' 'NR>2{n=split($5,a,"("); print a[1], $NF}' gpurun_out/r2c13_launches_c4a.csv | tail -24
