#!/bin/bash
# baseline of the restored tree: GPU suite, bench, every config, launch list + full ncu capture of the C2 data pass and the K > 64 kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c18_pytest.log
tail -8 gpurun_out/r2c18_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c18_bench.json 2> gpurun_out/r2c18_bench.err
cut -c1-300 gpurun_out/r2c18_bench.json; tail -3 gpurun_out/r2c18_bench.err
timeout 1200 python scripts/config_times.py C2 C3 C4a C4b C5 --steps 10 > gpurun_out/r2c18_config_times.jsonl 2> gpurun_out/r2c18_config_times.err
cut -c1-330 gpurun_out/r2c18_config_times.jsonl; tail -3 gpurun_out/r2c18_config_times.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c18_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2c18_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:data_pass_tc -s 5 -c 1 -o gpurun_out/r2c18_c2_tc python scripts/tc_time.py > gpurun_out/r2c18_ncu_tc.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:zlink|grad_gemm" -s 9 -c 3 -o gpurun_out/r2c18_c5_wide python scripts/wide_time.py 10000x50000x128 > gpurun_out/r2c18_ncu_wide.log 2>&1
tail -3 gpurun_out/r2c18_ncu_tc.log gpurun_out/r2c18_ncu_wide.log
ls -la gpurun_out
