#!/bin/bash
# final single-GPU evidence of the round: GPU suite, smoke, bench, every config, launch list, full ncu capture of the C2 data pass
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/r2c45_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c45_pytest.log
tail -4 gpurun_out/r2c45_pytest.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c45_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c45_smoke.log; tail -4 gpurun_out/r2c45_smoke.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c45_bench.json 2> gpurun_out/r2c45_bench.err
cut -c1-300 gpurun_out/r2c45_bench.json; tail -2 gpurun_out/r2c45_bench.err
timeout -s KILL 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c45_bench_reference.json 2> gpurun_out/r2c45_bench_reference.err
cut -c1-300 gpurun_out/r2c45_bench_reference.json
timeout -s KILL 900 python scripts/config_times.py C2 P25 C3 C4a C4b C5 --steps 10 > gpurun_out/r2c45_config_times.jsonl 2> gpurun_out/r2c45_config_times.err
cut -c1-260 gpurun_out/r2c45_config_times.jsonl; tail -2 gpurun_out/r2c45_config_times.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c45_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/r2c45_ncu_bench.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:data_pass_tc -s 5 -c 1 -o gpurun_out/r2c45_c2_tc python scripts/tc_time.py > gpurun_out/r2c45_ncu_tc.log 2>&1
tail -1 gpurun_out/r2c45_ncu_tc.log
