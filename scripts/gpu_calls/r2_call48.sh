#!/bin/bash
# last single-GPU call of the round (6 GPU-minutes left): GPU suite, smoke and bench of the final tree, then the numbers
# behind the looser test tolerances.  Most important first; every step under its own timeout.
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests -m gpu -q --timeout 100 -p no:cacheprovider > gpurun_out/r2c48_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c48_pytest.log
tail -6 gpurun_out/r2c48_pytest.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c48_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c48_smoke.log; tail -6 gpurun_out/r2c48_smoke.log
timeout -s KILL 150 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c48_bench.json 2> gpurun_out/r2c48_bench.err
cut -c1-400 gpurun_out/r2c48_bench.json; tail -2 gpurun_out/r2c48_bench.err
timeout -s KILL 90 python scripts/measure_tolerances.py > gpurun_out/r2c48_tolerances.jsonl 2> gpurun_out/r2c48_tolerances.err
cat gpurun_out/r2c48_tolerances.jsonl | cut -c1-600; tail -2 gpurun_out/r2c48_tolerances.err
timeout -s KILL 90 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --kernel-events separate > gpurun_out/r2c48_bench_separate.json 2> gpurun_out/r2c48_bench_separate.err
cut -c1-200 gpurun_out/r2c48_bench_separate.json; tail -2 gpurun_out/r2c48_bench_separate.err
