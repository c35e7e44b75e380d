#!/bin/bash
# final call of the round: the tree with the allocator cache (guard.cu) and the tightened tolerances -- GPU suite, smoke,
# bench (e2e = median of five), end-to-end breakdown
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests -m gpu -q --timeout 100 -p no:cacheprovider > gpurun_out/r2c50_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c50_pytest.log
tail -8 gpurun_out/r2c50_pytest.log
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c50_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2c50_smoke.log; tail -5 gpurun_out/r2c50_smoke.log
timeout -s KILL 120 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c50_bench.json 2> gpurun_out/r2c50_bench.err
cat gpurun_out/r2c50_bench.json; tail -2 gpurun_out/r2c50_bench.err
timeout -s KILL 60 python scripts/e2e_breakdown.py > gpurun_out/r2c50_e2e_breakdown.log 2>&1; tail -9 gpurun_out/r2c50_e2e_breakdown.log
