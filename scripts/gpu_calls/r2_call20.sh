#!/bin/bash
mkdir -p gpurun_out
for cta in 0 70; do
PMF_TC_TRACE=gpurun_out/r2c20_trace_cta$cta.bin PMF_TC_TRACE_CTA=$cta timeout -s KILL 90 python scripts/tc_time.py > gpurun_out/r2c20_trace_$cta.log 2>&1; echo "rc=$?" >> gpurun_out/r2c20_trace_$cta.log
done
cat gpurun_out/r2c20_trace_0.log
for ab in 16 3 8 27; do PMF_TC_ABLATE=$ab timeout -s KILL 90 python scripts/tc_time.py 2>&1 | tail -1; done > gpurun_out/r2c20_ablate.log
cat gpurun_out/r2c20_ablate.log
