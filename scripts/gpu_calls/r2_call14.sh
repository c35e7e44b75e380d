#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sanitize_case.py > gpurun_out/r2c14_plain.log 2>&1 || { cat gpurun_out/r2c14_plain.log; exit 1; }
cat gpurun_out/r2c14_plain.log
timeout 2400 compute-sanitizer --tool memcheck --log-file gpurun_out/r2c14_memcheck.log python scripts/sanitize_case.py > gpurun_out/r2c14_memcheck_stdout.log 2>&1
echo "memcheck rc=$?"; tail -15 gpurun_out/r2c14_memcheck.log; cat gpurun_out/r2c14_memcheck_stdout.log | tail -8
