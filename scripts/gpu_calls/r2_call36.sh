#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for sa in 3 4; do for b in C2 normal bernoulli poisson; do
  if [ $b = C2 ]; then unset PMF_BLOCKS; else export PMF_BLOCKS=$b; fi
  echo -n "sa=$sa "; PMF_LIB=$PWD/scripts/experiments/libpmf_sa$sa.so timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
done; done; done | tee gpurun_out/r2c36_ab.log
