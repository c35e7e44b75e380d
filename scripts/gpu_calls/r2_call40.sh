#!/bin/bash
mkdir -p gpurun_out
echo "== layout L1 (lane = (k%16)+32(k/16))"; timeout -s KILL 90 python scripts/tc_check.py 2>&1 | grep "prec=0" | tail -3
echo "== layout L2 (lane = k)"; PMF_TC_FLAGS=65536 timeout -s KILL 90 python scripts/tc_check.py 2>&1 | grep "prec=0" | tail -3
for b in C2 normal bernoulli poisson; do if [ $b = C2 ]; then unset PMF_BLOCKS; else export PMF_BLOCKS=$b; fi; timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1; done
