#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 90 python scripts/tc_check.py 2>&1 | grep "prec=0" | tail -3
for b in C2 normal bernoulli; do if [ $b = C2 ]; then unset PMF_BLOCKS; else export PMF_BLOCKS=$b; fi; timeout -s KILL 40 python scripts/tc_time.py 2>&1 | tail -1; done
