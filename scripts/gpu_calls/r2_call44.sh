#!/bin/bash
mkdir -p gpurun_out
for b in C2 normal bernoulli poisson; do if [ $b = C2 ]; then unset PMF_BLOCKS; else export PMF_BLOCKS=$b; fi; timeout -s KILL 40 python scripts/tc_time.py 2>&1 | tail -1; done
unset PMF_BLOCKS
timeout -s KILL 200 python -m pytest tests -m gpu -q -x -k "tc or smoke or fit" --timeout 60 2>&1 | tail -4
