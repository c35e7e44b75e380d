#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_BLOCKS=normal timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
PMF_BLOCKS=bernoulli timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1
timeout -s KILL 200 python -m pytest tests -m gpu -q -x -k "tc" --timeout 60 2>&1 | tail -3
