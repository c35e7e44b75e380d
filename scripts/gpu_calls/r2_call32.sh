#!/bin/bash
mkdir -p gpurun_out
export PMF_BLOCKS=normal
for fl in 0 768 1536 2560 0; do PMF_TC_FLAGS=$fl timeout -s KILL 60 python scripts/tc_time.py 2>&1 | tail -1; done
