#!/bin/bash
for lib in libpmf_2cd11dd libpmf_8cef424 libpmf_HEAD_1 current; do
  if [ $lib = current ]; then unset PMF_LIB; else export PMF_LIB=$PWD/scripts/experiments/$lib.so; fi
  echo -n "$lib: "; timeout -s KILL 200 python scripts/config_times.py C3 --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step %.4f data_pass_ms %.4f launches %.1f' % (d['ms_per_step'], d['data_pass_ms'], d['kernel_launches_per_step']))"
done
