#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
PMF_NO_CARVEOUT_HINT=1 timeout -s KILL 300 python bench.py --steps 20 --warmup 5 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no hint: ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
timeout -s KILL 300 python bench.py --steps 20 --warmup 5 2> /dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('hint   : ms_per_step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'])"
done
