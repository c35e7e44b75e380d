#!/bin/bash
timeout -s KILL 200 python scripts/config_times.py C2 C3 --steps 10 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config'], 'ms_per_step %.4f data_pass_ms %.4f' % (d['ms_per_step'], d['data_pass_ms']))"
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "tc or batch or c3" --timeout 100 2>&1 | tail -3
