#!/bin/bash
run() { python bench.py --workload $1 --steps 48 --warmup 8 --no-cpu-baseline --e2e-steps 0 "${@:2}" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print('  $* value %.3e  ms/step %.3f  push %.3f'%(d['value'],d['ms_per_step'],p['push']))"; }
for k in 3 4 5; do run c5 --sort-interval $k; done
for k in 4 5 6 8; do run c4 --sort-intervals ELECTRON=$k; done
