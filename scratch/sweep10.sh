#!/bin/bash
run() { python bench.py --workload c5 --steps 12 --warmup 6 --no-cpu-baseline --e2e-steps 0 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print('  value %.3e  ms/step %.3f  push %.3f sort %.3f solve %.3f frac %.3f'%(d['value'],d['ms_per_step'],p['push'],p['sort'],p['solve'],d['roofline']['frac']))"; }
for k in 1 2 3 4; do echo "sort $k"; run --sort-interval $k; done
