#!/bin/bash
run() { python bench.py --steps 24 --warmup 8 --no-cpu-baseline --e2e-steps 0 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print('  value %.3e  ms/step %.3f  push %.3f sort %.3f solve %.3f frac %.3f'%(d['value'],d['ms_per_step'],p['push'],p['sort'],p['solve'],d['roofline']['frac']))"; }
echo "standalone sort 8"; MAG2D_FUSED_SORT=0 run
echo "fused 8"; run
for e in 1 2 3 4 6; do echo "fused ions 64 electrons $e"; run --sort-intervals ARGON_POS=64,ELECTRON=$e; done
echo "fused ions 256 electrons 2"; run --sort-intervals ARGON_POS=256,ELECTRON=2
