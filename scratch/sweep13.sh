#!/bin/bash
run() { python bench.py --workload c4 --steps 36 --warmup 8 --no-cpu-baseline --e2e-steps 0 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print('  value %.3e  ms/step %.3f  push %.3f sort %.3f solve %.3f frac %.3f'%(d['value'],d['ms_per_step'],p['push'],p['sort'],p['solve'],d['roofline']['frac']))"; }
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run; run
for lib in scratch/variants/lib2_*.so; do [ -f "$lib" ] || continue; echo "$lib"; MAG2D_B200_LIB=$PWD/$lib run; done
