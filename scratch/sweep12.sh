#!/bin/bash
# 2-D push deposit variants on C4
run() { python bench.py --workload c4 --steps 36 --warmup 8 --no-cpu-baseline --e2e-steps 0 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print('  value %.3e  ms/step %.3f  push %.3f sort %.3f solve %.3f frac %.3f'%(d['value'],d['ms_per_step'],p['push'],p['sort'],p['solve'],d['roofline']['frac']))"; }
echo default; run; run
for lib in scratch/variants/lib2_*.so; do echo "$lib"; MAG2D_B200_LIB=$PWD/$lib python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deposit or selfconsistent" 2>&1 | tail -1; MAG2D_B200_LIB=$PWD/$lib run; done
