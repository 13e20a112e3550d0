#!/bin/bash
run() { python bench.py --workload c5 --steps 12 --warmup 6 --no-cpu-baseline --e2e-steps 0 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print('  value %.3e  ms/step %.3f  push %.3f sort %.3f solve %.3f frac %.3f'%(d['value'],d['ms_per_step'],p['push'],p['sort'],p['solve'],d['roofline']['frac']))"; }
echo "tile on"; timeout 300 python -m pytest tests/test_gpu_3d.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3; run
echo "tile off"; MAG3D_TILE=0 run
