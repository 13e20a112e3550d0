#!/bin/bash
run() { python bench.py --workload $1 --steps 36 --warmup 8 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print('  $1 value %.3e  ms/step %.3f  push %.3f solve %.3f frac %.3f'%(d['value'],d['ms_per_step'],p['push'],p['solve'],d['roofline']['frac']))"; }
echo "member redux"; python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -m gpu -x -q 2>&1 | tail -1; run c4; run c4; run c3
echo "full-mask redux"; export MAG2D_B200_LIB=$PWD/scratch/variants/lib_fullredux.so; run c4; run c4; run c3
