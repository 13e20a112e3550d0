#!/bin/bash
run() { python bench.py --workload $1 --steps 36 --warmup 8 --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print('  $1 value %.3e  ms/step %.3f  push %.3f solve %.3f frac %.3f'%(d['value'],d['ms_per_step'],p['push'],p['solve'],d['roofline']['frac']))"; }
echo "ticketless"; python -m pytest tests -m gpu -x -q 2>&1 | tail -2; run c4; run c5; run c3
echo "tickets"; export MAG2D_B200_LIB=$PWD/scratch/variants/lib_tickets.so; run c4; run c5; run c3
