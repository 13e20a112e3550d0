#!/bin/bash
# profiles/final_records.sh — the one-GPU records of a round: GPU test suite, smoke(), the default bench line (C4 + secondary C5 + e2e +
# cpu_baseline), the reference arm, and the side workloads.  Run on a B200 box from the repo root; everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r2_tests_final.log 2>&1
tail -4 gpurun_out/r2_tests_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null
for w in c1 c2 c3; do timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --e2e-steps 0 > gpurun_out/r2_bench_$w.json 2>/dev/null; done
python bench.py --workload c4 --storage f32 --steps 30 --warmup 10 --e2e-steps 0 --no-cpu-baseline --no-secondary > gpurun_out/r2_bench_c4_f32.json 2>/dev/null
python bench.py --workload c5 --steps 20 --warmup 6 > gpurun_out/r2_bench_c5.json 2>/dev/null
python - <<EOF
import json
def L(f): return json.loads(open("gpurun_out/"+f).read().strip().splitlines()[-1])
o=L("r2_bench_c4.json"); print("c4",o["value"],o["ms_per_step"],o["roofline"]["frac"],o["roofline"]["frac_moved"],o["e2e"]["value"],o["cpu_baseline"]["value"],o["phases_ms_per_step"]); s=o["secondary"]["c5"]; print("sec c5",s["value"],s["ms_per_step"],s["roofline"]["frac"])
print("ref",L("r2_bench_reference_arm.json")["value"])
for w in ("c1","c2","c3","c4_f32","c5"):
    o=L("r2_bench_%s.json"%w); print(w,o["value"],o["ms_per_step"],o["roofline"]["frac"],o["roofline"].get("frac_moved"),(o.get("e2e") or {}).get("value"))
EOF
