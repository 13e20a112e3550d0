#!/bin/bash
# refresh of profiles/ at the end of the round (one B200)
mkdir -p gpurun_out/final
for w in c4 c5 c1 c2 c3; do python bench.py --workload $w > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/final/launches_c4.csv python bench.py --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_c5.csv python bench.py --workload c5 --steps 8 --warmup 3 --no-cpu-baseline --e2e-steps 0 > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_push_boris --launch-skip 8 -c 2 -o gpurun_out/final/push2d -f python bench.py --steps 4 --warmup 3 --e2e-steps 0 --no-cpu-baseline --sort-interval 0 > /dev/null 2>&1
ls -la gpurun_out/final
