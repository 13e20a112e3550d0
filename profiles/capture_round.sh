#!/bin/bash
# profiles/capture_round.sh — every ncu capture the round's records come from, in one go (run on a B200 box, from the repo root):
#
#     /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash profiles/capture_round.sh'
#
# Writes text only into gpurun_out/ (the .ncu-rep files stay in /tmp: together they exceed what gpurun copies back):
#   r2_<workload>_ncu.txt     one-screen summary per captured push launch (profiles/ncu_summary.py)
#   r2_launches_<w>.csv       launch list of one short bench run (gpu__time_duration per launch: shares, not absolutes)
#   r2_push_dram.json         DRAM bytes per particle-step per workload, keyed by the sha1 of the kernel sources (bench.py reads it)
#   r2_solve3d_ncu.txt        the 3-D field solve kernel by kernel
# Copy what is to be kept into profiles/ afterwards.
set -u
O=/tmp/reps
mkdir -p $O gpurun_out
B="--e2e-steps 0 --no-cpu-baseline --no-secondary"
NCU="ncu --set full --clock-control none"
cap() {   # name, kernel regex, launches to skip, launches to capture, bench arguments...
    local name=$1 kern=$2 skip=$3 cnt=$4
    shift 4
    timeout 600 $NCU -k regex:$kern -s $skip -c $cnt -o $O/$name -f python bench.py "$@" --steps 8 --warmup 3 $B > /dev/null 2>&1
    python profiles/ncu_summary.py $O/$name.ncu-rep > gpurun_out/r2_${name}_ncu.txt 2>&1
}
# one sort cycle of both species of C4 (electrons are re-sorted every 5 pushes): 10 launches
cap c4 k_push_boris 10 10 --workload c4
cap c4_f32 k_push_boris 10 10 --workload c4 --storage f32
# one compaction cycle of the brick store: 4 launches (+ the placement kernels in between)
cap c5 "k_push3d|k_place3d" 8 8 --workload c5
cap c3 k_push_boris 6 4 --workload c3
cap c2 k_push_boris 6 4 --workload c2
timeout 600 $NCU --metrics smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum \
    -k regex:k_push_multicoll -s 4 -c 2 -o $O/c1 -f python bench.py --workload c1 --steps 4 --warmup 3 $B > /dev/null 2>&1
python profiles/ncu_summary.py $O/c1.ncu-rep > gpurun_out/r2_c1_ncu.txt 2>&1
# the 3-D solve, kernel by kernel (one step)
timeout 600 $NCU -k regex:"gemm3|thomas_solve|fold3d|rhs3d|edge_fields3d|capacitance|add_green" -s 36 -c 12 -o $O/solve3d -f python bench.py --workload c5 --steps 3 --warmup 3 $B > /dev/null 2>&1
python profiles/ncu_summary.py $O/solve3d.ncu-rep > gpurun_out/r2_solve3d_ncu.txt 2>&1
for w in c4 c5; do
    timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_$w.csv python bench.py --workload $w --steps 2 --warmup 3 $B > /dev/null 2>&1
done
python profiles/make_traffic.py c4=$O/c4.ncu-rep c4_f32=$O/c4_f32.ncu-rep c5=$O/c5.ncu-rep c3=$O/c3.ncu-rep c2=$O/c2.ncu-rep c1=$O/c1.ncu-rep flops_c1=${FLOPS_C1:-8915} > gpurun_out/r2_traffic_stdout.txt 2>&1
cp profiles/r2_push_dram.json gpurun_out/
ls -la gpurun_out
cat gpurun_out/r2_traffic_stdout.txt
