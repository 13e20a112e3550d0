#!/usr/bin/env python
"""one-screen summary of an ncu report (`ncu --set full`): python profiles/ncu_summary.py X.ncu-rep > profiles/rN_name_ncu.txt
prints per kernel: time, DRAM bytes, pipe utilisations, occupancy, instruction counts, FP64 flops, top stall reasons"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__inst_executed_op_global_red.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("---", d.get("Kernel Name", "?"))
    for k in KEYS:
        if k in d and d[k] != "":
            print("   %-78s %s %s" % (k, d[k], rows[1][hdr.index(k)]))
    fl = [d.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % o, "") for o in ("dfma", "dmul", "dadd")]
    if all(fl):
        f = 2 * float(fl[0]) + float(fl[1]) + float(fl[2])
        t = float(d["gpu__time_duration.sum"]) * (1e-3 if rows[1][hdr.index("gpu__time_duration.sum")] == "ms" else 1e-6 if rows[1][hdr.index("gpu__time_duration.sum")] in ("us", "usecond") else 1e-9)
        print("   FP64 FLOP (2 DFMA + DMUL + DADD, predicated-on threads)                         %.4g  -> %.2f TFLOP/s under ncu" % (f, f / t / 1e12))
    st = sorted(((float(v), k.split("issue_stalled_")[1].split("_per")[0]) for k, v in d.items()
                 if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v), reverse=True)[:6]
    print("   stalls per issue: " + ", ".join("%s %.2f" % (k, v) for v, k in st))
