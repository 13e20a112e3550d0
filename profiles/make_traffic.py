#!/usr/bin/env python
"""profiles/make_traffic.py — turn this round's ncu captures into profiles/r2_push_dram.json, the file bench.py reads for
`roofline.traffic` (and the FP64 work per particle-step of the multi-collision mover).

    python profiles/make_traffic.py c4=gpurun_out/r2_c4_push.ncu-rep c5=gpurun_out/r2_c5_push.ncu-rep ... [flops_c1=<number>]

For every workload: sum of dram__bytes_read.sum + dram__bytes_write.sum over the captured push launches divided by the live
particles those launches moved (argument `n_<workload>=<particles per launch>`, default the bench's named size).  The file is keyed
by the sha1 of the kernel sources (bench.kernel_source_hash): bench.py refuses it when the sources have changed since."""
import csv
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(rep, kernel_filter):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if kernel_filter not in d.get("Kernel Name", ""):
            continue
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(d[k]) * UNIT[units[hdr.index(k)]]
        out.append((d["Kernel Name"], tot, d.get("gpu__time_duration.sum"), units[hdr.index("gpu__time_duration.sum")]))
    return out


def main():
    table, notes, flops = {}, {}, {}
    per_launch = {"c4": 50_000_000, "c4_f32": 50_000_000, "c5": 125_000_000, "c3": 10_000_000, "c2": 1_000_000, "c1": 1_000_000}
    args = dict(a.split("=", 1) for a in sys.argv[1:])
    for k, v in list(args.items()):
        if k.startswith("n_"):
            per_launch[k[2:]] = float(v)
        elif k.startswith("flops_"):
            flops[k[6:]] = float(v)
    for wl, rep in args.items():
        if wl.startswith(("n_", "flops_")):
            continue
        filt = "k_push3d" if wl == "c5" else "k_push_multicoll" if wl == "c1" else "k_push_boris"
        launches = dram_bytes(rep, filt)
        if not launches:
            continue
        per = sum(b for _, b, _, _ in launches) / (len(launches) * per_launch[wl])
        table[wl] = per
        notes[wl] = {"report": os.path.basename(rep), "launches": [{"kernel": k[:80], "dram_bytes": b, "time": t, "time_unit": u} for k, b, t, u in launches],
                     "particles_per_launch": per_launch[wl]}
    out = {"kernel_source_sha1": bench.kernel_source_hash(), "sources": list(bench.KERNEL_SOURCES), "captured": time.strftime("%Y-%m-%d"),
           "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum per captured push launch / live particles per launch",
           "dram_bytes_per_particle_step": table, "fp64_flop_per_particle_step": flops, "detail": notes}
    with open(bench.TRAFFIC_FILE, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({"dram_bytes_per_particle_step": table, "fp64_flop_per_particle_step": flops}))


if __name__ == "__main__":
    main()
