#!/usr/bin/env python
"""per-source-line summary of an ncu report: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv; python ncu_lines.py f.csv [top]`
prints, for the CUDA lines with the most executed warp instructions: instructions, stall samples, shared wavefronts, L1 tag requests"""
import csv
import sys

rows, hdr, cur = [], None, None
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None or len(r) != len(hdr):
        continue
    if r[2] != "-":          # SASS rows repeat the line's totals
        continue
    rows.append((cur, r))
col = {k: hdr.index(k) for k in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared", "L1 Tag Requests Global", "Thread Instructions Executed")}
tot = {k: sum(int(r[c]) for _, r in rows) for k, c in col.items()}
print("totals:", tot)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
for f, r in sorted(rows, key=lambda fr: -int(fr[1][col["# Samples"]]))[:top]:
    print("%-16s %4s %-90s inst %5.1f%% samp %5.1f%% smemwf %5.1f%% l1tag %5.1f%% thr/inst %4.1f" % (
        f, r[0], r[1].strip()[:90], 100.0 * int(r[col["Instructions Executed"]]) / max(tot["Instructions Executed"], 1),
        100.0 * int(r[col["# Samples"]]) / max(tot["# Samples"], 1), 100.0 * int(r[col["L1 Wavefronts Shared"]]) / max(tot["L1 Wavefronts Shared"], 1),
        100.0 * int(r[col["L1 Tag Requests Global"]]) / max(tot["L1 Tag Requests Global"], 1),
        int(r[col["Thread Instructions Executed"]]) / max(int(r[col["Instructions Executed"]]), 1)))
