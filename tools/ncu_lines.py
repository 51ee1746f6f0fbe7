#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump per CUDA source line:
   usage: ncu -i X.ncu-rep --page source --csv --print-source sass,cuda | python tools/ncu_lines.py [top_n]"""
import csv, sys
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))
cur_file = None
hdr = None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed"); i_samp = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:110])
    a = agg.setdefault(key, [0, 0])
    num = lambda v: int(v) if v.lstrip("-").isdigit() else 0
    a[0] += num(r[i_inst]); a[1] += num(r[i_samp])
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot_i}, stall samples {tot_s}")
for (f, ln, src), (ins, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*ins/tot_i:5.1f}% inst {100*smp/max(1,tot_s):5.1f}% samp  {f}:{ln}  {src}")
