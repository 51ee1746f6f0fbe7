#!/usr/bin/env python3
"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): last N launches of our kernels."""
import csv, re, sys
path = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 38
lines = open(path).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = [r for r in csv.DictReader(lines[start:]) if 'adec::' in r['Kernel Name'] and 'replicate' not in r['Kernel Name'] and 'mma_probe' not in r['Kernel Name']]
last = rows[-n:]
tot = sum(float(r['Metric Value']) for r in last)
print(f"{len(rows)} adec launches in file; last {n}: total {tot/1e6:.3f} ms")
agg = {}
for r in last:
    name = re.sub(r'void adec::|\(adec::\w+\)|\(int\)|\(bool\)', '', r['Kernel Name'])
    us = float(r['Metric Value']) / 1e3
    print(f"{name:42s} grid {r['Grid Size']:18s} blk {r['Block Size']:12s} {us:9.1f} us {100*us*1e3/tot:5.1f}%")
    agg[name] = agg.get(name, 0) + us
print("-- by kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k:42s} {v:9.1f} us {100*v*1e3/tot:5.1f}%")
