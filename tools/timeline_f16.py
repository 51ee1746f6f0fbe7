#!/usr/bin/env python3
"""Summarise a -DADEC_TIMELINE event dump of tc_conv_f16_kernel (one CTA).  usage: timeline_f16.py file [first_group n_groups]
events: 1 TMA issue(c) 2 issuer enters group(c) 3 weights landed 4 partial free/issue start 5 issued+committed 6 issuer waits window(p)
        7 window ready 8 producer starts piece(p) 9 producer arrived 10 drain waits(c) 11 drain sees partial 12 drain released 13 epilogue start(tile)"""
import sys, collections
ev = collections.defaultdict(list)
rows = [tuple(map(int, l.split())) for l in open(sys.argv[1]) if not l.startswith("#")]
t0 = min(r[2] for r in rows)
for code, idx, t in rows: ev[code].append((idx, (t - t0) & 0xffffffff))
names = {1: "tma", 2: "enter", 3: "wland", 4: "start", 5: "done", 10: "dwait", 11: "dseen", 12: "drel"}
per = collections.defaultdict(dict)
for code in names:
    for idx, t in ev[code]: per[idx][code] = t
g0 = int(sys.argv[2]) if len(sys.argv) > 2 else 40
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print("group   tma  enter wland start  done | dwait dseen  drel | issue-len start-gap dseen-done")
prev = None
for c in range(g0, g0 + n):
    d = per.get(c, {})
    if 4 not in d: continue
    base = per[g0][4]
    f = lambda k: f"{d[k]-base:6d}" if k in d else "     -"
    gap = d[4] - prev if prev is not None else 0
    prev = d[4]
    print(f"{c:5d} {f(1)} {f(2)} {f(3)} {f(4)} {f(5)} | {f(10)} {f(11)} {f(12)} | {d.get(5,0)-d[4]:8d} {gap:9d} {d.get(11,0)-d.get(5,0):9d}")
starts = sorted(t for _, t in ev[4])
if len(starts) > 20:
    gaps = [b - a for a, b in zip(starts, starts[1:])]
    gaps_s = sorted(gaps)
    print(f"groups {len(starts)}: start-to-start median {gaps_s[len(gaps)//2]} mean {sum(gaps)/len(gaps):.0f} p90 {gaps_s[int(len(gaps)*0.9)]} max {gaps_s[-1]}; span {starts[-1]-starts[0]}")
ep = sorted(t for _, t in ev[13])
if len(ep) > 2: print("epilogue starts every", [b - a for a, b in zip(ep, ep[1:])][:12])
for code, nm in ((8, "producer piece start"), (9, "producer piece arrive"), (7, "issuer window ready")):
    ts = sorted(t for _, t in ev[code])
    if len(ts) > 4:
        g = sorted(b - a for a, b in zip(ts, ts[1:]))
        print(f"{nm}: n {len(ts)} interval median {g[len(g)//2]} max {g[-1]}")
# waits: issuer blocked on weights (3-2), on partial (4-3), on window (7-6)
def waits(a, b):
    A = dict(ev[a]); B = dict(ev[b])
    w = [B[k] - A[k] for k in A if k in B]
    return (sum(w), len(w), max(w) if w else 0)
print("issuer waits: weights", waits(2, 3), " partial", waits(3, 4), " issue", waits(4, 5))
print("drain waits for partial", waits(10, 11), " drain work", waits(11, 12))
