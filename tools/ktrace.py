#!/usr/bin/env python3
"""Kernel trace of back-to-back steps (ADEC_KTRACE=1): per-launch duration, effective SM MHz, gap to the previous launch."""
import ctypes, os, sys
os.environ["ADEC_KTRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from audiodec_b200 import _lib
dev = torch.device("cuda:0")
tx, rx, dec = bench.build_codec("symad", dev)
x = [(0.1 * torch.randn(64, 1, 48000)).to(dev) for _ in range(4)]
lib = _lib.load()
buf = (ctypes.c_ulonglong * (3 * 4096))()
def drain():
    out = []
    for g in (tx, dec):
        n = lib.adec_ktrace(g._h, buf, 4096)
        out += [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(max(n, 0))]
    return sorted(out)
for i in range(20): bench.codec_step(tx, rx, dec, x[i % 4])
torch.cuda.synchronize(); drain()
for mode in ("back-to-back", "synced"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5):
        bench.codec_step(tx, rx, dec, x[i % 4])
        if mode == "synced": torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    rec = drain()
    dur = sum(b - a for a, b, c in rec) / 1e6
    span = (rec[-1][1] - rec[0][0]) / 1e6
    gaps = [rec[i + 1][0] - rec[i][1] for i in range(len(rec) - 1)]
    mhz = sorted(c / (b - a) * 1e3 for a, b, c in rec if b > a)
    big = sorted(gaps)[-8:]
    print(f"{mode}: events {e0.elapsed_time(e1):.2f} ms for 5 steps; {len(rec)} conv launches: busy {dur:.2f} ms, span {span:.2f} ms; "
          f"SM MHz min/median/max {mhz[0]:.0f}/{mhz[len(mhz)//2]:.0f}/{mhz[-1]:.0f}; gap median {sorted(gaps)[len(gaps)//2]/1e3:.1f} us, largest {[round(g/1e3,1) for g in big]} us")
