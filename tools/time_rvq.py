#!/usr/bin/env python3
"""RVQ kernel alone (CUDA events): quantize (idx only) and the fused idx + packed + zq launch, 64 x 160 and 256 x 5 frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
tx, rx, dec = bench.build_codec("symad", dev)
for B, F in ((64, 160), (256, 5), (128, 160)):
    z = 0.6 * torch.randn(B, 64, F, device=dev)
    for name, fn in (("quantize", lambda: tx.quantize(z)), ("fused idx+packed+zq", lambda: tx.quantize_fused(z, want_idx=True, want_packed=True, want_zq=True))):
        for _ in range(5): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"rvq {B}x{F} frames  {name:22s} {e0.elapsed_time(e1) / 50:.4f} ms")
