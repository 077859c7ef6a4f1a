#!/usr/bin/env python3
"""Device time of 1-d problems with and without batch segments (FC_FLAG_NO_SEGMENT), L2 flushed, cached kernel spectrum."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import _lib as L
from fft_conv_pytorch_b200 import functional as Fn

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CASES = [((4, 8, 33000), (8, 8, 64), {}), ((8, 8, 40000), (8, 8, 129), dict(padding=64)), ((1, 8, 70000), (8, 8, 1025), {}),
         ((16, 16, 100000), (16, 16, 257), {}), ((2, 64, 70000), (64, 64, 513), {}), ((4, 128, 40000), (128, 128, 1000), {}),
         ((16, 256, 65536), (256, 256, 4097), {})]
for xs, ws, kw in CASES:
    x = torch.randn(*xs, device=dev)
    w = torch.randn(*ws, device=dev)
    b = torch.randn(ws[0], device=dev)
    res = []
    for flags in (0, L.FC_FLAG_NO_SEGMENT):
        Fn.set_default_flags(flags)
        Fn.clear_caches()
        with torch.no_grad():
            for _ in range(3):
                fcp.fft_conv(x, w, b, **kw)
            ts = []
            for _ in range(9):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fcp.fft_conv(x, w, b, **kw)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
        d = Fn._plans[next(reversed(Fn._plans))].plan.describe().splitlines()
        res.append((statistics.median(ts), d[0][:90] if "batch segments" in d[0] else "one transform"))
        del ts
    Fn.set_default_flags(0)
    print(xs, ws, kw, " | ".join(f"{t:.3f} ms ({n})" for t, n in res), flush=True)
