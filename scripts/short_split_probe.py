#!/usr/bin/env python3
"""1-d lines of 2048 ... 8192 points: four-step layout (64 x N2, specialised kernels) against the one-pass layout on the generic
kernels (FC_FLAG_NO_SHORT_SPLIT); batch-segment plans whose windows have these lengths included. CUDA-graph replay, L2 flushed."""
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
CASES = [((32, 32, 8192), (64, 32, 129)), ((8, 8, 4096), (8, 8, 65)), ((64, 16, 2048), (16, 16, 33)), ((4, 64, 8192), (64, 64, 257)),
         ((16, 128, 4096), (128, 128, 100)), ((2, 64, 70000), (64, 64, 513)), ((1, 16, 8192), (16, 16, 65)), ((256, 8, 2048), (8, 8, 17)),
         ((32, 8, 8192), (8, 8, 257)), ((4, 128, 40000), (128, 128, 1000)), ((1, 64, 2048), (64, 64, 9)), ((128, 32, 4096), (32, 32, 65))]
for xs, ws in CASES:
    x = torch.randn(*xs, device=dev)
    w = torch.randn(*ws, device=dev)
    b = torch.randn(ws[0], device=dev)
    res = []
    for flags in (L.FC_FLAG_NO_SHORT_SPLIT, 0):
        Fn.set_default_flags(flags)
        Fn.clear_caches()
        with torch.no_grad():
            for _ in range(3):
                fcp.fft_conv(x, w, b)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y = fcp.fft_conv(x, w, b)
        ts = []
        for i in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(e0.elapsed_time(e1) * 1e3)
        d = Fn._plans[next(reversed(Fn._plans))].plan
        ds = d.describe()
        res.append(f"{statistics.median(ts):.1f} us ({'four-step' if 'structure=2' in ds else 'one pass'} fft {d.fft_size[0]} x {int(d.info.segments)})")
        del g, y
    Fn.set_default_flags(0)
    print(xs, ws, " | ".join(res), flush=True)
