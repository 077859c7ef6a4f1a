#!/usr/bin/env python3
"""Small 1-d problems: one transform against forced batch segments (FC_FLAG_SEGMENT), CUDA-graph replay, L2 flushed."""
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
CASES = [((1, 8, 32768), (8, 8, 1025)), ((2, 4, 32768), (4, 4, 100)), ((1, 16, 20000), (16, 16, 65)), ((1, 8, 65536), (8, 8, 2049)),
         ((4, 8, 32768), (8, 8, 1025)), ((8, 8, 32768), (8, 8, 129)), ((1, 4, 131072), (4, 4, 513)), ((2, 16, 40000), (16, 16, 33))]
for xs, ws in CASES:
    x = torch.randn(*xs, device=dev)
    w = torch.randn(*ws, device=dev)
    b = torch.randn(ws[0], device=dev)
    res = []
    for flags in (L.FC_FLAG_NO_SEGMENT, L.FC_FLAG_SEGMENT):
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
        for i in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            if i >= 5:
                ts.append(e0.elapsed_time(e1) * 1e3)
        d = Fn._plans[next(reversed(Fn._plans))].plan
        res.append(f"{statistics.median(ts):.1f} us (fft {d.fft_size[0]} x {int(d.info.segments)})")
    Fn.set_default_flags(0)
    print(xs, ws, " | ".join(res), flush=True)
