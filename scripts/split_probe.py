#!/usr/bin/env python3
"""Does running BASELINE c2 as batch parts (so that the spectra between the three kernels stay in the 126 MB L2) beat one
full-batch call?  Device time of k back-to-back fc_conv calls on B/k items each, L2 flushed before the group, CUDA-graph replay."""
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import fft_conv_pytorch_b200 as fcp

dev = torch.device("cuda", 0)
x = torch.randn(8, 8, 512, 512, device=dev)
w = torch.randn(8, 8, 65, 65, device=dev)
b = torch.randn(8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for parts in (1, 2, 4, 8):
    xs = [t.contiguous() for t in x.chunk(parts)]
    with torch.no_grad():
        for t in xs:
            fcp.fft_conv(t, w, b)  # plans, kernel spectrum
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ys = [fcp.fft_conv(t, w, b) for t in xs]
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            ys = [fcp.fft_conv(t, w, b) for t in xs]
    ts = []
    for i in range(40):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        if i >= 5:
            ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"c2 as {parts} part(s) of {8 // parts}: {statistics.median(ts):.1f} us (min {min(ts):.1f})", flush=True)
