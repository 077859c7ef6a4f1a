#!/usr/bin/env python3
"""Host-side cost of one eager call (no GPU wait): wall time per call over a burst, plus a cProfile of the burst."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fft_conv_pytorch_b200 as fcp

m = fcp.FFTConv2d(8, 8, 65).cuda()
x = torch.randn(8, 8, 512, 512, device="cuda")
with torch.no_grad():
    for _ in range(5):
        m(x)
    torch.cuda.synchronize()
    for n in (50, 200):
        t0 = time.perf_counter()
        for _ in range(n):
            m(x)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"burst {n}: host {1e6 * (t1 - t0) / n:.1f} us/call, incl. drain {1e6 * (t2 - t0) / n:.1f} us/call")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        m(x)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
    g = fcp.graphed(m, x)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        g()
    torch.cuda.synchronize()
    print(f"graph replay: {1e6 * (time.perf_counter() - t0) / 200:.1f} us/call incl. GPU")
