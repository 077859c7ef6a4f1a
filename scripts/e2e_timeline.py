#!/usr/bin/env python3
"""Event timeline of the chunked host-buffer pipeline at BASELINE c2 (where the end-to-end time goes)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import functional as Fn

dev = torch.device("cuda", 0)
m = fcp.FFTConv2d(8, 8, 65).to(dev)
xh = torch.randn(8, 8, 512, 512).pin_memory()
xd = torch.empty_like(xh, device=dev)
yd = torch.empty(8, 8, 448, 448, device=dev)
yh = torch.empty(8, 8, 448, 448).pin_memory()
cur = torch.cuda.current_stream()
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def run(bounds, record=False):
    ev = []
    E = lambda: torch.cuda.Event(enable_timing=True)
    t0 = E(); t0.record(cur)
    s_in.wait_stream(cur); s_out.wait_stream(cur)
    ins = []
    with torch.cuda.stream(s_in):
        for a0, a1 in bounds:
            xd[a0:a1].copy_(xh[a0:a1], non_blocking=True)
            e = E(); e.record(s_in); ins.append(e)
    outs, comps = [], []
    with torch.no_grad():
        for c, (a0, a1) in enumerate(bounds):
            cur.wait_event(ins[c])
            yd[a0:a1] = m(xd[a0:a1])
            e = E(); e.record(cur); comps.append(e)
            with torch.cuda.stream(s_out):
                s_out.wait_event(e)
                yh[a0:a1].copy_(yd[a0:a1], non_blocking=True)
                e2 = E(); e2.record(s_out); outs.append(e2)
    cur.wait_stream(s_out)
    s_out.synchronize()
    if record:
        return [t0.elapsed_time(e) for e in ins], [t0.elapsed_time(e) for e in comps], [t0.elapsed_time(e) for e in outs]


for name, bounds in [("4 even", [(0, 2), (2, 4), (4, 6), (6, 8)]), ("8 even", [(i, i + 1) for i in range(8)]),
                     ("3,2,2,1", [(0, 3), (3, 5), (5, 7), (7, 8)]), ("2,2,2,1,1", [(0, 2), (2, 4), (4, 6), (6, 7), (7, 8)]),
                     ("1,2,2,2,1", [(0, 1), (1, 3), (3, 5), (5, 7), (7, 8)])]:
    for _ in range(3):
        run(bounds)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        run(bounds)
    torch.cuda.synchronize()
    wall = 1e3 * (time.perf_counter() - t) / 10
    i, c, o = run(bounds, True)
    print(f"{name}: wall {wall:.3f} ms/iter; H2D done {['%.2f' % v for v in i]} compute done {['%.2f' % v for v in c]} D2H done {['%.2f' % v for v in o]}")
with torch.no_grad():
    for ch in (4, 8):
        Fn._HOST_PIPELINE_CHUNKS = ch
        for _ in range(3):
            m(xh)
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(15):
            m(xh)
        torch.cuda.synchronize()
        print("module call, %d chunks: %.3f ms" % (ch, 1e3 * (time.perf_counter() - t) / 15))
