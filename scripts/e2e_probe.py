#!/usr/bin/env python3
"""Where the end-to-end (host buffers) time of the c2 call goes: raw PCIe copy times, kernel time, chunk counts."""
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


def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


print("H2D 67 MB: %.3f ms" % t(lambda: xd.copy_(xh, non_blocking=True)))
print("D2H 51 MB: %.3f ms" % t(lambda: yh.copy_(yd, non_blocking=True)))
s2 = torch.cuda.Stream()


def both():
    xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2):
        yh.copy_(yd, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


print("H2D + D2H concurrently: %.3f ms" % t(both))
with torch.no_grad():
    print("kernels only (device tensors): %.3f ms" % t(lambda: m(xd)))
    for ch in (1, 2, 3, 4, 6, 8):
        Fn._HOST_PIPELINE_CHUNKS = ch
        print("e2e module call, %d chunks: %.3f ms" % (ch, t(lambda: m(xh), 15)))
