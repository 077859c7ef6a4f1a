#!/usr/bin/env python3
"""Per-kernel device time of 1-d problems with 32 ... 128 channels per group (where does the contraction run, what does it cost)."""
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fft_conv_pytorch_b200 import _lib as L
from fft_conv_pytorch_b200 import functional as Fn

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CASES = [((64, 16, 1024), (16, 16, 33)), ((256, 64, 512), (64, 64, 9)), ((32, 3, 224, 224), (64, 3, 7, 7)), ((32, 64, 56, 56), (64, 64, 3, 3)),
         ((16, 64, 128, 128), (64, 64, 5, 5)), ((8, 32, 256, 256), (32, 32, 9, 9)), ((4, 16, 512, 512), (16, 16, 31, 31)), ((16, 8, 100, 100), (8, 8, 11, 11)),
         ((8, 4, 32, 32, 32), (8, 4, 5, 5, 5)), ((2, 16, 64, 64, 64), (16, 16, 7, 7, 7)), ((4, 1, 128, 128, 128), (4, 1, 9, 9, 9)), ((8, 128, 64, 64), (128, 128, 7, 7)),
         ((2, 8, 1024, 1024), (8, 8, 33, 33)), ((1, 3, 2048, 2048), (3, 3, 65, 65))]
if len(sys.argv) > 1:
    CASES = eval(sys.argv[1])
FLAGS = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # FC_FLAG_* for every plan (32: no tensor cores)
P = lambda t_: ctypes.c_void_p(t_.data_ptr())
for xs, ws in CASES:
    x = torch.randn(*xs, device=dev)
    w = torch.randn(*ws, device=dev)
    b = torch.randn(ws[0], device=dev)
    nd = len(xs) - 2
    entry = Fn.get_plan(False, xs[0], xs[1], ws[0], 1, tuple(xs[2:]), tuple(ws[2:]), (1,) * nd, (0,) * nd, (1,) * nd, (0,) * nd, "constant", FLAGS)
    plan = entry.plan
    lib = plan.lib
    kspec = Fn.kernel_spectrum(entry, w, dev)
    const = entry.const_for(dev)
    wsb = torch.empty(int(plan.info.workspace_bytes), dtype=torch.uint8, device=dev)
    y = torch.empty((xs[0], ws[0]) + plan.out_size, device=dev)
    nl = int(plan.info.n_launches)
    ms = (ctypes.c_float * nl)()
    n_out = ctypes.c_int(0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    acc = [[] for _ in range(nl)]
    for i in range(8):
        flush.zero_()
        L.check(lib, lib.fc_conv_profiled(plan.handle, P(const), P(x), P(kspec), P(b), P(y), P(wsb), st, ms, nl, ctypes.byref(n_out)), "prof")
        if i >= 3:
            for j in range(nl):
                acc[j].append(ms[j] * 1e3)
    names = []
    for j in range(nl):
        nm = ctypes.create_string_buffer(64)
        ab = ctypes.c_int64(0)
        lib.fc_plan_launch_info(plan.handle, j, nm, 64, ctypes.byref(ab))
        names.append((nm.value.decode(), ab.value))
    tot = sum(statistics.median(a) for a in acc)
    abytes = sum(ab for _, ab in names)
    print(xs, ws, f"fft {plan.fft_size} x {int(plan.info.segments)} total {tot:.0f} us = {abytes / tot / 1e3:.0f} GB/s:",
          " ".join(f"{n}={statistics.median(a):.0f}us({ab / statistics.median(a) / 1e3:.0f}GB/s)" for (n, ab), a in zip(names, acc)), flush=True)
    del x, w, y, wsb, kspec
    Fn.clear_caches()
    torch.cuda.empty_cache()
