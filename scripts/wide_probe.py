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
CASES = [((16, 64, 65536), (64, 64, 4097)), ((16, 64, 65536), (128, 64, 4097)), ((16, 128, 65536), (64, 128, 4097)), ((16, 96, 65536), (96, 96, 4097)),
         ((8, 64, 16384), (64, 64, 257)), ((32, 32, 8192), (64, 32, 129))]
P = lambda t_: ctypes.c_void_p(t_.data_ptr())
for xs, ws in CASES:
    x = torch.randn(*xs, device=dev)
    w = torch.randn(*ws, device=dev)
    b = torch.randn(ws[0], device=dev)
    entry = Fn.get_plan(False, xs[0], xs[1], ws[0], 1, tuple(xs[2:]), tuple(ws[2:]), (1,), (0,), (1,), (0,), "constant")
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
    print(xs, ws, f"fft {plan.fft_size[0]} x {int(plan.info.segments)} total {tot:.0f} us:",
          " ".join(f"{n}={statistics.median(a):.0f}us({ab / statistics.median(a) / 1e3:.0f}GB/s)" for (n, ab), a in zip(names, acc)), flush=True)
    del x, w, y, wsb, kspec
    Fn.clear_caches()
    torch.cuda.empty_cache()
