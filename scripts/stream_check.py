#!/usr/bin/env python3
"""Streaming K1 / K4 (csrc/fc_stream.cuh) against the register-path kernels: bit-identical outputs expected (same arithmetic),
and against F.conv2d in fp64. Prints one line per case; exits 1 on a mismatch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import _lib as L
from fft_conv_pytorch_b200 import functional as Fn

dev = torch.device("cuda", 0)
torch.manual_seed(0)
CASES = [
    # (x shape, w shape, kwargs)
    ((8, 8, 512, 512), (8, 8, 65, 65), {}),
    ((2, 8, 448, 448), (8, 8, 65, 65), dict(padding=32)),          # zero padding 32: per-row copies of 448 floats at offset 32
    ((3, 4, 200, 448), (4, 4, 17, 33), dict(padding=(8, 16))),     # rows of 448 + 2*16 = 480 -> 512; R = 200 + 16 rows
    ((2, 8, 250, 500), (8, 8, 9, 13), {}),                         # L = 500 (16-byte multiple), R = 250 -> partial last tile
    ((2, 8, 250, 498), (8, 8, 9, 13), {}),                         # L = 498: not a 16-byte multiple -> K1 falls back, K4 streams
    ((1, 8, 100, 512), (16, 8, 5, 5), dict(stride=1)),
]
bad = 0
for xs, ws, kw in CASES:
    x = torch.randn(*xs, device=dev)
    w = torch.randn(*ws, device=dev)
    b = torch.randn(ws[0], device=dev)
    outs = []
    for flags in (0, L.FC_FLAG_NO_STREAM):
        Fn.set_default_flags(flags)
        Fn.clear_caches()
        outs.append(fcp.fft_conv(x, w, b, **kw).clone())
    ref = F.conv2d(x.double(), w.double(), b.double(), **kw)
    same = torch.equal(outs[0], outs[1])
    err = ((outs[0].double() - ref).norm() / ref.norm()).item()
    maxd = (outs[0] - outs[1]).abs().max().item()
    print(f"{xs} {ws} {kw}: identical={same} maxdiff={maxd:.3e} rel_err_vs_fp64={err:.2e}", flush=True)
    if not same or not err < 1e-5:
        bad += 1
sys.exit(1 if bad else 0)
