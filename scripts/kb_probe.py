#!/usr/bin/env python3
"""Per-kernel device time of one config (CUDA events around every launch, L2 flushed, median of N calls): one line.
    python scripts/kb_probe.py [c2|c5|img256] [reps]      (tuning builds: set FFTCONV_B200_* knobs in the environment)"""
import ctypes
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fft_conv_pytorch_b200 import _lib as L
from fft_conv_pytorch_b200 import functional as Fn

CFG = {
    "c2": ((8, 8, 512, 512), (8, 8, 65, 65), False, {}),
    "img256": ((16, 8, 256, 256), (8, 8, 31, 31), False, {}),
    "c5": ((4, 64, 1024, 1024), (64, 16, 31, 31), True, dict(stride=2, dilation=2, groups=4)),
}
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 15
xs, ws, tr, kw = CFG[name]
Fn.set_default_flags(int(os.environ.get("FFTCONV_B200_PROBE_FLAGS", "0")))
dev = torch.device("cuda", 0)
x = torch.randn(*xs, device=dev)
w = torch.randn(*ws, device=dev)
cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
b = torch.randn(cout, device=dev)
nd = len(xs) - 2
t = lambda v: (v,) * nd
entry = Fn.get_plan(tr, xs[0], xs[1], cout, kw.get("groups", 1), tuple(xs[2:]), tuple(ws[2:]), t(kw.get("stride", 1)), t(0), t(kw.get("dilation", 1)),
                    t(0), "constant")
plan = entry.plan
lib = plan.lib
kspec = Fn.kernel_spectrum(entry, w, dev)
const = entry.const_for(dev)
wsb = torch.empty(int(plan.info.workspace_bytes), dtype=torch.uint8, device=dev)
y = torch.empty((xs[0], cout) + plan.out_size, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nl = int(plan.info.n_launches)
ms = (ctypes.c_float * nl)()
n_out = ctypes.c_int(0)
P = lambda t_: ctypes.c_void_p(t_.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
acc = [[] for _ in range(nl)]
for i in range(reps + 3):
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
    names.append(nm.value.decode())
env = " ".join(f"{k[13:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("FFTCONV_B200_") and k != "FFTCONV_B200_TUNING")
print(name, env or "-", " ".join(f"{n}={statistics.median(a):.1f}us" for n, a in zip(names, acc)), "total=%.1fus" % sum(statistics.median(a) for a in acc), flush=True)
