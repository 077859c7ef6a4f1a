#!/usr/bin/env python3
"""One small call of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_families.py

Families: generic passes + SIMT contraction, K1 / K4 warp and group engines, fused axis kernel (plain, general map,
segmented), the packed batch-pair kernels K1p / KBp / K4p (plain, general map, segmented, odd batch), contiguous C2C
passes, plane kernels (3-d), column kernels (four-step 1-d), tensor-core GEMM. Each result is compared with F.conv*.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import _lib as L
from fft_conv_pytorch_b200 import functional as Fn

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

CASES = [
    # name, x, w, transposed, kwargs, plan flags
    ("generic_2d", (2, 3, 40, 36), (4, 3, 5, 3), False, dict(padding=1), L.FC_FLAG_NO_FUSED),
    ("k1k4_group_c2c_contract", (4, 8, 100, 100), (8, 8, 7, 7), False, {}, L.FC_FLAG_NO_PAIR),
    ("fused_plain", (2, 8, 256, 256), (8, 8, 9, 9), False, {}, L.FC_FLAG_NO_PAIR),
    ("fused_general_seg", (1, 16, 560, 560), (16, 8, 9, 9), True, dict(stride=2, dilation=2, groups=2, padding=2), L.FC_FLAG_NO_PAIR),
    ("pair_plain", (4, 8, 256, 256), (8, 8, 9, 9), False, {}, 0),
    ("pair_general_odd_batch", (3, 8, 200, 180), (8, 8, 5, 7), False, dict(padding=(1, 2)), 0),
    ("pair_seg_lattice", (2, 32, 560, 560), (32, 16, 9, 9), True, dict(stride=2, dilation=2, groups=2, padding=2), 0),
    ("pair_long_rows", (2, 8, 200, 1700), (8, 8, 5, 301), False, {}, 0),
    ("plane_3d", (2, 8, 40, 40, 40), (8, 8, 5, 5, 5), False, {}, 0),
    ("column_1d", (2, 4, 40000), (4, 4, 129), False, {}, 0),
    ("c2c_1d_segments", (1, 8, 32768), (8, 8, 1025), False, {}, 0),
    ("tc_gemm", (4, 64, 600), (128, 64, 9), False, {}, 0),
]


def main():
    only = sys.argv[1:]
    dev = torch.device("cuda", 0)
    worst = 0.0
    for name, xs, ws, tr, kw, flags in CASES:
        if only and name not in only:
            continue
        g = torch.Generator().manual_seed(3)
        x = torch.randn(*xs, generator=g).to(dev)
        w = torch.randn(*ws, generator=g).to(dev)
        cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
        b = torch.randn(cout, generator=g).to(dev)
        Fn.set_default_flags(flags)
        Fn.clear_caches()
        with torch.no_grad():
            y = (fcp.fft_conv_transpose if tr else fcp.fft_conv)(x, w, b, **kw)
            nd = len(xs) - 2
            ref = getattr(F, ("conv_transpose%dd" if tr else "conv%dd") % nd)(x.double(), w.double(), b.double(), **kw)
        torch.cuda.synchronize()
        err = ((y.double() - ref).abs().max() / ref.abs().max()).item()
        worst = max(worst, err)
        print(f"{name}: rel err {err:.2e}", flush=True)
        assert err < 1e-4, name
    Fn.set_default_flags(0)
    print("all families ok, worst rel err %.2e" % worst)


if __name__ == "__main__":
    main()
