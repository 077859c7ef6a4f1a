#!/usr/bin/env python3
"""One small call of every kernel family on the host-thread emulation (tests/cpu_emul), for the sanitizer builds of
scripts/emul_sanitize.sh: every CUDA thread is a host thread, __syncthreads / __syncwarp / named barriers are real
barriers and shared memory is one heap block, so AddressSanitizer sees every out-of-bounds access of the kernel source
and ThreadSanitizer every pair of conflicting shared- or global-memory accesses that no barrier orders — what
compute-sanitizer memcheck / racecheck would report on the device (it is closed on this GPU pool)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fft_conv_pytorch_b200 import _lib as L  # noqa: E402
from oracle import fftconv_oracle as O  # noqa: E402
from tests.cpu_emul import emul  # noqa: E402

CASES = [
    ("generic_2d", (2, 3, 40, 36), (4, 3, 5, 3), False, dict(padding=1), L.FC_FLAG_NO_FUSED),
    ("k1k4_group_c2c_contract", (2, 8, 70, 70), (8, 8, 7, 7), False, {}, L.FC_FLAG_NO_PAIR),
    ("fused_plain", (2, 8, 256, 256), (8, 8, 9, 9), False, {}, L.FC_FLAG_NO_PAIR),
    ("fused_general_seg", (1, 16, 560, 300), (16, 8, 9, 9), True, dict(stride=2, dilation=2, groups=2, padding=2), L.FC_FLAG_NO_PAIR),
    ("pair_plain", (4, 8, 256, 256), (8, 8, 9, 9), False, {}, L.FC_FLAG_PAIR | L.FC_FLAG_NO_YSTAGE),
    ("pair_general_odd_batch", (3, 8, 200, 180), (8, 8, 5, 7), False, dict(padding=(1, 2)), L.FC_FLAG_PAIR),
    ("pair_seg_lattice_16", (2, 32, 560, 300), (32, 16, 9, 9), True, dict(stride=2, dilation=2, groups=2, padding=2), L.FC_FLAG_PAIR),
    ("pair_long_rows", (2, 8, 140, 1700), (8, 8, 5, 301), False, {}, L.FC_FLAG_PAIR),
    ("pair_ystage4_s64", (3, 8, 130, 250), (8, 8, 7, 3), False, {}, L.FC_FLAG_PAIR),
    ("pair_ystage4_s128_rowseg", (2, 8, 300, 600), (8, 8, 9, 5), False, {}, L.FC_FLAG_PAIR),
    ("pair_ystage8_s128", (2, 8, 600, 140), (8, 8, 9, 5), False, {}, L.FC_FLAG_PAIR | L.FC_FLAG_NO_SEGMENT),
    ("fused_bias_rows", (2, 16, 130, 136), (16, 4, 5, 5), True, dict(groups=2, stride=2, dilation=2, padding=1), 0),
    ("plane_3d", (1, 8, 40, 40, 40), (8, 8, 5, 5, 5), False, {}, 0),
    ("column_1d", (1, 4, 40000), (4, 4, 129), False, {}, 0),
    ("fused_1d_split", (1, 8, 32768), (8, 8, 1025), False, {}, 0),
]


def main():
    only = sys.argv[1:]
    for name, xs, ws, tr, kw, flags in CASES:
        if only and name not in only:
            continue
        rng = np.random.RandomState(3)
        x = rng.standard_normal(xs).astype(np.float32)
        w = rng.standard_normal(ws).astype(np.float32)
        cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
        b = rng.standard_normal(cout).astype(np.float32)
        y, plan = emul.conv(x, w, b, transposed=tr, flags=flags, **kw)
        ref = (O.fft_conv_transpose if tr else O.fft_conv)(x, w, b, **kw)
        err = float(np.abs(y - ref).max() / np.abs(ref).max())
        names = [l.split()[1] for l in plan.describe().splitlines() if l.strip().startswith("launch")]
        print(f"{name}: {names} rel err {err:.2e}", flush=True)
        assert err < 1e-4, name
    print("all families ok")


if __name__ == "__main__":
    main()
