#!/usr/bin/env python3
"""Generate tests/golden/ref_cases.npz by running the UNMODIFIED reference (authoring container only).

Usage (in the authoring container, where /root/reference exists):
    python tests/golden/make_golden.py

For every case it stores the inputs, the reference's own ``fft_conv`` / ``fft_conv_transpose`` output (fp32; its
fp64 run is checked here against the direct result to 1e-9 and not stored) and torch's direct ``F.conv{n}d`` / ``F.conv_transpose{n}d`` output (computed in fp64) — the quantity the reference's tests
pin the path to (reference tests/test_functional.py:56-59). The GPU box has no /root/reference; tests read only
the committed .npz.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("FFTCONV_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
from fft_conv_pytorch.functional import fft_conv, fft_conv_transpose  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    out = []
    # the reference's own test grid, thinned (reference tests/test_functional.py:11-20)
    for ndim in (1, 2, 3):
        for size in (7, 8):
            for (cin, cout, g) in ((2, 2, 1), (3, 3, 3), (2, 2, 2), (3, 2, 1)):
                for (k, p, s, d) in ((2, 0, 1, 1), (3, 1, 2, 2), (3, 1, 1, 2), (2, 1, 2, 1)):
                    out.append(dict(kind="fwd", x=(2, cin) + (size,) * ndim, w=(cout, cin // g) + (k,) * ndim,
                                    kw=dict(stride=s, padding=p, dilation=d, groups=g)))
    # transposed grid (reference tests/test_functional_transpose.py:11-21, 67-68)
    for ndim in (1, 2, 3):
        for size in (7, 8):
            for (cin, cout, g) in ((2, 2, 1), (3, 3, 3), (2, 2, 2), (2, 3, 1)):
                for (k, p, op, s, d) in ((2, 0, 0, 1, 1), (3, 1, 1, 3, 3), (3, 1, 0, 2, 2), (2, 1, 2, 4, 3), (3, 0, 1, 2, 1)):
                    if ndim == 3 and s > 2:
                        continue  # keeps the fixture file small (3-d outputs grow as stride^3)
                    out.append(dict(kind="tr", x=(2, cin) + (size,) * ndim, w=(cin, cout // g) + (k,) * ndim,
                                    kw=dict(stride=s, padding=p, output_padding=op, dilation=d, groups=g)))
    # behaviours the reference supports but never tests
    for mode in ("constant", "reflect", "replicate", "circular"):
        out.append(dict(kind="fwd", x=(2, 4, 9, 12), w=(6, 2, 3, 4), kw=dict(stride=(2, 1), padding=(2, 3), dilation=(1, 2), groups=2, padding_mode=mode)))
        out.append(dict(kind="fwd", x=(1, 2, 21), w=(2, 2, 5), kw=dict(stride=1, padding=4, dilation=1, groups=1, padding_mode=mode)))
    out.append(dict(kind="fwd", x=(3, 2, 33), w=(4, 2, 1), kw=dict(), bias=False))
    out.append(dict(kind="fwd", x=(1, 1, 5, 6, 7), w=(2, 1, 2, 3, 1), kw=dict(stride=(1, 2, 1), padding=(1, 0, 2), dilation=(2, 1, 3))))
    out.append(dict(kind="tr", x=(2, 4, 5, 6), w=(4, 3, 3, 2), kw=dict(stride=(2, 3), padding=(1, 0), output_padding=(1, 2), dilation=(1, 2), groups=2)))
    out.append(dict(kind="tr", x=(1, 2, 6, 5, 4), w=(2, 2, 2, 3, 2), kw=dict(stride=(1, 2, 2), padding=(0, 1, 1), output_padding=(0, 1, 0), dilation=(2, 2, 1))))
    # medium sizes exercising multi-stage transforms and the four-step 1-d layout (N > 8192)
    out.append(dict(kind="fwd", x=(1, 2, 300), w=(2, 2, 31), kw=dict(padding=3)))
    out.append(dict(kind="fwd", x=(1, 2, 9000), w=(2, 2, 65), kw=dict()))
    out.append(dict(kind="fwd", x=(1, 2, 20000), w=(2, 1, 129), kw=dict(groups=2, stride=3, padding=7, padding_mode="reflect")))
    out.append(dict(kind="tr", x=(1, 2, 5000), w=(2, 2, 33), kw=dict(stride=2, dilation=2, padding=5, output_padding=1)))
    out.append(dict(kind="fwd", x=(1, 2, 70, 90), w=(3, 2, 9, 7), kw=dict(padding=(2, 1))))
    out.append(dict(kind="tr", x=(1, 2, 40, 30), w=(2, 2, 5, 5), kw=dict(stride=2, dilation=2)))
    out.append(dict(kind="fwd", x=(1, 2, 20, 18, 22), w=(2, 2, 5, 3, 4), kw=dict()))
    return out


def main():
    rng = np.random.RandomState(1234)
    store = {}
    meta = []
    cs = cases()
    for idx, c in enumerate(cs):
        x = rng.standard_normal(c["x"]).astype(np.float32)
        w = rng.standard_normal(c["w"]).astype(np.float32)
        kw = c["kw"]
        transposed = c["kind"] == "tr"
        cout = c["w"][1] * kw.get("groups", 1) if transposed else c["w"][0]
        b = rng.standard_normal(cout).astype(np.float32) if c.get("bias", True) else None
        tx, tw = torch.from_numpy(x), torch.from_numpy(w)
        tb = None if b is None else torch.from_numpy(b)
        fn = fft_conv_transpose if transposed else fft_conv
        y_ref32 = fn(tx, tw, tb, **kw).contiguous().numpy()
        dbl = lambda t: None if t is None else t.double()
        y_ref64 = fn(tx.double(), tw.double(), dbl(tb), **kw).contiguous().numpy()
        n = x.ndim - 2
        dkw = dict(kw)
        mode = dkw.pop("padding_mode", "constant")
        if transposed:
            y_dir = getattr(F, f"conv_transpose{n}d")(tx.double(), tw.double(), dbl(tb), **dkw).numpy()
        else:
            pads = dkw.pop("padding", 0)
            pads = (pads,) * n if isinstance(pads, int) else tuple(pads)
            xp = tx.double()
            if any(pads):
                xp = F.pad(xp, [p for p in pads[::-1] for _ in range(2)], mode=mode)
            y_dir = getattr(F, f"conv{n}d")(xp, tw.double(), dbl(tb), **dkw).numpy()
        short = transposed and y_ref64.shape != y_dir.shape  # reference bug A.5 (K=1, output_padding > padding)
        store[f"x{idx}"] = x
        store[f"w{idx}"] = w
        if b is not None:
            store[f"b{idx}"] = b
        store[f"ref32_{idx}"] = y_ref32
        store[f"dir_{idx}"] = y_dir.astype(np.float32)  # computed in fp64, stored rounded to fp32
        meta.append(dict(kind=c["kind"], kw={k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()},
                         bias=b is not None, ref_short=bool(short)))
        if not short:
            err = np.abs(y_ref64 - y_dir).max()
            assert err < 1e-9, (idx, c, err)
    store["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    path = os.path.join(HERE, "ref_cases.npz")
    np.savez_compressed(path, **store)
    print(f"wrote {len(cs)} cases to {path} ({os.path.getsize(path) / 1e6:.2f} MB); torch {torch.__version__}")


if __name__ == "__main__":
    main()
