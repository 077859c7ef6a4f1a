"""Shared test helpers: golden fixtures of the unmodified reference and the parity metric."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Golden:
    """tests/golden/ref_cases.npz — outputs of the unmodified reference (see tests/golden/make_golden.py)."""

    def __init__(self):
        self.npz = np.load(os.path.join(ROOT, "tests", "golden", "ref_cases.npz"))
        self.meta = json.loads(bytes(self.npz["meta"]).decode())

    def __len__(self):
        return len(self.meta)

    def case(self, i):
        m = self.meta[i]
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in m["kw"].items()}
        return dict(
            idx=i, transposed=m["kind"] == "tr", kw=kw, x=self.npz[f"x{i}"], w=self.npz[f"w{i}"],
            b=self.npz[f"b{i}"] if m["bias"] else None, ref32=self.npz[f"ref32_{i}"], direct=self.npz[f"dir_{i}"],
            ref_short=m["ref_short"],
        )


_golden = None


def golden():
    global _golden
    if _golden is None:
        _golden = Golden()
    return _golden


def rel_err(y, ref):
    """SURVEY §8c parity metric: max|y - ref| / max|ref|."""
    y = np.asarray(y, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(y - ref).max() / max(np.abs(ref).max(), 1e-30))


def spot_check(x, w, b, y, n=100, transposed=False, stride=1, padding=0, dilation=1, groups=1, seed=0):
    """Check `n` random output elements of a full-size result `y` (torch tensor, any device) against the
    convolution's definition evaluated in float64 (zero padding mode).
    Returns (max abs error / max|y|, n). Used where a full direct convolution is intractable (BASELINE c4, c5)."""
    rng = np.random.RandomState(seed)
    nd = x.ndim - 2
    tup = lambda v: tuple(v) if hasattr(v, "__iter__") else (v,) * nd
    st, pd, dl = tup(stride), tup(padding), tup(dilation)
    B, cin = x.shape[:2]
    cout = y.shape[1]
    out_sp = tuple(y.shape[2:])
    K = w.shape[2:]
    L = x.shape[2:]
    ig, og = cin // groups, cout // groups
    ymax = float(y.abs().max())
    worst = 0.0
    for _ in range(n):
        bb, o = rng.randint(B), rng.randint(cout)
        j = [rng.randint(s) for s in out_sp]
        g, ol = o // og, o % og
        idx, masks = [], []
        for a in range(nd):
            m = np.arange(K[a])
            if not transposed:
                q = j[a] * st[a] + m * dl[a] - pd[a]
                ok = (q >= 0) & (q < L[a])
            else:
                t = j[a] + pd[a] - m * dl[a]
                ok = (t >= 0) & (t % st[a] == 0) & (t // st[a] < L[a])
                q = t // st[a]
            idx.append(np.where(ok, q, 0))
            masks.append(ok)
        xs = x[bb, g * ig:(g + 1) * ig]
        patch = xs[np.ix_(np.arange(ig), *idx)].astype(np.float64)
        mask = masks[0].astype(np.float64)
        for a in range(1, nd):
            mask = np.multiply.outer(mask, masks[a].astype(np.float64))
        wsel = (w[g * ig:(g + 1) * ig, ol] if transposed else w[o]).astype(np.float64)
        val = float((patch * mask[None] * wsel).sum()) + (0.0 if b is None else float(b[o]))
        got = float(y[(bb, o) + tuple(j)])
        worst = max(worst, abs(got - val))
    return worst / max(ymax, 1e-30), n
