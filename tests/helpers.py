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
