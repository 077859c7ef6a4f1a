"""ctypes wrapper of oracle/direct_conv.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see direct_conv.c)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdirect_conv.so")
_PAD = {"constant": 0, "zeros": 0, "reflect": 1, "replicate": 2, "circular": 3}


def build() -> str:
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "direct_conv.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def _lib():
    lib = ctypes.CDLL(build())
    return lib


def _tup3(v, n, fill):
    v = tuple(v) if hasattr(v, "__iter__") else (v,) * n
    return (ctypes.c_int * 3)(*([fill] * (3 - n) + list(v)))


def direct_conv(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1, padding_mode="constant"):
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    n = x.ndim - 2
    L, K = _tup3(x.shape[2:], n, 1), _tup3(w.shape[2:], n, 1)
    s, p, d = _tup3(stride, n, 1), _tup3(padding, n, 0), _tup3(dilation, n, 1)
    out_sp = [(L[a] + 2 * p[a] - (K[a] - 1) * d[a] - 1) // s[a] + 1 for a in range(3 - n, 3)]
    y = np.empty((x.shape[0], w.shape[0], *out_sp), dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = _lib().fc_oracle_direct_conv(
        x.ctypes.data_as(fp), w.ctypes.data_as(fp), None if b is None else b.ctypes.data_as(fp), y.ctypes.data_as(fp),
        x.shape[0], x.shape[1], w.shape[0], groups, L, K, s, p, d, _PAD[padding_mode])
    if rc:
        raise ValueError("direct_conv: kernel larger than padded signal")
    return y


def direct_conv_transpose(x, w, bias=None, stride=1, padding=0, output_padding=0, dilation=1, groups=1):
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    n = x.ndim - 2
    L, K = _tup3(x.shape[2:], n, 1), _tup3(w.shape[2:], n, 1)
    s, p, d, op = _tup3(stride, n, 1), _tup3(padding, n, 0), _tup3(dilation, n, 1), _tup3(output_padding, n, 0)
    out_sp = [(L[a] - 1) * s[a] - 2 * p[a] + d[a] * (K[a] - 1) + op[a] + 1 for a in range(3 - n, 3)]
    cout = w.shape[1] * groups
    y = np.empty((x.shape[0], cout, *out_sp), dtype=np.float32)
    b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    rc = _lib().fc_oracle_direct_conv_transpose(
        x.ctypes.data_as(fp), w.ctypes.data_as(fp), None if b is None else b.ctypes.data_as(fp), y.ctypes.data_as(fp),
        x.shape[0], x.shape[1], cout, groups, L, K, s, p, d, op)
    if rc:
        raise ValueError("direct_conv_transpose: empty output")
    return y
