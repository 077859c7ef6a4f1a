"""TEST INFRASTRUCTURE ONLY: run the generic CUDA kernels on host threads through the same C ABI.

`libfftconv_emul.so` is fc_api.cu + fc_plan.cpp + fc_kernels.cuh compiled as host C++ against cuda_shim.h.
It lets the `-m "not gpu"` suite check the kernels' index algebra (passes, maps, layouts, twiddles) against the
oracle in a container without a GPU. The product package never loads it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from fft_conv_pytorch_b200 import _lib as L

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.environ.get("FFTCONV_EMUL_SO") or os.path.join(_HERE, "libfftconv_emul.so")  # (a sanitizer build: scripts/emul_sanitize.sh)
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.environ.get("FFTCONV_EMUL_SO"):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _lib = L.bind(ctypes.CDLL(_SO))
    return _lib


def _ptr(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _tup(v, n):
    return tuple(v) if hasattr(v, "__iter__") else (v,) * n


def plan_for(x_shape, w_shape, transposed=False, stride=1, padding=0, dilation=1, groups=1, output_padding=0,
             padding_mode="constant", threads=64, flags=0):
    n = len(x_shape) - 2
    cout = w_shape[1] * groups if transposed else w_shape[0]
    prob = L.make_problem(transposed, x_shape[0], x_shape[1], cout, groups, x_shape[2:], w_shape[2:], _tup(stride, n),
                          _tup(padding, n), _tup(dilation, n), _tup(output_padding, n), padding_mode, threads, flags)
    return L.Plan(lib(), prob)


def conv(x, w, b=None, transposed=False, staged=False, **kw):
    """Full pipeline on numpy arrays. staged=True calls the four stage entry points instead of fc_conv."""
    lb = lib()
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    b = None if b is None else np.ascontiguousarray(b, np.float32)
    if staged:  # the stage calls need the unsegmented layout and the channel-major kernel spectrum
        kw = dict(kw, flags=kw.get("flags", 0) | L.FC_FLAG_NO_SEGMENT | L.FC_FLAG_NO_FUSED_MID)
    plan = plan_for(x.shape, w.shape, transposed, **kw)
    info = plan.info
    const = np.zeros(info.const_bytes, np.uint8)
    L.check(lb, lb.fc_plan_init_const(plan.handle, _ptr(const), None), "init_const")
    ws = np.zeros(info.workspace_bytes, np.uint8)
    kspec = np.zeros(info.kspec_bytes // 4, np.float32)
    L.check(lb, lb.fc_kernel_spectrum(plan.handle, _ptr(const), _ptr(w), _ptr(kspec), _ptr(ws), None), "kernel_spectrum")
    cout = plan.problem.cout
    y = np.full((x.shape[0], cout) + plan.out_size, np.nan, np.float32)
    if staged:
        xs = np.zeros(info.xspec_bytes // 4, np.float32)
        ys = np.zeros(info.yspec_bytes // 4, np.float32)
        L.check(lb, lb.fc_signal_spectrum(plan.handle, _ptr(const), _ptr(x), _ptr(xs), _ptr(ws), None), "signal_spectrum")
        L.check(lb, lb.fc_contract(plan.handle, _ptr(xs), _ptr(kspec), _ptr(ys), None), "contract")
        L.check(lb, lb.fc_inverse(plan.handle, _ptr(const), _ptr(ys), _ptr(b), _ptr(y), _ptr(ws), None), "inverse")
    else:
        L.check(lb, lb.fc_conv(plan.handle, _ptr(const), _ptr(x), _ptr(kspec), _ptr(b), _ptr(y), _ptr(ws), None), "conv")
    return y, plan
