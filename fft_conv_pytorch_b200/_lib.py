"""ctypes binding of the C ABI in include/fftconv_b200.h (no torch types cross this boundary).

The shared library is built in-tree by ``build.py`` (nvcc, sm_100a) as ``libfftconv_b200.so``. There is no CPU
fallback: if the library is missing, importing the operators raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Sequence

FC_MAX_ND = 3
PAD_MODES = {"constant": 0, "zeros": 0, "reflect": 1, "replicate": 2, "circular": 3}
FC_FLAG_NO_FUSED = 1
FC_FLAG_NO_POLYPHASE = 2
FC_FLAG_NO_FAST_R2C = 4
FC_FLAG_NO_FAST_C2R = 8
FC_FLAG_NO_FUSED_MID = 16
FC_FLAG_NO_TC = 32
FC_FLAG_NO_FAST_C2C = 64
FC_FLAG_NO_SEGMENT = 128
FC_FLAG_NO_PAIR = 256
FC_FLAG_PAIR = 512
FC_FLAG_NO_YSTAGE = 1024
FC_FLAG_NO_ROW_FILL = 2048
FC_FLAG_NO_STREAM = 4096
FC_FLAG_STREAM_R2C = 8192
FC_FLAG_SEGMENT = 16384
FC_FLAG_NO_SHORT_SPLIT = 32768

_I3 = ctypes.c_int32 * FC_MAX_ND


class FcProblem(ctypes.Structure):
    _fields_ = [
        ("ndim", ctypes.c_int32),
        ("transposed", ctypes.c_int32),
        ("batch", ctypes.c_int32),
        ("cin", ctypes.c_int32),
        ("cout", ctypes.c_int32),
        ("groups", ctypes.c_int32),
        ("in_size", _I3),
        ("kernel_size", _I3),
        ("stride", _I3),
        ("padding", _I3),
        ("dilation", _I3),
        ("output_padding", _I3),
        ("padding_mode", ctypes.c_int32),
        ("threads", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class FcPlanInfo(ctypes.Structure):
    _fields_ = [
        ("ndim", ctypes.c_int32),
        ("out_size", _I3),
        ("fft_size", _I3),
        ("n_launches", ctypes.c_int32),
        ("n_launches_kspec", ctypes.c_int32),
        ("fused", ctypes.c_int32),
        ("segments", ctypes.c_int32),
        ("bins", ctypes.c_int64),
        ("out_elems", ctypes.c_int64),
        ("xspec_bytes", ctypes.c_int64),
        ("kspec_bytes", ctypes.c_int64),
        ("yspec_bytes", ctypes.c_int64),
        ("workspace_bytes", ctypes.c_int64),
        ("const_bytes", ctypes.c_int64),
        ("algo_bytes_s1", ctypes.c_int64),
        ("algo_bytes_s2", ctypes.c_int64),
        ("algo_bytes_s3", ctypes.c_int64),
        ("algo_bytes_s4", ctypes.c_int64),
        ("kspec_workspace_bytes", ctypes.c_int64),
        ("tensor_core", ctypes.c_int32),
        ("reserved2", ctypes.c_int32),
    ]


_P = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/fftconv_b200.h declares
SYMBOLS = {
    "fc_last_error": (ctypes.c_char_p, []),
    "fc_version": (ctypes.c_char_p, []),
    "fc_plan_create": (ctypes.c_int, [ctypes.POINTER(_P), ctypes.POINTER(FcProblem)]),
    "fc_plan_destroy": (None, [_P]),
    "fc_plan_get_info": (ctypes.c_int, [_P, ctypes.POINTER(FcPlanInfo)]),
    "fc_plan_describe": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_size_t]),
    "fc_plan_init_const": (ctypes.c_int, [_P, _P, _P]),
    "fc_signal_spectrum": (ctypes.c_int, [_P, _P, _P, _P, _P, _P]),
    "fc_kernel_spectrum": (ctypes.c_int, [_P, _P, _P, _P, _P, _P]),
    "fc_contract": (ctypes.c_int, [_P, _P, _P, _P, _P]),
    "fc_inverse": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "fc_conv": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "fc_conv_host": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fc_conv_profiled": (ctypes.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "fc_plan_launch_info": (ctypes.c_int, [_P, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_int64)]),
    "fc_tc_supported": (ctypes.c_int, [ctypes.c_int64] * 4),
    "fc_tc_scratch_bytes": (ctypes.c_int64, [ctypes.c_int64] * 5),
    "fc_tc_prepare_kernel": (ctypes.c_int, [_P, _P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _P]),
    "fc_tc_complex_matmul": (ctypes.c_int, [_P, _P, _P, _P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _P]),
    "fc_complex_matmul": (ctypes.c_int, [_P, _P, _P, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, _P]),
}

LIB_NAME = "libfftconv_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


def bind(cdll: ctypes.CDLL) -> ctypes.CDLL:
    """Attach signatures; raises AttributeError if a declared symbol is not exported."""
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib: Optional[ctypes.CDLL] = None


def load(path: Optional[str] = None) -> ctypes.CDLL:
    """Load the CUDA library (once). No fallback: a missing library is an error."""
    global _lib
    if path is None and _lib is not None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"{LIB_NAME} not found at {p}: build it with `python -m fft_conv_pytorch_b200.build` (nvcc, sm_100a). "
            "fft_conv_pytorch_b200 has no CPU fallback."
        )
    lib = bind(ctypes.CDLL(p))
    if path is None:
        _lib = lib
    return lib


class FcError(RuntimeError):
    pass


def check(lib: ctypes.CDLL, rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = (lib.fc_last_error() or b"").decode()
    if rc < 0:
        raise ValueError(f"{what}: {msg}")
    raise FcError(f"{what}: CUDA error {rc}: {msg}")


def make_problem(
    transposed: bool,
    batch: int,
    cin: int,
    cout: int,
    groups: int,
    in_size: Sequence[int],
    kernel_size: Sequence[int],
    stride: Sequence[int],
    padding: Sequence[int],
    dilation: Sequence[int],
    output_padding: Optional[Sequence[int]] = None,
    padding_mode: str = "constant",
    threads: int = 0,
    flags: int = 0,
) -> FcProblem:
    n = len(in_size)
    if padding_mode not in PAD_MODES:
        raise ValueError(f"Unknown padding_mode {padding_mode!r}; expected one of {sorted(PAD_MODES)}")
    p = FcProblem()
    p.ndim, p.transposed = n, int(bool(transposed))
    p.batch, p.cin, p.cout, p.groups = int(batch), int(cin), int(cout), int(groups)
    opad = output_padding if output_padding is not None else (0,) * n
    for i in range(min(n, FC_MAX_ND)):
        p.in_size[i], p.kernel_size[i] = int(in_size[i]), int(kernel_size[i])
        p.stride[i], p.padding[i], p.dilation[i] = int(stride[i]), int(padding[i]), int(dilation[i])
        p.output_padding[i] = int(opad[i])
    p.padding_mode = PAD_MODES[padding_mode]
    p.threads, p.flags = int(threads), int(flags)
    return p


class Plan:
    """Host-side plan handle (immutable once created)."""

    def __init__(self, lib: ctypes.CDLL, problem: FcProblem):
        self.lib = lib
        self.problem = problem
        h = _P()
        check(lib, lib.fc_plan_create(ctypes.byref(h), ctypes.byref(problem)), "fc_plan_create")
        self.handle = h
        self.info = FcPlanInfo()
        check(lib, lib.fc_plan_get_info(h, ctypes.byref(self.info)), "fc_plan_get_info")

    @property
    def out_size(self):
        return tuple(self.info.out_size[i] for i in range(self.info.ndim))

    @property
    def fft_size(self):
        return tuple(self.info.fft_size[i] for i in range(self.info.ndim))

    def describe(self) -> str:
        buf = ctypes.create_string_buffer(16384)
        self.lib.fc_plan_describe(self.handle, buf, len(buf))
        return buf.value.decode()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.fc_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
