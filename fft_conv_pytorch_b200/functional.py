"""Functional API: ``fft_conv``, ``fft_conv_transpose``, ``complex_matmul`` with the reference's signatures.

Mirrors fft_conv_pytorch/functional.py of the reference (``fft_conv`` :19-28, ``fft_conv_transpose`` :92-101,
``complex_matmul`` :11-16): same argument names, meaning and defaults. The work is done by the sm_100a kernels of
``libfftconv_b200.so`` through the C ABI of include/fftconv_b200.h; torch only supplies device memory and the
current stream. There is no CPU implementation: CPU tensors are staged to the GPU and back (the "host buffer"
path, ``fc_conv_host``), and without a CUDA device or without the library every call raises.

Differences from the reference, all documented in DESIGN.md:
  * the result is a contiguous tensor (the reference returns a strided view of its irfftn buffer);
  * float32 only (the reference also runs float64 on CPU);
  * malformed problems that the reference silently mis-computes (kernel larger than the padded signal,
    ``Cin % groups != 0`` ...) raise ``ValueError``.
"""
from __future__ import annotations

import ctypes
import os
import threading
import weakref
from collections import OrderedDict
from typing import Iterable, Optional, Tuple, Union

import torch
from torch import Tensor

from . import _lib as L
from .utils import to_ntuple

IntOrSeq = Union[int, Iterable[int]]

_lock = threading.RLock()
_plans: "OrderedDict[tuple, _PlanEntry]" = OrderedDict()
_PLAN_CACHE_MAX = 64
_kspec_cache: "OrderedDict[tuple, tuple]" = OrderedDict()
_kspec_cache_bytes = 0
_kspec_cap = {}  # device index -> byte cap of the spectrum cache (a quarter of the device memory)
_launch_counter = 0  # kernels queued by this module (bench.py reports it)
# FC_FLAG_* bits OR-ed into every plan (set_default_flags: A/B timing of the specialised kernels, 1 = generic kernels only)
_DEFAULT_FLAGS = 0


def set_default_flags(flags: int) -> int:
    """OR `flags` (FC_FLAG_*) into every plan created from now on; returns the previous value. For tests and A/B
    timing; no environment variable is consulted anywhere in the call path."""
    global _DEFAULT_FLAGS
    old, _DEFAULT_FLAGS = _DEFAULT_FLAGS, int(flags)
    return old


def launches() -> int:
    return _launch_counter


class _PlanEntry:
    """A host plan plus its per-device constant table."""

    def __init__(self, plan: L.Plan):
        self.plan = plan
        self.const = {}  # device index -> (uint8 tensor, (event, stream) of the fill)

    def const_for(self, device: torch.device) -> Tensor:
        idx = device.index if device.index is not None else torch.cuda.current_device()
        hit = self.const.get(idx)
        if hit is not None:
            return _serve_spectrum(hit[0], hit[1], device)
        lib = self.plan.lib
        t = torch.empty(int(self.plan.info.const_bytes), dtype=torch.uint8, device=device)
        cur = torch.cuda.current_stream(device)
        L.check(lib, lib.fc_plan_init_const(self.plan.handle, ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(cur.cuda_stream)), "fc_plan_init_const")
        ev = torch.cuda.Event()
        ev.record(cur)
        self.const[idx] = (t, (ev, cur.cuda_stream))
        return t


def _require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("fft_conv_pytorch_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")


def get_plan(
    transposed: bool,
    batch: int,
    cin: int,
    cout: int,
    groups: int,
    in_size: Tuple[int, ...],
    kernel_size: Tuple[int, ...],
    stride: Tuple[int, ...],
    padding: Tuple[int, ...],
    dilation: Tuple[int, ...],
    output_padding: Tuple[int, ...],
    padding_mode: str,
    flags: int = 0,
    threads: int = 0,
) -> _PlanEntry:
    flags |= _DEFAULT_FLAGS
    key = (transposed, batch, cin, cout, groups, in_size, kernel_size, stride, padding, dilation, output_padding, padding_mode, flags, threads)
    with _lock:
        e = _plans.get(key)
        if e is not None:
            _plans.move_to_end(key)
            return e
        prob = L.make_problem(transposed, batch, cin, cout, groups, in_size, kernel_size, stride, padding, dilation, output_padding,
                              padding_mode, threads, flags)
        e = _PlanEntry(L.Plan(L.load(), prob))
        _plans[key] = e
        while len(_plans) > _PLAN_CACHE_MAX:
            _, old = _plans.popitem(last=False)
            _drop_plan_spectra(old.plan)
        return e


def _drop_plan_spectra(plan) -> None:
    """An evicted plan takes its cached kernel spectra with it (they are laid out for that plan only)."""
    with _lock:
        for k in [k for k, v in _kspec_cache.items() if v[4] is plan]:
            _drop_kspec(k)


def _kspec_cache_cap(dev_idx: int) -> int:
    cap = _kspec_cap.get(dev_idx)
    if cap is None:
        try:
            cap = int(torch.cuda.get_device_properties(dev_idx).total_memory) // 4
        except Exception:
            cap = 8 << 30
        _kspec_cap[dev_idx] = cap
    return cap


def _serve_spectrum(spec: Tensor, built, device: torch.device) -> Tensor:
    """A cached spectrum was produced on the stream that was current at its first use; a consumer on another stream
    waits for that work (one event wait, nothing when the stream is the same)."""
    ev, stream_id = built
    if ev is None:
        return spec
    cur = torch.cuda.current_stream(device)
    if cur.cuda_stream != stream_id and not torch.cuda.is_current_stream_capturing():
        cur.wait_event(ev)
    return spec


def clear_caches(plans: bool = True) -> None:
    """Drop the cached kernel spectra and (unless plans=False) the plans."""
    global _kspec_cache_bytes
    with _lock:
        if plans:
            _plans.clear()
        _kspec_cache.clear()
        _kspec_cache_bytes = 0


def _ptr(t: Optional[Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _check_tensor(name: str, t: Tensor) -> None:
    if not isinstance(t, Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype}); fft_conv_pytorch_b200 computes in fp32 only")


def kernel_spectrum(entry: _PlanEntry, kernel: Tensor, device: torch.device, use_cache: bool = True) -> Tensor:
    """Cached spectrum of the weight tensor (stage 2): dilation scatter / transposed regroup + transform +
    conjugate + 1/N scale, in the layout the contraction reads.

    The cache entry is tied to the weight *object* (weak reference) and to its storage pointer and version
    counter, so an optimizer step, an in-place edit or a re-used allocation can never serve a stale spectrum.
    It is a derived, non-persistent object: never part of a ``state_dict``."""
    global _kspec_cache_bytes, _launch_counter
    plan = entry.plan
    dev_idx = device.index if device.index is not None else (torch.cuda.current_device() if device.type == "cuda" else -1)
    # id(plan) alone could be recycled for a later plan once this one is evicted and freed: the entry therefore holds
    # the plan itself (so the id stays taken while the entry lives) and a hit also requires `hit_plan is plan`
    key = (id(plan), id(kernel), dev_idx)
    if use_cache:
        with _lock:
            hit = _kspec_cache.get(key)
            if hit is not None:
                ref, ver, ptr, spec, hit_plan, built = hit
                if hit_plan is plan and ref() is kernel and ver == kernel._version and ptr == kernel.data_ptr():
                    _kspec_cache.move_to_end(key)
                    return _serve_spectrum(spec, built, device)
                _drop_kspec(key)
    lib = plan.lib
    const = entry.const_for(device)
    kspec = torch.empty(int(plan.info.kspec_bytes) // 4, dtype=torch.float32, device=device)
    ws = torch.empty(int(plan.info.kspec_workspace_bytes), dtype=torch.uint8, device=device)
    cur = torch.cuda.current_stream(device)
    stream = cur.cuda_stream
    w = kernel.detach().to(device=device).contiguous()
    L.check(lib, lib.fc_kernel_spectrum(plan.handle, _ptr(const), _ptr(w), _ptr(kspec), _ptr(ws), ctypes.c_void_p(stream)), "fc_kernel_spectrum")
    _launch_counter += int(plan.info.n_launches_kspec)
    if use_cache:
        with _lock:
            try:
                ref = weakref.ref(kernel, lambda _r, k=key: _drop_kspec(k))
            except TypeError:
                return kspec
            ev = torch.cuda.Event()
            ev.record(cur)
            _kspec_cache[key] = (ref, kernel._version, kernel.data_ptr(), kspec, plan, (ev, stream))
            _kspec_cache_bytes += kspec.numel() * 4
            cap = _kspec_cache_cap(dev_idx)
            while _kspec_cache_bytes > cap and len(_kspec_cache) > 1:
                _drop_kspec(next(iter(_kspec_cache)))
    return kspec


def install_kernel_spectrum(entry: _PlanEntry, kernel: Tensor, device: torch.device, kspec: Tensor) -> None:
    """Put an externally produced spectrum of `kernel` (same plan layout, e.g. received from another rank:
    ``dist.broadcast_kernel_spectrum``) into the cache; it is served until the weight changes, like a local one."""
    global _kspec_cache_bytes
    plan = entry.plan
    if kspec.numel() * kspec.element_size() != int(plan.info.kspec_bytes):
        raise ValueError(f"kernel spectrum has {kspec.numel() * kspec.element_size()} bytes, the plan expects {int(plan.info.kspec_bytes)}")
    dev_idx = device.index if device.index is not None else (torch.cuda.current_device() if device.type == "cuda" else -1)
    key = (id(plan), id(kernel), dev_idx)
    with _lock:
        _drop_kspec(key)
        ref = weakref.ref(kernel, lambda _r, k=key: _drop_kspec(k))
        built = (None, None)
        if device.type == "cuda":
            cur = torch.cuda.current_stream(device)
            ev = torch.cuda.Event()
            ev.record(cur)
            built = (ev, cur.cuda_stream)
        _kspec_cache[key] = (ref, kernel._version, kernel.data_ptr(), kspec, plan, built)
        _kspec_cache_bytes += kspec.numel() * 4


def _drop_kspec(key) -> None:
    global _kspec_cache_bytes
    with _lock:
        old = _kspec_cache.pop(key, None)
        if old is not None:
            _kspec_cache_bytes -= old[3].numel() * 4


def _run(transposed: bool, signal: Tensor, kernel: Tensor, bias: Optional[Tensor], stride, padding, output_padding, dilation,
         groups: int, padding_mode: str, flags: int = 0) -> Tensor:
    global _launch_counter
    _check_tensor("signal", signal)
    _check_tensor("kernel", kernel)
    if bias is not None:
        _check_tensor("bias", bias)
    n = signal.ndim - 2
    if n < 1:
        raise ValueError(f"signal must have shape (batch, channels, *spatial); got {tuple(signal.shape)}")
    padding_ = to_ntuple(padding, n)
    stride_ = to_ntuple(stride, n)
    dilation_ = to_ntuple(dilation, n)
    opad_ = to_ntuple(output_padding, n)
    if kernel.ndim != signal.ndim:
        raise ValueError(f"kernel must have {signal.ndim} dims like the signal; got {kernel.ndim}")
    if not isinstance(groups, int) or groups < 1:
        raise ValueError(f"groups must be a positive int, got {groups!r}")
    B, cin = int(signal.shape[0]), int(signal.shape[1])
    if transposed:
        if kernel.shape[0] != cin:
            raise ValueError(f"transposed kernel must have shape (in_channels={cin}, out_channels/groups, ...); got {tuple(kernel.shape)}")
        cout = int(kernel.shape[1]) * groups
    else:
        cout = int(kernel.shape[0])
        if cin % groups or kernel.shape[1] != cin // groups:
            raise ValueError(
                f"kernel must have shape (out_channels, in_channels/groups={cin}/{groups}, ...); got {tuple(kernel.shape)}")
    if bias is not None and tuple(bias.shape) != (cout,):
        raise ValueError(f"bias must have shape ({cout},); got {tuple(bias.shape)}")
    if torch.is_grad_enabled() and (signal.requires_grad or kernel.requires_grad or (bias is not None and bias.requires_grad)):
        from .autograd import conv_with_grad  # local import: autograd wraps this module

        return conv_with_grad(transposed, signal, kernel, bias, stride_, padding_, opad_, dilation_, groups, padding_mode)

    _require_cuda()
    if B == 0:  # empty batch: shape algebra only (the plan of a one-sample problem gives the output extents)
        e1 = get_plan(transposed, 1, cin, cout, groups, tuple(int(s) for s in signal.shape[2:]), tuple(int(s) for s in kernel.shape[2:]),
                      stride_, padding_, dilation_, opad_, padding_mode, flags)
        return signal.new_empty((0, cout) + e1.plan.out_size)
    entry = get_plan(transposed, B, cin, cout, groups, tuple(int(s) for s in signal.shape[2:]), tuple(int(s) for s in kernel.shape[2:]),
                     stride_, padding_, dilation_, opad_, padding_mode, flags)
    plan = entry.plan
    lib = plan.lib
    host_path = not signal.is_cuda
    dev = kernel.device if kernel.is_cuda else (signal.device if signal.is_cuda else torch.device("cuda", torch.cuda.current_device()))
    if signal.is_cuda and kernel.is_cuda and signal.device != kernel.device:
        raise ValueError(f"signal is on {signal.device} but kernel is on {kernel.device}")
    with torch.cuda.device(dev):
        b_dev = None if bias is None else (bias.detach().contiguous() if bias.is_cuda else bias.detach().to(dev))
        const = entry.const_for(dev)
        kspec = kernel_spectrum(entry, kernel, dev)
        ws = torch.empty(int(plan.info.workspace_bytes), dtype=torch.uint8, device=dev)
        out_shape = (B, cout) + plan.out_size
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if not host_path:
            x = signal.detach().contiguous()
            y = torch.empty(out_shape, dtype=torch.float32, device=dev)
            L.check(lib, lib.fc_conv(plan.handle, _ptr(const), _ptr(x), _ptr(kspec), _ptr(b_dev), _ptr(y), _ptr(ws), stream), "fc_conv")
            _launch_counter += int(plan.info.n_launches)
            return y
        # host buffers in, host buffers out
        x_host = signal.detach().contiguous()
        if not x_host.is_pinned():
            x_host = x_host.pin_memory()
        y_host = torch.empty(out_shape, dtype=torch.float32, pin_memory=True)
        x_stage = torch.empty(x_host.shape, dtype=torch.float32, device=dev)
        y_stage = torch.empty(out_shape, dtype=torch.float32, device=dev)
        n_chunks = min(B, _HOST_PIPELINE_CHUNKS)
        if n_chunks < 2:
            # one shot: H2D + kernels + D2H on the current stream (fc_conv_host)
            L.check(lib, lib.fc_conv_host(plan.handle, _ptr(const), _ptr(x_host), _ptr(x_stage), _ptr(kspec), _ptr(b_dev), _ptr(y_stage),
                                          _ptr(y_host), _ptr(ws), stream), "fc_conv_host")
            _launch_counter += int(plan.info.n_launches)
            torch.cuda.current_stream(dev).synchronize()
            return y_host
        # batch-chunked pipeline: the upload of chunk c+1 and the download of chunk c-1 overlap the kernels of chunk c
        # (PCIe is full duplex); every chunk is an independent convolution, so the result is unchanged
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = _side_streams(dev)
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        bounds = _chunk_bounds(B, n_chunks)
        n_chunks = len(bounds) - 1
        ev_in = []
        with torch.cuda.stream(s_in):
            for c in range(n_chunks):
                a0, a1 = bounds[c], bounds[c + 1]
                x_stage[a0:a1].copy_(x_host[a0:a1], non_blocking=True)
                e = torch.cuda.Event()
                e.record(s_in)
                ev_in.append(e)
        for c in range(n_chunks):
            a0, a1 = bounds[c], bounds[c + 1]
            # a chunk runs on the kernel spectrum of the full-batch plan, so its plan must choose the same spectrum layout:
            # the tensor-core layout depends on the batch (<= 32), so a full batch on the SIMT path keeps its chunks there
            cflags = flags | (0 if int(plan.info.tensor_core) else L.FC_FLAG_NO_TC)
            if signal.dim() == 3 and not transposed:  # 1-d batch segments: the (batch-dependent) choice of the full-batch plan
                cflags |= L.FC_FLAG_SEGMENT if int(plan.info.segments) > 1 else L.FC_FLAG_NO_SEGMENT
            sub = entry if a1 - a0 == B else get_plan(transposed, a1 - a0, cin, cout, groups, tuple(int(v) for v in signal.shape[2:]),
                                                      tuple(int(v) for v in kernel.shape[2:]), stride_, padding_, dilation_, opad_, padding_mode, cflags)
            sp = sub.plan
            if _kspec_layout(sp) != _kspec_layout(plan):
                raise RuntimeError("internal: the plan of a batch chunk does not share the kernel-spectrum layout of the full-batch plan")
            cur.wait_event(ev_in[c])
            L.check(lib, lib.fc_conv(sp.handle, _ptr(sub.const_for(dev)), ctypes.c_void_p(x_stage[a0:a1].data_ptr()), _ptr(kspec), _ptr(b_dev),
                                     ctypes.c_void_p(y_stage[a0:a1].data_ptr()), _ptr(ws), stream), "fc_conv")
            _launch_counter += int(sp.info.n_launches)
            e = torch.cuda.Event()
            e.record(cur)
            with torch.cuda.stream(s_out):
                s_out.wait_event(e)
                y_host[a0:a1].copy_(y_stage[a0:a1], non_blocking=True)
        cur.wait_stream(s_out)
        s_out.synchronize()
        for t in (x_stage, y_stage, ws, kspec):
            t.record_stream(s_in)
            t.record_stream(s_out)
        return y_host


_HOST_PIPELINE_CHUNKS = 6  # batch chunks of the host-buffer pipeline (see _chunk_bounds)


def _kspec_layout(plan) -> tuple:
    """What two plans must agree on to share one kernel spectrum."""
    i = plan.info
    return (plan.fft_size, int(i.segments), int(i.kspec_bytes), int(i.tensor_core), int(i.fused))


def _chunk_bounds(B: int, n_chunks: int):
    """Batch ranges of the host pipeline: an even split into n_chunks - 1 parts whose last part is halved again, so
    that the tail nothing can overlap with (kernels + download of the last chunk) is short. B = 8, 6 chunks: 1,2,1,2,1,1
    (BASELINE c2 through bench.py: 1.72 ms per call; 1.75 with 5 chunks, 1.84 with 4, 1.87 with 8, 2.32 unchunked)."""
    if n_chunks < 3 or B < n_chunks:
        n = min(B, n_chunks)
        return [(B * c) // n for c in range(n + 1)]
    n = n_chunks - 1
    b = [(B * c) // n for c in range(n + 1)]
    mid = (b[-2] + b[-1]) // 2
    return b[:-1] + ([mid] if b[-2] < mid < b[-1] else []) + [b[-1]]
_side = {}


def _side_streams(dev: torch.device):
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    s = _side.get(idx)
    if s is None:
        s = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        _side[idx] = s
    return s


def fft_conv(
    signal: Tensor,
    kernel: Tensor,
    bias: Tensor = None,
    stride: IntOrSeq = 1,
    padding: IntOrSeq = 0,
    dilation: IntOrSeq = 1,
    groups: int = 1,
    padding_mode: str = "constant",
) -> Tensor:
    """N-d (1, 2 or 3 spatial dims) convolution through the frequency domain; equals ``F.conv{n}d``.

    Same signature as the reference's ``fft_conv`` (reference functional.py:19-28).
    signal (B, Cin, *L), kernel (Cout, Cin/groups, *K), bias (Cout,) or None -> (B, Cout, *Lout).
    """
    return _run(False, signal, kernel, bias, stride, padding, 0, dilation, groups, padding_mode)


def fft_conv_transpose(
    signal: Tensor,
    kernel: Tensor,
    bias: Tensor = None,
    stride: IntOrSeq = 1,
    padding: IntOrSeq = 0,
    output_padding: IntOrSeq = 0,
    dilation: IntOrSeq = 1,
    groups: int = 1,
) -> Tensor:
    """Transposed convolution through the frequency domain; equals ``F.conv_transpose{n}d``.

    Same signature as the reference's ``fft_conv_transpose`` (reference functional.py:92-101).
    signal (B, Cin, *L), kernel (Cin, Cout/groups, *K), bias (Cout,) or None -> (B, Cout, *Lout).
    """
    return _run(True, signal, kernel, bias, stride, padding, output_padding, dilation, groups, "constant")


def complex_matmul(a: Tensor, b: Tensor, groups: int = 1) -> Tensor:
    """Grouped per-bin channel contraction ``einsum("bgi...,goi...->bgo...")`` (reference functional.py:11-16).

    a: (B, Cin, *bins) complex64, b: (Cout, Cin/groups, *bins) complex64 -> (B, Cout, *bins) complex64.
    """
    global _launch_counter
    if a.dtype != torch.complex64 or b.dtype != torch.complex64:
        raise TypeError("complex_matmul expects complex64 tensors")
    _require_cuda()
    if not (a.is_cuda and b.is_cuda):
        raise ValueError("complex_matmul expects CUDA tensors")
    B, cin = int(a.shape[0]), int(a.shape[1])
    cout = int(b.shape[0])
    if cin % groups or cout % groups or b.shape[1] != cin // groups or tuple(a.shape[2:]) != tuple(b.shape[2:]):
        raise ValueError(f"complex_matmul: incompatible shapes {tuple(a.shape)} x {tuple(b.shape)} with groups={groups}")
    bins = 1
    for s in a.shape[2:]:
        bins *= int(s)
    ar = torch.view_as_real(a.contiguous())
    br = torch.view_as_real(b.contiguous())
    y = torch.empty((B, cout) + tuple(a.shape[2:]), dtype=torch.complex64, device=a.device)
    yr = torch.view_as_real(y)
    lib = L.load()
    with torch.cuda.device(a.device):
        stream = ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)
        L.check(lib, lib.fc_complex_matmul(_ptr(ar), _ptr(br), _ptr(yr), B, cin, cout, groups, bins, stream), "fc_complex_matmul")
    _launch_counter += 1
    return y
