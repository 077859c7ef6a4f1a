"""CPU oracle for the fft_conv hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the algorithm of klae01/fft-conv-pytorch's convolution path, written from the
behavioural spec with each step citing the reference lines it follows. Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import this module; the
product package (``fft_conv_pytorch_b200``) never does and fails loudly without its CUDA library.

Where the arithmetic lives: the reference delegates every FLOP to a third-party dependency that is not under
/root/reference — PyTorch (``torch>=1.8``, reference setup.py:33; 2.11.0+cu128 in this image): ``torch.fft.rfftn`` /
``irfftn`` (functional.py:68-75, 155-162), ``torch.einsum`` (functional.py:12) and ``F.pad`` (functional.py:62).
Their published semantics are the unnormalised forward DFT over the last ``n`` axes with a one-sided last axis,
its 1/prod(N)-normalised inverse, an Einstein-summation contraction, and numpy-compatible pad modes; they are
restated here with ``numpy.fft`` (pocketfft) / ``numpy.einsum`` / ``numpy.pad``.

Parity pinning: ``tests/golden/make_golden.py`` imported the UNMODIFIED reference from /root/reference in the
authoring container and stored its outputs (and ``torch.nn.functional.conv*`` outputs) for a grid of cases;
``tests/test_oracle.py`` checks this module against every stored vector. Parity is therefore pinned.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence, Tuple, Union

import numpy as np

try:  # multi-threaded pocketfft when scipy is present (same algorithm as numpy.fft)
    import scipy.fft as _fft

    def _rfftn(a, s, axes, workers=None):
        return _fft.rfftn(a, s=s, axes=axes, workers=workers)

    def _irfftn(a, s, axes, workers=None):
        return _fft.irfftn(a, s=s, axes=axes, workers=workers)

except Exception:  # pragma: no cover
    def _rfftn(a, s, axes, workers=None):
        return np.fft.rfftn(a, s=s, axes=axes)

    def _irfftn(a, s, axes, workers=None):
        return np.fft.irfftn(a, s=s, axes=axes)


IntOrSeq = Union[int, Iterable[int]]

_NP_PAD_MODE = {"constant": "constant", "zeros": "constant", "reflect": "reflect", "replicate": "edge", "circular": "wrap"}


def to_ntuple(val: IntOrSeq, n: int) -> Tuple[int, ...]:
    """reference utils.py:4-20 — an int is repeated n times, an iterable must already have length n."""
    if isinstance(val, Iterable):
        out = tuple(val)
        if len(out) != n:
            raise ValueError(f"Cannot cast tuple of length {len(out)} to length {n}.")
        return out
    return n * (val,)


def complex_matmul(a: np.ndarray, b: np.ndarray, groups: int = 1) -> np.ndarray:
    """reference functional.py:11-16 — y[b, (g,o), f] = sum_i a[b, (g,i), f] * b[(g,o), i, f]."""
    B, cin = a.shape[:2]
    cout = b.shape[0]
    ag = a.reshape(B, groups, cin // groups, *a.shape[2:])
    bg = b.reshape(groups, cout // groups, *b.shape[1:])
    y = np.einsum("bgi...,goi...->bgo...", ag, bg)
    return y.reshape(B, cout, *a.shape[2:])


def _dilate(kernel: np.ndarray, dilation: Sequence[int]) -> np.ndarray:
    """reference functional.py:49-57 / 115-124 — scatter the taps onto a zero lattice of pitch `dilation`."""
    if all(d == 1 for d in dilation):
        return kernel
    shape = list(kernel.shape[:2]) + [(k - 1) * d + 1 for k, d in zip(kernel.shape[2:], dilation)]
    out = np.zeros(shape, dtype=kernel.dtype)
    out[(slice(None), slice(None)) + tuple(slice(None, None, d) for d in dilation)] = kernel
    return out


def fft_conv(
    signal: np.ndarray,
    kernel: np.ndarray,
    bias: Optional[np.ndarray] = None,
    stride: IntOrSeq = 1,
    padding: IntOrSeq = 0,
    dilation: IntOrSeq = 1,
    groups: int = 1,
    padding_mode: str = "constant",
    workers: Optional[int] = None,
) -> np.ndarray:
    """reference functional.py:19-89."""
    n = signal.ndim - 2
    padding_ = to_ntuple(padding, n)  # functional.py:44-47
    stride_ = to_ntuple(stride, n)
    dilation_ = to_ntuple(dilation, n)
    kernel = _dilate(kernel, dilation_)  # functional.py:49-57
    if any(p != 0 for p in padding_):  # functional.py:60-62
        pads = [(0, 0), (0, 0)] + [(p, p) for p in padding_]
        signal = np.pad(signal, pads, mode=_NP_PAD_MODE[padding_mode])
    axes = tuple(range(-n, 0))
    interm = [(s + 1) // 2 * 2 for s in signal.shape[2:]]  # functional.py:66 — every axis rounded up to even
    sig_f = _rfftn(signal, interm, axes, workers)  # functional.py:70
    ker_f = np.conj(_rfftn(kernel, interm, axes, workers))  # functional.py:71
    out = _irfftn(complex_matmul(sig_f, ker_f, groups), interm, axes, workers)  # functional.py:68-75
    crop = (slice(None), slice(None)) + tuple(  # functional.py:76-82
        slice(0, signal.shape[i] - kernel.shape[i] + 1, stride_[i - 2]) for i in range(2, signal.ndim)
    )
    out = np.ascontiguousarray(out[crop]).astype(signal.dtype, copy=False)
    if bias is not None:  # functional.py:85-87
        out = out + bias.reshape((1, -1) + (1,) * n).astype(out.dtype)
    return out


def fft_conv_transpose(
    signal: np.ndarray,
    kernel: np.ndarray,
    bias: Optional[np.ndarray] = None,
    stride: IntOrSeq = 1,
    padding: IntOrSeq = 0,
    output_padding: IntOrSeq = 0,
    dilation: IntOrSeq = 1,
    groups: int = 1,
    workers: Optional[int] = None,
) -> np.ndarray:
    """reference functional.py:92-176.

    One deliberate difference: for kernel extent 1 with output_padding > padding the reference's crop runs off
    its buffer and returns a short result; here the transform extent is raised so that the result
    has torch's shape.
    """
    n = signal.ndim - 2
    padding_ = to_ntuple(padding, n)  # functional.py:103-107
    opad_ = to_ntuple(output_padding, n)
    stride_ = to_ntuple(stride, n)
    dilation_ = to_ntuple(dilation, n)
    cin = kernel.shape[0]
    # functional.py:109-114 — flip spatially, (Cin, Cout/g, k) -> (g, Cin/g, Cout/g, k) -> swap -> (Cout, Cin/g, k)
    k = kernel[(slice(None), slice(None)) + (slice(None, None, -1),) * n]
    k = k.reshape(groups, cin // groups, *k.shape[1:])
    k = np.swapaxes(k, 1, 2)
    k = np.ascontiguousarray(k).reshape(-1, cin // groups, *kernel.shape[2:])
    k_ = _dilate(k, dilation_)  # functional.py:115-124
    # functional.py:126-139 — zero-stuff by `stride`, left-pad by Kd-1
    stuffed_shape = list(signal.shape[:2]) + [
        (s - 1) * t + 1 + (kk - 1) for s, kk, t in zip(signal.shape[2:], k_.shape[2:], stride_)
    ]
    sig_ = np.zeros(stuffed_shape, dtype=signal.dtype)
    sig_[(slice(None), slice(None)) + tuple(slice(kk - 1, None, t) for kk, t in zip(k_.shape[2:], stride_))] = signal
    out_shape = [  # functional.py:144-154
        (s - 1) * t - 2 * p + d * (kk - 1) + o + 1
        for s, kk, t, p, d, o in zip(signal.shape[2:], kernel.shape[2:], stride_, padding_, dilation_, opad_)
    ]
    interm = [(s + kk) // 2 * 2 for s, kk in zip(sig_.shape[2:], k_.shape[2:])]  # functional.py:143
    interm = [max(m, ((o + p) + 1) // 2 * 2) for m, o, p in zip(interm, out_shape, padding_)]  # A.5 fix, see docstring
    axes = tuple(range(-n, 0))
    sig_f = _rfftn(sig_, interm, axes, workers)  # functional.py:157
    ker_f = np.conj(_rfftn(k_, interm, axes, workers))  # functional.py:158
    out = _irfftn(complex_matmul(sig_f, ker_f, groups), interm, axes, workers)  # functional.py:155-162
    crop = (slice(None), slice(None)) + tuple(slice(p, o + p) for o, p in zip(out_shape, padding_))  # functional.py:163-169
    out = np.ascontiguousarray(out[crop]).astype(signal.dtype, copy=False)
    if bias is not None:  # functional.py:172-174
        out = out + bias.reshape((1, -1) + (1,) * n).astype(out.dtype)
    return out
