"""Backward pass of the FFT convolution, built from the same forward kernels.

The reference has no custom autograd: torch differentiates through rfftn / einsum / irfftn
(reference tests pin ``weight.grad`` and ``bias.grad``: tests/test_functional.py:72-117,
tests/test_functional_transpose.py:73-124). Here the op is opaque to autograd, so the adjoints are written out;
every one of them is again a (transposed) convolution and runs on the same sm_100a kernels:

  y = conv(x, w)                      grad_x = conv_transpose(grad_y, w)        grad_w = corr(x_pad, grad_y) sampled on the dilation lattice
  y = conv_transpose(x, w)            grad_x = conv(grad_y, w)                  grad_w = corr(grad_y_pad, x) sampled on the dilation lattice
  grad_bias = sum of grad_y over batch and space

The weight gradients are evaluated as convolutions with batch and channel roles swapped (stride <-> dilation).
Non-zero padding modes are differentiated by padding explicitly with ``F.pad`` first.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor


def _raw_conv(transposed: bool, x: Tensor, w: Tensor, b: Optional[Tensor], stride, padding, opad, dilation, groups: int) -> Tensor:
    """The non-differentiable op (zero padding mode). Tests substitute an oracle-backed implementation."""
    from . import functional as Fn

    with torch.no_grad():
        return Fn._run(transposed, x, w, b, stride, padding, opad, dilation, groups, "constant")


def _crop(t: Tensor, sizes: Tuple[int, ...]) -> Tensor:
    idx = (slice(None), slice(None)) + tuple(slice(0, s) for s in sizes)
    return t[idx]


def _to_group_major(t: Tensor, groups: int, lead_is_channel: bool) -> Tensor:
    """(B, G*c, *sp) -> (c, G*B, *sp) when lead_is_channel (a "signal" whose batch is the channel index and whose channels
    are (group, batch)), else -> (G*c, B, *sp) (a "kernel" with B input channels per group). One contiguous copy."""
    B = t.shape[0]
    c = t.shape[1] // groups
    sp = tuple(t.shape[2:])
    v = t.reshape(B, groups, c, *sp)
    tail = tuple(range(3, 3 + len(sp)))
    if lead_is_channel:
        return v.permute(2, 1, 0, *tail).reshape(c, groups * B, *sp)
    return v.permute(1, 2, 0, *tail).reshape(groups * c, B, *sp)


def _grad_weight_fwd(x: Tensor, gy: Tensor, ksize, stride, padding, dilation, groups: int) -> Tensor:
    """grad_w[o, i, m] = sum_{b, j} gy[b, o, j] * xpad[b, i, j*s + m*d]  -> (Cout, Cin/g, *K).

    One grouped convolution for all groups: batch and channel roles swapped (signal (Cin/g, G*B, *L), kernel
    (G*Cout/g, B, *Lout), groups = G, stride <-> dilation)."""
    cin, cout = x.shape[1], gy.shape[1]
    ig, og = cin // groups, cout // groups
    xs = _to_group_major(x, groups, True)    # (ig, G*B, *L)
    gs = _to_group_major(gy, groups, False)  # (G*og, B, *Lout)
    r = _crop(_raw_conv(False, xs, gs, None, dilation, padding, 0, stride, groups), ksize)  # (ig, G*og, *K)
    tail = tuple(range(3, 3 + len(ksize)))
    return r.reshape(ig, groups, og, *ksize).permute(1, 2, 0, *tail).reshape(cout, ig, *ksize)


def _grad_weight_tr(x: Tensor, gy: Tensor, ksize, stride, padding, dilation, groups: int) -> Tensor:
    """grad_w[i, o, m] = sum_{b, q} x[b, i, q] * gypad[b, o, q*t + m*d]  -> (Cin, Cout/g, *K); one grouped convolution."""
    cin, cout = x.shape[1], gy.shape[1]
    ig, og = cin // groups, cout // groups
    gs = _to_group_major(gy, groups, True)  # (og, G*B, *Lout)  signal
    xs = _to_group_major(x, groups, False)  # (G*ig, B, *L)      kernel
    r = _crop(_raw_conv(False, gs, xs, None, dilation, padding, 0, stride, groups), ksize)  # (og, G*ig, *K)
    tail = tuple(range(3, 3 + len(ksize)))
    return r.reshape(og, groups, ig, *ksize).permute(1, 2, 0, *tail).reshape(cin, og, *ksize)


class _FFTConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, transposed, stride, padding, opad, dilation, groups):
        ctx.save_for_backward(x, w)
        ctx.cfg = (transposed, stride, padding, opad, dilation, groups, b is not None)
        return _raw_conv(transposed, x, w, b, stride, padding, opad, dilation, groups)

    @staticmethod
    @torch.autograd.function.once_differentiable  # the adjoints run on the opaque kernels: no double backward
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        transposed, stride, padding, opad, dilation, groups, has_bias = ctx.cfg
        gy = gy.contiguous()
        n = x.ndim - 2
        ksize = tuple(w.shape[2:])
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if not transposed:
                # adjoint of a strided conv is a transposed conv; output_padding restores the rows the stride dropped
                op = tuple(
                    x.shape[2 + i] - ((gy.shape[2 + i] - 1) * stride[i] - 2 * padding[i] + dilation[i] * (ksize[i] - 1) + 1)
                    for i in range(n)
                )
                gx = _raw_conv(True, gy, w, None, stride, padding, op, dilation, groups)
            else:
                gx = _crop(_raw_conv(False, gy, w, None, stride, padding, 0, dilation, groups), tuple(x.shape[2:]))
        if ctx.needs_input_grad[1]:
            if not transposed:
                gw = _grad_weight_fwd(x, gy, ksize, stride, padding, dilation, groups)
            else:
                gw = _grad_weight_tr(x, gy, ksize, stride, padding, dilation, groups)
        if has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(dim=(0,) + tuple(range(2, gy.ndim)))
        return gx, gw, gb, None, None, None, None, None, None


def conv_with_grad(transposed: bool, signal: Tensor, kernel: Tensor, bias: Optional[Tensor], stride, padding, opad, dilation,
                   groups: int, padding_mode: str) -> Tensor:
    if not transposed and padding_mode != "constant" and any(p != 0 for p in padding):
        pads = [p for p in padding[::-1] for _ in range(2)]
        signal = F.pad(signal, pads, mode=padding_mode)  # differentiable; the conv itself then runs unpadded
        padding = (0,) * len(padding)
    return _FFTConvFn.apply(signal, kernel, bias, transposed, tuple(stride), tuple(padding), tuple(opad), tuple(dilation), groups)
