"""The adjoint formulas of fft_conv_pytorch_b200/autograd.py, checked on CPU with the oracle standing in for
the device op (the `-m gpu` suite repeats this with the real kernels). Mirrors the reference's backward tests
(reference tests/test_functional.py:72-117, tests/test_functional_transpose.py:73-124): weight and bias gradients
of y.sum() against torch's direct convolution; the input gradient is checked as well."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from fft_conv_pytorch_b200 import autograd as ag
from oracle import fftconv_oracle as O


def _oracle_raw(transposed, x, w, b, stride, padding, opad, dilation, groups):
    xn, wn = x.detach().numpy().astype(np.float64), w.detach().numpy().astype(np.float64)
    bn = None if b is None else b.detach().numpy().astype(np.float64)
    if transposed:
        y = O.fft_conv_transpose(xn, wn, bn, stride=stride, padding=padding, output_padding=opad, dilation=dilation, groups=groups)
    else:
        y = O.fft_conv(xn, wn, bn, stride=stride, padding=padding, dilation=dilation, groups=groups)
    return torch.from_numpy(np.ascontiguousarray(y)).to(x.dtype)


@pytest.fixture(autouse=True)
def _patch(monkeypatch):
    monkeypatch.setattr(ag, "_raw_conv", _oracle_raw)


def _gcd(a, b):
    while b:
        a, b = b, a % b
    return a


@pytest.mark.parametrize("ndim", [1, 2, 3])
@pytest.mark.parametrize("size", [7, 8])
@pytest.mark.parametrize("cin,cout,groups", [(2, 2, 1), (3, 3, 3), (2, 4, 2), (3, 2, 1)])
@pytest.mark.parametrize("k,p,s,d", [(2, 0, 1, 1), (3, 1, 2, 2), (3, 1, 1, 2), (2, 1, 2, 1), (3, 0, 3, 1)])
def test_forward_conv_grads(ndim, size, cin, cout, groups, k, p, s, d):
    torch.manual_seed(0)
    x0 = torch.randn(2, cin, *([size] * ndim), dtype=torch.float64, requires_grad=True)
    w0 = torch.randn(cout, cin // groups, *([k] * ndim), dtype=torch.float64, requires_grad=True)
    b0 = torch.randn(cout, dtype=torch.float64, requires_grad=True)
    x1, w1, b1 = (t.detach().clone().requires_grad_() for t in (x0, w0, b0))
    nt = lambda v: (v,) * ndim
    y0 = ag.conv_with_grad(False, x0, w0, b0, nt(s), nt(p), nt(0), nt(d), groups, "constant")
    y1 = getattr(F, f"conv{ndim}d")(x1, w1, b1, stride=s, padding=p, dilation=d, groups=groups)
    g = torch.randn_like(y1)
    (y0 * g).sum().backward()
    (y1 * g).sum().backward()
    assert torch.allclose(y0, y1, atol=1e-9)
    assert torch.allclose(w0.grad, w1.grad, atol=1e-9)
    assert torch.allclose(b0.grad, b1.grad, atol=1e-9)
    assert torch.allclose(x0.grad, x1.grad, atol=1e-9)


@pytest.mark.parametrize("ndim", [1, 2])
@pytest.mark.parametrize("size", [7, 8])
@pytest.mark.parametrize("cin,cout,groups", [(2, 2, 1), (3, 3, 3), (2, 4, 2), (2, 3, 1)])
@pytest.mark.parametrize("k,p,op,s,d", [(2, 0, 0, 1, 1), (3, 1, 1, 3, 3), (3, 1, 0, 2, 2), (2, 1, 2, 4, 3), (3, 0, 1, 2, 1)])
def test_transposed_conv_grads(ndim, size, cin, cout, groups, k, p, op, s, d):
    torch.manual_seed(1)
    x0 = torch.randn(2, cin, *([size] * ndim), dtype=torch.float64, requires_grad=True)
    w0 = torch.randn(cin, cout // groups, *([k] * ndim), dtype=torch.float64, requires_grad=True)
    b0 = torch.randn(cout, dtype=torch.float64, requires_grad=True)
    x1, w1, b1 = (t.detach().clone().requires_grad_() for t in (x0, w0, b0))
    nt = lambda v: (v,) * ndim
    y0 = ag.conv_with_grad(True, x0, w0, b0, nt(s), nt(p), nt(op), nt(d), groups, "constant")
    y1 = getattr(F, f"conv_transpose{ndim}d")(x1, w1, b1, stride=s, padding=p, output_padding=op, dilation=d, groups=groups)
    g = torch.randn_like(y1)
    (y0 * g).sum().backward()
    (y1 * g).sum().backward()
    assert torch.allclose(y0, y1, atol=1e-9)
    assert torch.allclose(w0.grad, w1.grad, atol=1e-9)
    assert torch.allclose(b0.grad, b1.grad, atol=1e-9)
    assert torch.allclose(x0.grad, x1.grad, atol=1e-9)


@pytest.mark.parametrize("mode", ["reflect", "replicate", "circular"])
def test_padding_mode_grads(mode):
    torch.manual_seed(2)
    x0 = torch.randn(2, 2, 9, 8, dtype=torch.float64, requires_grad=True)
    w0 = torch.randn(3, 2, 3, 3, dtype=torch.float64, requires_grad=True)
    x1, w1 = (t.detach().clone().requires_grad_() for t in (x0, w0))
    y0 = ag.conv_with_grad(False, x0, w0, None, (1, 2), (2, 1), (0, 0), (1, 1), 1, mode)
    y1 = F.conv2d(F.pad(x1, [1, 1, 2, 2], mode=mode), w1, None, stride=(1, 2))
    y0.sum().backward()
    y1.sum().backward()
    assert torch.allclose(y0, y1, atol=1e-9)
    assert torch.allclose(w0.grad, w1.grad, atol=1e-9)
    assert torch.allclose(x0.grad, x1.grad, atol=1e-9)
