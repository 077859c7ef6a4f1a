"""Host-side contract of the drop-in: exported symbols, argument normalisation, module layout, error behaviour.
No compute is launched here (there is no GPU in the `-m "not gpu"` run)."""
import ctypes
import os
import re

import pytest
import torch

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "fftconv_b200.h")).read()
    declared = set(re.findall(r"\b(fc_[a-z_0-9]+)\s*\(", header))
    declared -= {"fc_problem", "fc_plan", "fc_plan_info"}
    assert declared == set(L.SYMBOLS), (declared ^ set(L.SYMBOLS))
    lib = L.load()  # raises if the library is missing or a symbol is not exported
    assert b"sm_100a" in lib.fc_version()


def test_struct_layout_matches_header():
    assert ctypes.sizeof(L.FcProblem) == 4 * (6 + 6 * 3 + 4)
    assert ctypes.sizeof(L.FcPlanInfo) == 48 + 8 * 11 + 8 + 8


def test_plan_creation_needs_no_gpu():
    prob = L.make_problem(False, 8, 8, 8, 1, (512, 512), (65, 65), (1, 1), (0, 0), (1, 1))
    plan = L.Plan(L.load(), prob)
    assert plan.out_size == (448, 448)
    assert "R2C" in plan.describe()


def test_package_exports():
    for name in ("fft_conv", "fft_conv_transpose", "complex_matmul", "FFTConv1d", "FFTConv2d", "FFTConv3d",
                 "FFTConvTranspose1d", "FFTConvTranspose2d", "FFTConvTranspose3d", "functional", "nn", "to_ntuple"):
        assert hasattr(fcp, name)


def test_to_ntuple():
    assert fcp.to_ntuple(2, 3) == (2, 2, 2)
    assert fcp.to_ntuple((1, 2), 2) == (1, 2)
    with pytest.raises(ValueError, match="Cannot cast tuple of length 3 to length 2"):
        fcp.to_ntuple((1, 2, 3), 2)
    with pytest.raises(ValueError):
        fcp.to_ntuple("same", 2)  # like the reference: string paddings are rejected


@pytest.mark.parametrize("cls,base", [
    (fcp.FFTConv1d, torch.nn.Conv1d), (fcp.FFTConv2d, torch.nn.Conv2d), (fcp.FFTConv3d, torch.nn.Conv3d),
    (fcp.FFTConvTranspose1d, torch.nn.ConvTranspose1d), (fcp.FFTConvTranspose2d, torch.nn.ConvTranspose2d),
    (fcp.FFTConvTranspose3d, torch.nn.ConvTranspose3d)])
def test_modules_share_torch_layout(cls, base):
    m = cls(4, 6, 3, stride=2, padding=1, groups=2)
    t = base(4, 6, 3, stride=2, padding=1, groups=2)
    assert isinstance(m, base)
    assert list(m.state_dict().keys()) == ["weight", "bias"]
    assert m.weight.shape == t.weight.shape
    m.load_state_dict(t.state_dict())
    t.load_state_dict(m.state_dict())
    assert torch.equal(m.weight, t.weight)


def test_module_asserts_batched_input():
    m = fcp.FFTConv1d(2, 2, 3)
    with pytest.raises(AssertionError):
        m(torch.randn(2, 8))  # reference nn.py:11


def test_argument_errors_before_any_launch():
    x = torch.randn(1, 2, 8)
    w = torch.randn(2, 2, 3)
    with pytest.raises(ValueError, match="Cannot cast tuple"):
        fcp.fft_conv(x, w, padding=(1, 1))
    with pytest.raises(TypeError, match="float32"):
        fcp.fft_conv(x.double(), w.double())
    with pytest.raises(TypeError, match="float32"):
        fcp.fft_conv(x.half(), w.half())
    with pytest.raises(ValueError, match="kernel must have shape"):
        fcp.fft_conv(x, torch.randn(2, 3, 3))
    with pytest.raises(ValueError, match="bias must have shape"):
        fcp.fft_conv(x, w, torch.randn(3))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    x = torch.randn(1, 2, 8)
    w = torch.randn(2, 2, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            fcp.fft_conv(x, w)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        L.load(str(tmp_path / "libfftconv_b200.so"))
