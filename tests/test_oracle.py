"""The oracle (oracle/) pinned against outputs of the unmodified reference (tests/golden/ref_cases.npz)."""
import numpy as np
import pytest

from tests.helpers import golden, rel_err
from oracle import direct, fftconv_oracle as O

N_CASES = len(golden())


@pytest.mark.parametrize("i", range(N_CASES))
def test_oracle_matches_reference_outputs(i):
    c = golden().case(i)
    fn = O.fft_conv_transpose if c["transposed"] else O.fft_conv
    y32 = fn(c["x"], c["w"], c["b"], **c["kw"])
    y64 = fn(c["x"].astype(np.float64), c["w"].astype(np.float64), None if c["b"] is None else c["b"].astype(np.float64), **c["kw"])
    assert y64.shape == c["direct"].shape
    # fp64 restatement == direct convolution (the quantity the reference's tests pin), to fp32 storage precision
    assert rel_err(y64, c["direct"]) < 1e-6
    if not c["ref_short"]:  # reference bug A.5 returns a short tensor for K=1, output_padding>padding
        assert y32.shape == c["ref32"].shape
        assert rel_err(y32, c["ref32"]) < 2e-5  # two fp32 FFT pipelines (pocketfft vs MKL)
        # reference's own tolerance for small cases (benchmark_utils.py:53-57)
        if c["x"].size < 5000:
            d = np.abs(y32 - c["ref32"])
            assert d.mean() < 5e-5 and d.max() < 1e-4


@pytest.mark.parametrize("i", [i for i in range(N_CASES) if golden().npz[f"dir_{i}"].size * golden().npz[f"w{i}"].size < 3e7])
def test_direct_c_oracle_matches_reference_outputs(i):
    c = golden().case(i)
    if c["transposed"]:
        y = direct.direct_conv_transpose(c["x"], c["w"], c["b"], **c["kw"])
    else:
        y = direct.direct_conv(c["x"], c["w"], c["b"], **c["kw"])
    assert y.shape == c["direct"].shape
    assert rel_err(y, c["direct"]) < 1e-6


def test_to_ntuple_contract():
    assert O.to_ntuple(3, 2) == (3, 3)
    assert O.to_ntuple([1, 2], 2) == (1, 2)
    with pytest.raises(ValueError):
        O.to_ntuple((1, 2, 3), 2)
    with pytest.raises(ValueError):
        O.to_ntuple("same", 2)


def test_complex_matmul_groups():
    rng = np.random.RandomState(0)
    a = (rng.standard_normal((2, 6, 5)) + 1j * rng.standard_normal((2, 6, 5))).astype(np.complex64)
    b = (rng.standard_normal((4, 3, 5)) + 1j * rng.standard_normal((4, 3, 5))).astype(np.complex64)
    y = O.complex_matmul(a, b, groups=2)
    ref = np.zeros((2, 4, 5), np.complex64)
    for bb in range(2):
        for o in range(4):
            g = o // 2
            for i in range(3):
                ref[bb, o] += a[bb, g * 3 + i] * b[o, i]
    assert np.abs(y - ref).max() < 1e-5
