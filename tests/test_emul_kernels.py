"""The generic CUDA kernels (fc_kernels.cuh), executed on host threads by tests/cpu_emul, against the golden
outputs of the reference. This checks the kernels' index algebra — pass program, gather/scatter maps, layouts,
four-step twiddles, Stockham stages — in a container without a GPU. The `-m gpu` tests repeat the same cases on
the real device through the product library."""
import numpy as np
import pytest

from tests.cpu_emul import emul
from tests.helpers import golden, rel_err
from fft_conv_pytorch_b200 import _lib as L

N_CASES = len(golden())
# every 3rd small case + all the medium / special ones at the end of the file
SEL = sorted(set(range(0, N_CASES, 3)) | set(range(N_CASES - 24, N_CASES)))


@pytest.mark.parametrize("i", SEL)
def test_emulated_kernels_match_reference(i):
    c = golden().case(i)
    y, plan = emul.conv(c["x"], c["w"], c["b"], transposed=c["transposed"], **c["kw"])
    assert y.shape == c["direct"].shape
    assert not np.isnan(y).any()  # every output element was written
    assert rel_err(y, c["direct"]) < 1e-5
    if not c["ref_short"]:
        assert rel_err(y, c["ref32"]) < 2e-5


@pytest.mark.parametrize("i", SEL[::7])
def test_stage_calls_equal_fused_call(i):
    c = golden().case(i)
    y0, p0 = emul.conv(c["x"], c["w"], c["b"], transposed=c["transposed"], **c["kw"])
    y1, _ = emul.conv(c["x"], c["w"], c["b"], transposed=c["transposed"], staged=True, **c["kw"])
    d = p0.describe()
    if any(t in d for t in ("fast_", "fused_", "col_", "tc_")):  # fc_conv ran specialised kernels, the stages the generic ones
        assert rel_err(y0, y1) < 2e-6
    else:
        assert np.array_equal(y0, y1)


@pytest.mark.parametrize("i", SEL[::5])
def test_polyphase_off_gives_same_result(i):
    c = golden().case(i)
    y0, p0 = emul.conv(c["x"], c["w"], c["b"], transposed=c["transposed"], **c["kw"])
    y1, p1 = emul.conv(c["x"], c["w"], c["b"], transposed=c["transposed"], flags=L.FC_FLAG_NO_POLYPHASE, **c["kw"])
    assert rel_err(y0, y1) < 1e-5


@pytest.mark.parametrize("threads", [32, 96, 256])
def test_thread_count_independent(threads):
    c = golden().case(len(golden()) - 3)  # 2-d medium case
    y, _ = emul.conv(c["x"], c["w"], c["b"], transposed=c["transposed"], threads=threads, **c["kw"])
    assert rel_err(y, c["direct"]) < 1e-5


def test_linearity_and_shift_property():
    """Size-independent properties used at BASELINE sizes on the GPU: linearity in the signal and
    conv(delta kernel) == shifted crop of the signal."""
    rng = np.random.RandomState(5)
    x1 = rng.standard_normal((1, 2, 40, 24)).astype(np.float32)
    x2 = rng.standard_normal((1, 2, 40, 24)).astype(np.float32)
    w = rng.standard_normal((2, 2, 7, 5)).astype(np.float32)
    ya, _ = emul.conv(x1, w)
    yb, _ = emul.conv(x2, w)
    yc, _ = emul.conv(x1 + 2 * x2, w)
    assert rel_err(yc, ya + 2 * yb) < 1e-5
    d = np.zeros((2, 2, 7, 5), np.float32)
    d[0, 0, 3, 2] = 1.0
    d[1, 1, 0, 4] = 1.0
    yd, _ = emul.conv(x1, d)
    assert rel_err(yd[0, 0], x1[0, 0, 3:3 + 34, 2:2 + 20]) < 1e-5
    assert rel_err(yd[0, 1], x1[0, 1, 0:34, 4:24]) < 1e-5


def test_bad_problems_raise():
    with pytest.raises(ValueError, match="exceeds the padded signal"):
        emul.plan_for((1, 1, 4), (1, 1, 7))
    with pytest.raises(ValueError, match="divisible by groups"):
        emul.plan_for((1, 3, 8), (2, 1, 3), groups=2)
    with pytest.raises(ValueError, match="reflect"):
        emul.plan_for((1, 1, 4), (1, 1, 3), padding=4, padding_mode="reflect")
    with pytest.raises(ValueError):
        emul.plan_for((1, 1, 8), (1, 1, 3), stride=0)


def test_plan_shapes_follow_reference_formulas():
    # forward: Lout = (L + 2p - d(K-1) - 1)//s + 1 (reference functional.py:79); transposed: functional.py:144-154
    p = emul.plan_for((8, 8, 512, 512), (8, 8, 65, 65))
    assert p.out_size == (448, 448) and p.fft_size == (512, 512)
    assert p.info.bins == 512 * 257
    p = emul.plan_for((4, 64, 1024, 1024), (64, 16, 31, 31), transposed=True, stride=2, dilation=2, groups=4)
    assert p.out_size == (2107, 2107)
    # polyphase: dense 1024 (*) 31 -> 1054 per axis, run as 5 x 5 overlap-save segments of 256 (f3)
    assert p.fft_size == (256, 256) and p.info.segments == 25
    p = emul.plan_for((4, 64, 1024, 1024), (64, 16, 31, 31), transposed=True, stride=2, dilation=2, groups=4, flags=L.FC_FLAG_NO_SEGMENT)
    assert p.fft_size == (2048, 2048) and p.info.segments == 1
    p = emul.plan_for((16, 256, 65536), (256, 256, 4097))  # 5 windows of 16384 points as batch items (f3, 1-d)
    assert p.out_size == (61440,) and p.fft_size == (16384,) and p.info.segments == 5
    p = emul.plan_for((16, 256, 65536), (256, 256, 4097), flags=L.FC_FLAG_NO_SEGMENT)
    assert p.out_size == (61440,) and p.fft_size == (65536,) and p.info.segments == 1
    p = emul.plan_for((1, 8, 32768), (8, 8, 1025))
    assert p.out_size == (31744,)
    # SURVEY §8d algorithmic bytes at the reference extents for c1/c2
    p = emul.plan_for((8, 8, 512, 512), (8, 8, 65, 65))
    i = p.info
    assert abs((i.algo_bytes_s1 + i.algo_bytes_s3 + i.algo_bytes_s4) / 1e6 - 455.3) < 0.5


# ------------------------------------------------------------------------------------ specialised kernels (fc_fused.cuh)
_FAST_SHAPES = [
    # x, w, kwargs -> transform sizes that select the warp-FFT kernels (last axis 512/1024, fused axis 256/512)
    ((2, 2, 150, 300), (3, 2, 7, 9), {}),
    ((1, 8, 256, 260), (8, 8, 3, 5), {}),  # full 8x8 channel groups, full lines: the predicate-free instantiation
    ((2, 4, 130, 270), (4, 2, 5, 3), dict(groups=2, padding=(3, 0), stride=(2, 1))),
    ((1, 2, 260, 600), (2, 2, 3, 3), dict(padding=(1, 1), padding_mode="reflect")),
    ((1, 2, 20, 1100), (2, 2, 3, 5), {}),  # last axis 2048: the one-line-per-warp, 16-warp variant
    ((1, 2, 18, 1030), (2, 1, 3, 4), dict(groups=2, stride=(1, 3))),  # ... with a strided scatter on store
    ((2, 2, 150, 300), (3, 2, 7, 9), dict(padding=(3, 4))),  # "same" convolution: even zero padding stays on the row kernel K1
]


def test_row_kernel_takes_even_zero_padding():
    p = emul.plan_for((2, 2, 150, 300), (3, 2, 7, 9), padding=(3, 4), threads=256)
    assert "fast_r2c" in p.describe()
    p = emul.plan_for((2, 2, 150, 300), (3, 2, 7, 9), padding=(3, 3), threads=256)  # odd: pairs would straddle the edge
    assert "fast_r2c" not in p.describe()


@pytest.mark.parametrize("xs,ws,kw", _FAST_SHAPES)
def test_fast_kernels_match_generic_and_oracle(xs, ws, kw):
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(11)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    ref = O.fft_conv(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), **kw)
    y_gen, p_gen = emul.conv(x, w, b, threads=256, flags=L.FC_FLAG_NO_FUSED, **kw)
    y_fast, p_fast = emul.conv(x, w, b, threads=256, **kw)
    assert p_gen.info.n_launches == 5 and "fast" not in p_gen.describe().split("launch", 1)[1]
    assert "fused_axis" in p_fast.describe() or "fast_" in p_fast.describe()
    assert rel_err(y_gen, ref) < 1e-5
    assert rel_err(y_fast, ref) < 1e-5
    assert not np.isnan(y_fast).any()


def test_fast_kernels_one_at_a_time():
    rng = np.random.RandomState(12)
    x = rng.standard_normal((2, 2, 140, 280)).astype(np.float32)
    w = rng.standard_normal((3, 2, 7, 9)).astype(np.float32)
    y0, _ = emul.conv(x, w, None, threads=256, flags=L.FC_FLAG_NO_FUSED)
    for fl in (L.FC_FLAG_NO_FAST_C2R | L.FC_FLAG_NO_FUSED_MID, L.FC_FLAG_NO_FAST_R2C | L.FC_FLAG_NO_FUSED_MID,
               L.FC_FLAG_NO_FAST_R2C | L.FC_FLAG_NO_FAST_C2R):
        y1, p = emul.conv(x, w, None, threads=256, flags=fl)
        assert rel_err(y1, y0) < 2e-6, fl


def test_fast_kernels_transposed_row_lattice():
    """stride = dilation = 2 transposed conv at sizes that select the warp-FFT kernels: polyphase reduction keeps
    only the dense rows between the passes; the C2R kernel scatters them and fills the bias-only rows."""
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(13)
    x = rng.standard_normal((2, 4, 130, 270)).astype(np.float32)
    w = rng.standard_normal((4, 3, 3, 5)).astype(np.float32)
    b = rng.standard_normal(6).astype(np.float32)
    kw = dict(stride=2, dilation=2, padding=(1, 0), output_padding=(1, 0), groups=2)
    ref = O.fft_conv_transpose(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), **kw)
    y, p = emul.conv(x, w, b, transposed=True, threads=256, **kw)
    assert "fast_c2r" in p.describe() and "fused_axis" in p.describe()
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5
    y2, _ = emul.conv(x, w, b, transposed=True, threads=256, flags=L.FC_FLAG_NO_FUSED, **kw)
    assert rel_err(y2, ref) < 1e-5


_SHORT_ROW_SHAPES = [
    # last axis 64 / 128 / 256 real points: K1 / K4 with a group of M/8 lanes per row (M = 32, 64, 128)
    ((2, 2, 70, 60), (2, 2, 5, 7), {}, False),
    ((1, 3, 40, 100), (2, 3, 3, 9), {}, False),
    ((2, 2, 30, 200), (3, 2, 3, 11), dict(groups=1), False),
    ((1, 2, 50, 64), (2, 1, 3, 5), dict(groups=2, stride=(1, 2)), False),  # strided scatter in K4
    ((1, 2, 20, 20, 60), (2, 2, 3, 3, 5), {}, False),  # 3-d, x extent 64
    ((1, 2, 33, 50), (2, 2, 3, 4), dict(stride=(1, 2), dilation=(1, 2)), True),  # lattice on the last axis (K4 general path)
]


@pytest.mark.parametrize("xs,ws,kw,tr", _SHORT_ROW_SHAPES)
def test_short_row_kernels_match_generic_and_oracle(xs, ws, kw, tr):
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(51)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
    b = rng.standard_normal(cout).astype(np.float32)
    ofn = O.fft_conv_transpose if tr else O.fft_conv
    ref = ofn(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), **kw)
    y, p = emul.conv(x, w, b, transposed=tr, threads=256, **kw)
    d = p.describe()
    assert "fast_c2r" in d and (tr or "fast_r2c" in d), d
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5
    y2, p2 = emul.conv(x, w, b, transposed=tr, threads=256, flags=L.FC_FLAG_NO_FAST_R2C | L.FC_FLAG_NO_FAST_C2R, **kw)
    assert "fast_r2c" not in p2.describe() and "fast_c2r" not in p2.describe()
    assert rel_err(y2, ref) < 1e-5


_COLUMN_SHAPES = [
    # 1-d lines long enough for the four-step split (transform length 16384 .. 65536 = 64 x N2)
    ((1, 2, 9000), (2, 2, 33), {}, False),
    ((2, 2, 20000), (3, 2, 129), dict(padding=64, padding_mode="reflect"), False),
    ((1, 4, 40000), (4, 2, 9), dict(groups=2, stride=3, padding=5, dilation=2), False),
    ((1, 2, 12000), (2, 3, 17), dict(stride=1, padding=3), True),
    ((1, 2, 6000), (2, 1, 5), dict(stride=3, output_padding=2, groups=2), True),  # zero-stuffed signal
    ((1, 2, 10000), (2, 2, 7), dict(stride=2, dilation=2, padding=1), True),  # polyphase lattice on store
]


@pytest.mark.parametrize("xs,ws,kw,tr", _COLUMN_SHAPES)
def test_column_kernels_match_generic_and_oracle(xs, ws, kw, tr):
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(31)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
    b = rng.standard_normal(cout).astype(np.float32)
    ofn = O.fft_conv_transpose if tr else O.fft_conv
    ref = ofn(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), **kw)
    y, p = emul.conv(x, w, b, transposed=tr, threads=256, **kw)
    d = p.describe()
    assert "col_r2c_N64" in d and "col_c2r_N64" in d and ("fused_axis" in d or ("fast_c2c_fwd" in d and "fast_c2c_inv" in d)), d
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5
    y2, p2 = emul.conv(x, w, b, transposed=tr, threads=256, flags=L.FC_FLAG_NO_FUSED, **kw)
    assert "col_" not in p2.describe() and "fast_" not in p2.describe()
    assert rel_err(y2, ref) < 1e-5
    y3, p3 = emul.conv(x, w, b, transposed=tr, threads=256, flags=L.FC_FLAG_NO_FUSED_MID, **kw)
    assert "fast_c2c_fwd" in p3.describe() and "fast_c2c_inv" in p3.describe()
    assert rel_err(y3, ref) < 1e-5
    print("column", rel_err(y, ref), "generic", rel_err(y2, ref))


_C2C_SHAPES = [
    # first axis 256 .. 2048 with the fused middle switched off: the contiguous complex passes run on fc_fast_c2c_kernel
    ((1, 2, 200, 40), (2, 2, 9, 3), {}, False),
    ((2, 3, 300, 40), (3, 3, 5, 3), dict(padding=(2, 1), padding_mode="reflect"), False),  # non-trivial gather map
    ((1, 2, 600, 36), (2, 1, 4, 3), dict(groups=2, stride=(3, 1), padding=(5, 0)), False),  # strided scatter on store
    ((1, 1, 1100, 34), (1, 1, 3, 3), {}, False),  # N = 2048: 64 points per lane
    ((1, 2, 150, 40), (2, 2, 3, 3), dict(stride=2, padding=1, output_padding=1), True),  # zero-stuffed signal
    ((1, 2, 140, 40), (2, 1, 5, 3), dict(stride=2, dilation=2, groups=2), True),  # polyphase lattice on store
    # short lines: a group of N/8 lanes per line (N = 32, 64, 128), several lines per warp, ragged line counts
    ((3, 2, 30, 20), (2, 2, 3, 3), {}, False),
    ((2, 3, 50, 36), (3, 3, 5, 3), dict(padding=(2, 1), padding_mode="circular"), False),
    ((1, 2, 100, 20), (2, 1, 4, 3), dict(groups=2, stride=(3, 1), padding=(5, 0)), False),
    ((1, 2, 40, 20), (2, 2, 3, 3), dict(stride=2, padding=1, output_padding=1), True),
    ((1, 2, 20, 130, 12), (2, 2, 3, 3, 3), {}, False),  # 3-d: the z passes (y extent 256, so no plane kernel)
    ((1, 1, 20, 130, 10), (1, 2, 5, 3, 3), dict(stride=(2, 1, 1), dilation=(2, 1, 1)), True),  # (y extent 256 again)
]


@pytest.mark.parametrize("xs,ws,kw,tr", _C2C_SHAPES)
def test_fast_c2c_pass_matches_generic_and_oracle(xs, ws, kw, tr):
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(21)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
    b = rng.standard_normal(cout).astype(np.float32)
    ofn = O.fft_conv_transpose if tr else O.fft_conv
    ref = ofn(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), **kw)
    y, p = emul.conv(x, w, b, transposed=tr, threads=256, flags=L.FC_FLAG_NO_FUSED_MID, **kw)
    d = p.describe()
    assert "fast_c2c_fwd" in d and "fast_c2c_inv" in d, d
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5
    y2, p2 = emul.conv(x, w, b, transposed=tr, threads=256, flags=L.FC_FLAG_NO_FUSED_MID | L.FC_FLAG_NO_FAST_C2C, **kw)
    assert "fast_c2c" not in p2.describe()
    assert rel_err(y2, ref) < 1e-5


_PLANE_SHAPES = [
    # 3-d problems whose y and z transform extents are 32 or 64: the two middle passes run as one plane kernel
    ((2, 2, 20, 20, 20), (2, 2, 3, 3, 3), {}, False),
    ((1, 3, 40, 20, 18), (2, 3, 5, 3, 3), dict(padding=(2, 1, 0)), False),  # 64 x 32 plane, zero padding on load
    ((1, 2, 24, 50, 12), (2, 2, 3, 7, 3), dict(padding=(0, 3, 1)), False),  # 32 x 64 plane
    ((1, 2, 60, 60, 10), (2, 1, 17, 17, 3), dict(groups=2), False),  # 64 x 64 plane
    ((1, 2, 20, 20, 20), (2, 2, 3, 3, 3), dict(padding=1), True),  # transposed, stride 1: crop offset on store
]


@pytest.mark.parametrize("xs,ws,kw,tr", _PLANE_SHAPES)
def test_plane_kernels_match_generic_and_oracle(xs, ws, kw, tr):
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(41)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
    b = rng.standard_normal(cout).astype(np.float32)
    ofn = O.fft_conv_transpose if tr else O.fft_conv
    ref = ofn(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), **kw)
    y, p = emul.conv(x, w, b, transposed=tr, threads=256, **kw)
    d = p.describe()
    assert "plane_fwd" in d and "plane_inv" in d and p.info.n_launches == 5, d
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5
    y2, p2 = emul.conv(x, w, b, transposed=tr, threads=256, flags=L.FC_FLAG_NO_FAST_C2C, **kw)
    assert "plane" not in p2.describe()
    assert rel_err(y2, ref) < 1e-5


@pytest.mark.parametrize("xs,ws,groups", [((10, 40, 70), (36, 40, 5), 1), ((17, 70, 40), (66, 35, 3), 2)])
def test_wide_channel_contraction_tiling(xs, ws, groups):
    """>= 32 output channels per group select the layout with several output tiles per CTA (BASELINE c4's
    contraction); ragged batch / channel tails included."""
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(14)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    y, _ = emul.conv(x, w, b, threads=64, groups=groups)
    ref = O.fft_conv(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), groups=groups)
    assert not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5


_SEGMENT_SHAPES = [
    # first axis long enough (and the kernel short enough) that the fused axis kernel runs overlap-save segments
    ((2, 2, 530, 40), (3, 2, 9, 3), {}, False),
    ((1, 12, 530, 36), (12, 12, 7, 3), {}, False),  # 9..16 channels per group: the one-bin-per-thread contraction
    ((1, 18, 520, 36), (18, 9, 5, 3), dict(stride=(2, 2), dilation=(2, 2), groups=2), True),  # BASELINE c5 in small
    ((2, 3, 700, 40), (3, 3, 5, 3), dict(padding=(2, 1), padding_mode="reflect"), False),  # general gather map
    ((1, 2, 1200, 36), (2, 1, 4, 3), dict(groups=2, stride=(3, 1), padding=(5, 0)), False),  # strided scatter on store
    ((1, 2, 330, 40), (2, 2, 3, 3), dict(stride=2, padding=1, output_padding=1), True),  # zero-stuffed signal
    ((1, 2, 5000, 34), (2, 2, 31, 3), {}, False),  # longer than any single-line transform of this axis
    ((3, 2, 600, 34), (2, 2, 17, 3), dict(padding=(40, 0)), True),  # crop larger than the segment overlap
    # ... and segments on the last axis as well: the row kernels K1 / K4 treat a (row, segment) pair as a line (the strided /
    # dilated / 2500-point variants of these run on the device only: tests/test_gpu_parity.py _SEGMENT_CASES)
    ((1, 1, 300, 600), (2, 1, 9, 5), {}, False),
    ((1, 3, 130, 600), (3, 3, 3, 7), dict(stride=(2, 2), dilation=(2, 2)), True),  # BASELINE c5 in small
    ((1, 2, 130, 700), (2, 2, 3, 4), {}, False),  # even kernel extent: segment stride rounded down to even
    ((1, 1, 130, 1100), (1, 1, 3, 10), dict(padding=(0, 20)), True),  # crop across a segment boundary
    ((1, 2, 130, 700), (2, 2, 3, 5), dict(padding=(1, 2)), False),  # zero padding inside the first / last row segment
]


@pytest.mark.parametrize("xs,ws,kw,tr", _SEGMENT_SHAPES)
def test_overlap_save_segments_match_unsegmented_and_oracle(xs, ws, kw, tr):
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(61)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
    b = rng.standard_normal(cout).astype(np.float32)
    ofn = O.fft_conv_transpose if tr else O.fft_conv
    ref = ofn(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), **kw)
    y, p = emul.conv(x, w, b, transposed=tr, threads=256, **kw)
    assert p.info.segments > 1 and p.info.fused == 1 and "segments(n=" in p.describe(), p.describe()
    assert p.fft_size[0] in (256, 512, 1024)
    if xs[3] >= 600:
        assert "axis1" in [ln.split(":")[0].strip() for ln in p.describe().splitlines() if "segments(n=" in ln]
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5
    if xs[2] <= 4096 and xs[3] <= 640 and xs[1] <= 3:  # the unsegmented layout of the same call (one transform over the whole axis)
        y2, p2 = emul.conv(x, w, b, transposed=tr, threads=256, flags=L.FC_FLAG_NO_SEGMENT, **kw)
        assert p2.info.segments == 1 and "segments(n=" not in p2.describe()
        assert rel_err(y2, ref) < 1e-5
        assert rel_err(y, y2) < 5e-6


def test_segmented_plans_refuse_the_stage_calls():
    p = emul.plan_for((1, 2, 530, 40), (2, 2, 9, 3))
    assert p.info.segments > 1
    lb = emul.lib()
    buf = np.zeros(16, np.float32)
    rc = lb.fc_contract(p.handle, emul._ptr(buf), emul._ptr(buf), emul._ptr(buf), None)
    assert rc == -2 and b"FC_FLAG_NO_SEGMENT" in lb.fc_last_error()
    assert emul.plan_for((1, 2, 530, 40), (2, 2, 9, 3), flags=L.FC_FLAG_NO_SEGMENT).info.segments == 1


@pytest.mark.parametrize("xs,ws,kw,flags,expect", [
    ((3, 8, 200, 180), (8, 8, 5, 7), dict(padding=(1, 2)), L.FC_FLAG_PAIR, "pair_fused_N256"),              # packed batch pairs, odd batch
    ((2, 8, 130, 250), (8, 8, 7, 3), {}, L.FC_FLAG_PAIR, "pair_fused64_N256"),                               # y stage, 64-point sub-problems
    ((2, 8, 300, 140), (8, 8, 9, 5), {}, L.FC_FLAG_PAIR, "pair_fused64_N512"),                               # y stage, 128-point sub-problems
])
def test_pair_programs_on_the_emulation(xs, ws, kw, flags, expect):
    """The packed batch-pair kernels (fc_pair.cuh) and the y-stage program, same source on host threads."""
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(2)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    y, plan = emul.conv(x, w, b, flags=flags, **kw)
    assert expect in plan.describe()
    assert not np.isnan(y).any()
    assert rel_err(y, O.fft_conv(x, w, b, **kw)) < 1e-5


def test_bias_only_rows_come_from_the_fused_kernel():
    """Transposed row lattice: the bias-only rows are written by the fused kernel (default) or by the last one (flag)."""
    from oracle import fftconv_oracle as O

    rng = np.random.RandomState(4)
    x = rng.standard_normal((2, 16, 130, 136)).astype(np.float32)
    w = rng.standard_normal((16, 4, 5, 5)).astype(np.float32)
    b = rng.standard_normal(8).astype(np.float32)
    kw = dict(groups=2, stride=2, dilation=2, padding=1)
    ref = O.fft_conv_transpose(x, w, b, **kw)
    for flags in (0, L.FC_FLAG_NO_ROW_FILL):
        y, plan = emul.conv(x, w, b, transposed=True, flags=flags, **kw)
        assert "fused_axis" in plan.describe() and not np.isnan(y).any()
        assert rel_err(y, ref) < 1e-5


# ---- 1-d overlap-save with the segments as extra batch items (fc_plan.cpp "batch segments")
_BSEG_CASES = [
    # more than 16 channels per group: windows of 1024 ... 4096 points through the one-pass layout
    ((2, 20, 9000), (4, 20, 33), {}),
    ((2, 3, 9000), (18, 3, 33), dict(padding=16)),
    ((1, 40, 9001), (4, 20, 40), dict(padding=7, stride=3, groups=2)),
    ((1, 17, 20000), (2, 17, 300), dict(padding=150, dilation=2)),
    # channel groups the fused axis kernel serves: windows of 16384 points, the four-step layout with the column kernels
    ((1, 2, 70000), (3, 2, 129), dict(padding=64)),
    ((2, 2, 70001), (2, 2, 4100), dict(stride=2)),
]


@pytest.mark.parametrize("xs,ws,kw", _BSEG_CASES)
def test_batch_segments_match_direct_convolution(xs, ws, kw):
    import torch
    import torch.nn.functional as F

    rng = np.random.RandomState(1)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    y, plan = emul.conv(x, w, b, **kw)
    assert "batch segments" in plan.describe()
    assert int(plan.info.segments) > 1 and plan.out_size == (y.shape[-1],)
    ref = F.conv1d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), **kw).numpy()
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5
    y1, plan1 = emul.conv(x, w, b, flags=L.FC_FLAG_NO_SEGMENT, **kw)  # the one-transform plan of the same problem
    assert "batch segments" not in plan1.describe()
    assert rel_err(y, y1) < 1e-5


def test_batch_segment_choice():
    """BASELINE c4 runs as 5 windows of 16384 points (kernel spectrum 4x smaller); BASELINE c1 (8 channels) keeps one
    transform; the choice never depends on the batch (the host pipeline's batch chunks share the kernel spectrum)."""
    c4 = emul.plan_for((16, 256, 65536), (256, 256, 4097))
    assert int(c4.info.segments) == 5 and c4.fft_size == (16384,) and c4.out_size == (61440,)
    for B in (1, 3, 40):
        p = emul.plan_for((B, 256, 65536), (256, 256, 4097))
        assert (int(p.info.segments), p.fft_size, int(p.info.kspec_bytes)) == (5, (16384,), int(c4.info.kspec_bytes))
    # channel groups the fused axis kernel serves: windows of 16384 points only; a small call (BASELINE c1) takes them even at
    # 1.5 x the points (more CTAs on a latency-bound problem), a large batch of the same lines only when they move fewer bytes
    c1 = emul.plan_for((1, 8, 32768), (8, 8, 1025))
    assert int(c1.info.segments) == 3 and c1.fft_size == (16384,) and c1.out_size == (31744,)
    big = emul.plan_for((64, 8, 32768), (8, 8, 1025))
    assert int(big.info.segments) == 1 and big.fft_size == (32768,)
    for flags, segs in ((L.FC_FLAG_NO_SEGMENT, 1), (L.FC_FLAG_SEGMENT, 3)):  # what the host pipeline pins for its batch chunks
        for B in (1, 64):
            assert int(emul.plan_for((B, 8, 32768), (8, 8, 1025), flags=flags).info.segments) == segs
    just_above = emul.plan_for((4, 32, 33000), (32, 32, 64))
    assert int(just_above.info.segments) > 1 and just_above.fft_size[0] <= 8192
    small = emul.plan_for((4, 8, 33000), (8, 8, 64))  # 8 channels: only windows that keep the fused-kernel program
    assert int(small.info.segments) == 1 or small.fft_size[0] >= 16384


# ---- 1-d lines of 2048 ... 8192 points on the four-step layout (64 x 32 ... 128: column + warp-engine kernels)
_SHORT_SPLIT_CASES = [
    ((2, 64, 4000), (64, 64, 33), dict(padding=16)),
    ((1, 40, 8000), (24, 20, 100), dict(groups=2, stride=2)),
    ((3, 7, 2048), (6, 7, 7), {}),
    ((1, 12, 2000), (8, 12, 9), dict(padding=4, dilation=3)),
    ((2, 4, 3000), (4, 4, 40), dict(padding=20, padding_mode="reflect")),
]


@pytest.mark.parametrize("xs,ws,kw", _SHORT_SPLIT_CASES)
def test_short_lines_on_the_four_step_layout(xs, ws, kw):
    import torch
    import torch.nn.functional as F

    rng = np.random.RandomState(2)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    y, plan = emul.conv(x, w, b, **kw)
    d = plan.describe()
    assert "structure=2" in d and "col_r2c_N64" in d and "col_c2r_N64" in d, d
    y1, plan1 = emul.conv(x, w, b, flags=L.FC_FLAG_NO_SHORT_SPLIT, **kw)
    assert "structure=1" in plan1.describe()
    kw2 = dict(kw)
    mode = kw2.pop("padding_mode", "constant")
    xt = torch.from_numpy(x).double()
    if mode != "constant":
        p = kw2.pop("padding")
        xt = F.pad(xt, (p, p), mode=mode)
    ref = F.conv1d(xt, torch.from_numpy(w).double(), torch.from_numpy(b).double(), **kw2).numpy()
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5 and rel_err(y, y1) < 1e-5


@pytest.mark.parametrize("xs,ws,kw", [((1, 1, 70, 100, 12), (2, 1, 3, 5, 3), {}), ((1, 1, 100, 40, 10), (1, 1, 5, 3, 3), dict(padding=(2, 1, 1)))])
def test_plane_kernels_with_128_point_axes(xs, ws, kw):
    """3-d programs whose middle axes are 128 points long run on the plane kernels (a 128 x 128 plane is 129 KB of shared memory)."""
    import torch
    import torch.nn.functional as F

    rng = np.random.RandomState(3)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    y, plan = emul.conv(x, w, b, **kw)
    d = plan.describe()
    assert "plane_fwd_" in d and "plane_inv_" in d and "128" in d, d
    ref = F.conv3d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), **kw).numpy()
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5


@pytest.mark.parametrize("xs,ws,kw,line_out", [((5, 6, 1000), (4, 6, 33), dict(padding=7), True), ((3, 4, 512), (4, 2, 9), dict(groups=2), True),
                                               ((1, 3, 700), (5, 3, 100), dict(padding=50, dilation=2), True), ((7, 2, 1024), (2, 2, 5), {}, True),
                                               ((2, 2, 900), (2, 2, 5), dict(stride=2), False)])
def test_short_1d_lines_on_the_warp_engine(xs, ws, kw, line_out):
    """One-pass 1-d programs of 512 / 1024 points: fc_line.cuh (a strided output keeps the generic last pass)."""
    import torch
    import torch.nn.functional as F

    rng = np.random.RandomState(4)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    y, plan = emul.conv(x, w, b, **kw)
    d = plan.describe()
    assert "line_r2c_N" in d and ("line_c2r_N" in d) == line_out, d
    y1, plan1 = emul.conv(x, w, b, flags=L.FC_FLAG_NO_FAST_R2C | L.FC_FLAG_NO_FAST_C2R, **kw)
    assert "line_" not in plan1.describe()
    ref = F.conv1d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), **kw).numpy()
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5 and rel_err(y, y1) < 2e-6


@pytest.mark.parametrize("xs,ws,kw", [((9, 48, 8100), (13, 48, 5), {}), ((8, 100, 120, 60), (26, 50, 3, 3), dict(groups=2, padding=1))])
def test_shared_memory_tiled_contraction(xs, ws, kw):
    """Wide channel groups on >= 4096 bins off the tensor-core path: fc_contract_tiled_kernel (ragged batch / channel tiles)."""
    import torch
    import torch.nn.functional as F

    rng = np.random.RandomState(5)
    x = rng.standard_normal(xs).astype(np.float32)
    w = rng.standard_normal(ws).astype(np.float32)
    b = rng.standard_normal(ws[0]).astype(np.float32)
    y, plan = emul.conv(x, w, b, **kw)
    assert int(plan.info.bins) >= 4096 and "launch contract" in plan.describe()
    conv = F.conv1d if len(xs) == 3 else F.conv2d
    ref = conv(torch.from_numpy(x).double(), torch.from_numpy(w).double(), torch.from_numpy(b).double(), **kw).numpy()
    assert y.shape == ref.shape and not np.isnan(y).any()
    assert rel_err(y, ref) < 1e-5
