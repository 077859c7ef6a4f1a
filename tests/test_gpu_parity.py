"""Parity of the CUDA path (through the C ABI) with the oracle, the reference's golden outputs and torch's direct
convolution, on a real B200.  Tolerance: max|y - ref| / max|ref| <= 1e-4 (BASELINE.json north_star);
small grid cases additionally keep the reference's own absolute tolerance (benchmark_utils.py:53-57)."""
import itertools

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import functional as Fn
from tests.helpers import golden, rel_err, spot_check

pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(autouse=True, scope="module")
def _no_tf32():
    # direct convolutions used as the expected value must not run in TF32
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_native_library_is_loaded():
    import ctypes

    lib = Fn.L.load()
    assert isinstance(lib, ctypes.CDLL)
    assert torch.cuda.get_device_capability()[0] >= 10, "these kernels are built for sm_100a only"


@pytest.mark.parametrize("i", range(len(golden())))
def test_golden_cases(i):
    c = golden().case(i)
    fn = fcp.fft_conv_transpose if c["transposed"] else fcp.fft_conv
    with torch.no_grad():
        y = fn(_dev(c["x"]), _dev(c["w"]), _dev(c["b"]), **c["kw"])
    assert y.is_contiguous() and y.dtype == torch.float32
    y = y.cpu().numpy()
    assert y.shape == c["direct"].shape
    assert rel_err(y, c["direct"]) < TOL
    if not c["ref_short"]:
        assert rel_err(y, c["ref32"]) < TOL
        if c["x"].size < 5000:
            d = np.abs(y - c["ref32"])
            assert d.mean() < 5e-5 and d.max() < 1e-4


def _gcd(a, b):
    while b:
        a, b = b, a % b
    return a


@pytest.mark.parametrize("ndim", [1, 2, 3])
@pytest.mark.parametrize("input_size", [7, 8])
def test_reference_forward_grid(ndim, input_size):
    """The reference's full forward grid (reference tests/test_functional.py:11-20) against F.conv{n}d."""
    torch.manual_seed(0)
    conv = getattr(F, f"conv{ndim}d")
    worst = 0.0
    for cin, cout, groups, k, p, s, d in itertools.product([2, 3], [2, 3], [1, 2, 3], [2, 3], [0, 1], [1, 2], [1, 2]):
        g = _gcd(cin, _gcd(cout, groups))
        x = torch.randn(2, cin, *([input_size] * ndim), device="cuda")
        w = torch.randn(cout, cin // g, *([k] * ndim), device="cuda")
        b = torch.randn(cout, device="cuda")
        with torch.no_grad():
            y0 = fcp.fft_conv(x, w, bias=b, padding=p, stride=s, dilation=d, groups=g)
            y1 = conv(x, w, bias=b, padding=p, stride=s, dilation=d, groups=g)
        assert y0.shape == y1.shape
        err = (y0 - y1).abs()
        assert err.mean().item() < 5e-5 and err.max().item() < 1e-4, (cin, cout, g, k, p, s, d)
        worst = max(worst, err.max().item())
    print("worst abs err", worst)


@pytest.mark.parametrize("ndim", [1, 2, 3])
@pytest.mark.parametrize("input_size", [7, 8])
def test_reference_transposed_grid(ndim, input_size):
    """The reference's transposed grid (reference tests/test_functional_transpose.py:60-88) against F.conv_transpose{n}d."""
    torch.manual_seed(1)
    conv = getattr(F, f"conv_transpose{ndim}d")
    for cin, cout, groups, k, p, op, s, d in itertools.product([2, 3], [2, 3], [1, 2, 3], [2, 3], [0, 1], [0, 1, 2], [1, 2], [1, 2]):
        d, s = d + op, s + op
        if ndim == 3 and s > 3:
            continue
        g = _gcd(cin, _gcd(cout, groups))
        x = torch.randn(2, cin, *([input_size] * ndim), device="cuda")
        w = torch.randn(cin, cout // g, *([k] * ndim), device="cuda")
        b = torch.randn(cout, device="cuda")
        kw = dict(padding=p, output_padding=op, stride=s, dilation=d, groups=g)
        with torch.no_grad():
            y0 = fcp.fft_conv_transpose(x, w, bias=b, **kw)
            y1 = conv(x, w, bias=b, **kw)
        assert y0.shape == y1.shape
        err = (y0 - y1).abs()
        assert err.mean().item() < 5e-5 and err.max().item() < 1e-4, (cin, cout, g, k, kw)


@pytest.mark.parametrize("mode", ["zeros", "reflect", "replicate", "circular"])
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_modules_match_torch_layers(mode, ndim):
    torch.manual_seed(2)
    cls = getattr(fcp, f"FFTConv{ndim}d")
    ref_cls = getattr(torch.nn, f"Conv{ndim}d")
    size = {1: 50, 2: 20, 3: 10}[ndim]
    m = cls(4, 6, 3, stride=2, padding=2, dilation=1, groups=2, padding_mode=mode).cuda()
    r = ref_cls(4, 6, 3, stride=2, padding=2, dilation=1, groups=2, padding_mode=mode).cuda()
    r.load_state_dict(m.state_dict())
    x = torch.randn(3, 4, *([size] * ndim), device="cuda")
    with torch.no_grad():
        y0, y1 = m(x), r(x)
    assert y0.shape == y1.shape
    assert rel_err(y0.cpu().numpy(), y1.cpu().numpy()) < TOL


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_transposed_modules_match_torch_layers(ndim):
    torch.manual_seed(3)
    cls = getattr(fcp, f"FFTConvTranspose{ndim}d")
    ref_cls = getattr(torch.nn, f"ConvTranspose{ndim}d")
    size = {1: 50, 2: 20, 3: 10}[ndim]
    m = cls(4, 6, 3, stride=3, padding=1, output_padding=2, dilation=2, groups=2).cuda()
    r = ref_cls(4, 6, 3, stride=3, padding=1, output_padding=2, dilation=2, groups=2).cuda()
    r.load_state_dict(m.state_dict())
    x = torch.randn(3, 4, *([size] * ndim), device="cuda")
    with torch.no_grad():
        y0, y1 = m(x), r(x)
    assert y0.shape == y1.shape
    assert rel_err(y0.cpu().numpy(), y1.cpu().numpy()) < TOL


def test_anisotropic_arguments_and_no_bias():
    torch.manual_seed(4)
    x = torch.randn(2, 6, 33, 47, device="cuda")
    w = torch.randn(9, 2, 5, 3, device="cuda")
    kw = dict(stride=(2, 3), padding=(4, 1), dilation=(3, 2), groups=3)
    with torch.no_grad():
        y0 = fcp.fft_conv(x, w, None, **kw)
        y1 = F.conv2d(x, w, None, **kw)
    assert rel_err(y0.cpu().numpy(), y1.cpu().numpy()) < TOL
    w1 = torch.randn(6, 1, 1, 1, device="cuda")  # kernel size 1, depthwise
    with torch.no_grad():
        assert rel_err(fcp.fft_conv(x, w1, groups=6).cpu().numpy(), F.conv2d(x, w1, groups=6).cpu().numpy()) < TOL


def test_kernel_spectrum_cache_tracks_weight_version():
    torch.manual_seed(5)
    m = fcp.FFTConv1d(2, 2, 9).cuda()
    x = torch.randn(1, 2, 100, device="cuda")
    with torch.no_grad():
        y0 = m(x).clone()
        n0 = Fn.launches()
        y0b = m(x)
        per_call = Fn.launches() - n0  # cached: no kernel-spectrum launches
        m.weight.mul_(2.0)  # in-place edit bumps the version counter
        y1 = m(x)
        ref = F.conv1d(x, m.weight, m.bias)
    assert torch.equal(y0, y0b)
    assert per_call == 3  # R2C, contraction, C2R
    assert rel_err(y1.cpu().numpy(), ref.cpu().numpy()) < TOL


def test_host_buffer_path_roundtrip():
    """CPU tensors in -> CPU tensor out through fc_conv_host (what a CPU-tensor caller of the reference sees)."""
    torch.manual_seed(6)
    x = torch.randn(2, 3, 40, 36)
    w = torch.randn(4, 3, 5, 5)
    b = torch.randn(4)
    with torch.no_grad():
        y = fcp.fft_conv(x, w, b, padding=2)
    assert not y.is_cuda
    assert rel_err(y.numpy(), F.conv2d(x, w, b, padding=2).numpy()) < TOL


def test_complex_matmul_matches_einsum():
    torch.manual_seed(7)
    a = torch.randn(3, 6, 5, 7, dtype=torch.complex64, device="cuda")
    b = torch.randn(8, 3, 5, 7, dtype=torch.complex64, device="cuda")
    y = fcp.complex_matmul(a, b, groups=2)
    ref = torch.einsum("bgi...,goi...->bgo...", a.unflatten(1, [2, 3]), b.unflatten(0, [2, 4])).flatten(1, 2)
    assert rel_err(torch.view_as_real(y).cpu().numpy(), torch.view_as_real(ref).cpu().numpy()) < 1e-5


def test_empty_batch_noncontiguous_input_and_side_stream():
    torch.manual_seed(12)
    w = torch.randn(4, 3, 5, 5, device="cuda")
    b = torch.randn(4, device="cuda")
    with torch.no_grad():
        y = fcp.fft_conv(torch.empty(0, 3, 20, 20, device="cuda"), w, b, padding=1)
        assert y.shape == (0, 4, 18, 18)
        x = torch.randn(2, 20, 3, 24, device="cuda").permute(0, 2, 3, 1)  # (2, 3, 24, 20), non-contiguous
        assert not x.is_contiguous()
        ref = F.conv2d(x, w, b, stride=2)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            y = fcp.fft_conv(x, w, b, stride=2)
        torch.cuda.current_stream().wait_stream(s)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < TOL


def test_inputs_are_not_mutated():
    torch.manual_seed(8)
    x = torch.randn(1, 4, 30, device="cuda")
    w = torch.randn(4, 2, 5, device="cuda")
    x0, w0 = x.clone(), w.clone()
    with torch.no_grad():
        fcp.fft_conv_transpose(x, w, stride=2, groups=2)
    assert torch.equal(x, x0) and torch.equal(w, w0)


# ------------------------------------------------------------------------------------------- BASELINE.json shapes
def _seeded(shape_x, shape_w, cout, seed=0):
    g = torch.Generator().manual_seed(seed)  # order: signal, weight, bias
    x = torch.randn(*shape_x, generator=g)
    w = torch.randn(*shape_w, generator=g)
    b = torch.randn(cout, generator=g)
    return x, w, b


def test_baseline_c1_full():
    x, w, b = _seeded((1, 8, 32768), (8, 8, 1025), 8)
    with torch.no_grad():
        y = fcp.fft_conv(x.cuda(), w.cuda(), b.cuda())
        ref = F.conv1d(x.double(), w.double(), b.double())
    assert y.shape == (1, 8, 31744)
    assert rel_err(y.cpu().numpy(), ref.numpy()) < TOL


def test_baseline_c2_full():
    torch.manual_seed(0)
    m = fcp.FFTConv2d(8, 8, 65).cuda()
    x = torch.randn(8, 8, 512, 512, device="cuda")
    with torch.no_grad():
        y = m(x)
        ref = F.conv2d(x, m.weight, m.bias)
    assert y.shape == (8, 8, 448, 448)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < TOL
    err, n = spot_check(x.cpu().numpy(), m.weight.detach().cpu().numpy(), m.bias.detach().cpu().numpy(), y, n=200)
    assert err < TOL, (err, n)
    # the host-buffer path bench.py times as `e2e` (pinned tensors, batch-chunked pipeline) gives the same numbers
    with torch.no_grad():
        yh = m(x.cpu().pin_memory())
    assert not yh.is_cuda and rel_err(yh.numpy(), y.cpu().numpy()) < 1e-6


def test_baseline_c3_full():
    torch.manual_seed(0)
    m = fcp.FFTConv3d(8, 8, 17).cuda()
    x = torch.randn(4, 8, 64, 64, 64, device="cuda")
    with torch.no_grad():
        y = m(x)
        ref = F.conv3d(x, m.weight, m.bias)
    assert y.shape == (4, 8, 48, 48, 48)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < TOL


def test_baseline_c4_spot_checked_and_linear():
    """c4 (16,256,65536) x (256,256,4097): a direct convolution is intractable here, so the full-size result is
    checked at random output positions against the definition in fp64, and through linearity."""
    x, w, b = _seeded((16, 256, 65536), (256, 256, 4097), 256)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    with torch.no_grad():
        y = fcp.fft_conv(xd, wd, bd)
    assert y.shape == (16, 256, 61440)
    err, n = spot_check(x.numpy(), w.numpy(), b.numpy(), y, n=64)
    assert err < TOL, (err, n)
    with torch.no_grad():
        y2 = fcp.fft_conv(xd * 0.5, wd, None)
    lin = (y - bd.view(1, -1, 1)) * 0.5
    assert rel_err(y2[:2].cpu().numpy(), lin[:2].cpu().numpy()) < 1e-5


def test_baseline_c5_shard_spot_checked():
    """c5 per-GPU shard (B=4 of 32): fft_conv_transpose (4,64,1024,1024), kernel (64,16,31,31), stride 2, dilation 2,
    groups 4. Checked at random positions against the definition, plus the lattice property: with
    stride = dilation = 2 every odd output position is exactly the bias."""
    x, w, b = _seeded((4, 64, 1024, 1024), (64, 16, 31, 31), 64)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    with torch.no_grad():
        y = fcp.fft_conv_transpose(xd, wd, bd, stride=2, dilation=2, groups=4)
    assert y.shape == (4, 64, 2107, 2107)
    err, n = spot_check(x.numpy(), w.numpy(), b.numpy(), y, n=200, transposed=True, stride=2, dilation=2, groups=4)
    assert err < TOL, (err, n)
    odd = y[:, :, 1::2, :]
    assert torch.equal(odd, bd.view(1, -1, 1, 1).expand_as(odd))


_SEGMENT_CASES = [
    # 2-d problems whose first axis runs as overlap-save segments inside the fused axis kernel
    ((2, 8, 600, 300), (8, 8, 9, 7), {}, False),
    ((2, 12, 530, 200), (12, 12, 7, 3), {}, False),
    ((1, 32, 520, 260), (32, 16, 5, 3), dict(stride=(2, 2), dilation=(2, 2), groups=2), True),
    ((2, 3, 700, 140), (3, 3, 5, 3), dict(padding=(2, 1), padding_mode="reflect"), False),
    ((1, 2, 1200, 136), (2, 1, 4, 3), dict(groups=2, stride=(3, 1), padding=(5, 0)), False),
    ((1, 2, 330, 140), (2, 2, 3, 3), dict(stride=2, padding=1, output_padding=1), True),
    ((1, 4, 5000, 134), (4, 4, 31, 3), {}, False),
    ((3, 2, 600, 134), (2, 2, 17, 3), dict(padding=(40, 0)), True),
    # segments on both axes (row kernels with (row, segment) lines)
    ((2, 8, 300, 600), (8, 8, 9, 5), {}, False),
    ((1, 32, 300, 600), (32, 16, 5, 7), dict(stride=(2, 2), dilation=(2, 2), groups=2), True),
    ((1, 2, 130, 700), (2, 2, 3, 4), {}, False),
    ((1, 2, 270, 640), (2, 1, 4, 6), dict(groups=2, stride=(1, 3)), False),
    ((1, 2, 280, 660), (2, 2, 3, 4), dict(padding=(1, 3), dilation=(1, 2), output_padding=(0, 1)), True),
    ((2, 2, 260, 2500), (2, 2, 3, 33), {}, False),
    ((1, 3, 200, 1100), (3, 3, 3, 10), dict(padding=(0, 20)), True),
    ((2, 16, 1100, 1300), (16, 16, 15, 15), {}, False),
    ((2, 8, 640, 720), (8, 8, 9, 9), dict(padding=(4, 4)), False),  # "same" convolution
    ((1, 2, 600, 700), (2, 2, 5, 5), dict(dilation=(3, 2)), False),  # dilated kernel: longer segment overlap
    ((1, 2, 300, 4200), (2, 2, 1, 1), {}, False),  # 1 x 1 kernel: segments without overlap
]


@pytest.mark.parametrize("xs,ws,kw,tr", _SEGMENT_CASES)
def test_overlap_save_segments(xs, ws, kw, tr):
    x, w, b = _seeded(xs, ws, ws[1] * kw.get("groups", 1) if tr else ws[0], seed=3)
    fn = fcp.fft_conv_transpose if tr else fcp.fft_conv
    tfn = F.conv_transpose2d if tr else F.conv2d
    tkw = dict(kw)
    mode = tkw.pop("padding_mode", None)
    xd = x.double()
    if mode:
        pd = tkw.pop("padding")
        xd = F.pad(xd, (pd[1], pd[1], pd[0], pd[0]), mode=mode)
    with torch.no_grad():
        y = fn(x.cuda(), w.cuda(), b.cuda(), **kw)
        ref = tfn(xd, w.double(), b.double(), **tkw)
        y_one = Fn._run(tr, x.cuda(), w.cuda(), b.cuda(), kw.get("stride", 1), kw.get("padding", 0), kw.get("output_padding", 0),
                        kw.get("dilation", 1), kw.get("groups", 1), kw.get("padding_mode", "constant"), flags=Fn.L.FC_FLAG_NO_SEGMENT) \
            if xs[2] <= 4096 and xs[3] <= 4096 else None
    nd = 2
    tup = lambda v: tuple(v) if hasattr(v, "__iter__") else (v,) * nd
    cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
    entry = Fn.get_plan(tr, xs[0], xs[1], cout, kw.get("groups", 1), tuple(xs[2:]), tuple(ws[2:]), tup(kw.get("stride", 1)),
                        tup(kw.get("padding", 0)), tup(kw.get("dilation", 1)), tup(kw.get("output_padding", 0)), kw.get("padding_mode", "constant"))
    assert entry.plan.info.segments > 1 and entry.plan.info.fused == 1
    assert y.shape == ref.shape
    assert rel_err(y.cpu().numpy(), ref.numpy()) < TOL
    if y_one is not None:
        assert rel_err(y.cpu().numpy(), y_one.cpu().numpy()) < 1e-5


def _random_2d_case(rng):
    """A random 2-d problem in the size range where the plan chooses between segmented and unsegmented layouts, fast and
    generic kernels (odd / even extents and paddings, lattices, groups, channel bounds 8 / 16 / generic)."""
    tr = bool(rng.randint(2))
    groups = int(rng.choice([1, 1, 2, 4]))
    ig, og = int(rng.choice([1, 2, 3, 8, 12, 16, 20])), int(rng.choice([1, 2, 5, 8, 16]))
    cin, cout = ig * groups, og * groups
    k = (int(rng.randint(1, 34)), int(rng.randint(1, 34)))
    dil = (int(rng.choice([1, 1, 1, 2, 3])), int(rng.choice([1, 1, 1, 2, 3])))
    stride = (int(rng.choice([1, 1, 2, 3])), int(rng.choice([1, 1, 2, 3])))
    size = (int(rng.randint(120, 1500)), int(rng.randint(120, 1500)))
    size = tuple(max(s, (kk - 1) * d + 1) for s, kk, d in zip(size, k, dil))
    kw = dict(stride=stride, dilation=dil, groups=groups)
    if tr:
        kw["padding"] = tuple(int(rng.randint(0, min(6, (kk - 1) * d + 1))) for kk, d in zip(k, dil))
        kw["output_padding"] = tuple(int(rng.randint(0, max(s, d))) for s, d in zip(stride, dil))
        w_shape = (cin, og) + k
    else:
        kw["padding"] = (int(rng.randint(0, 9)), int(rng.randint(0, 9)))
        w_shape = (cout, ig) + k
    B = 1 if cin * cout * size[0] * size[1] > 6e7 else int(rng.randint(1, 4))
    return tr, (B, cin) + size, w_shape, cout, kw


@pytest.mark.parametrize("seed", range(48))
def test_random_2d_problems_match_torch(seed):
    rng = np.random.RandomState(1000 + seed)
    tr, xs, ws, cout, kw = _random_2d_case(rng)
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(*xs, device="cuda", generator=g)
    w = torch.randn(*ws, device="cuda", generator=g)
    b = torch.randn(cout, device="cuda", generator=g)
    with torch.no_grad():
        if tr:
            y = fcp.fft_conv_transpose(x, w, b, **kw)
            ref = F.conv_transpose2d(x.double(), w.double(), b.double(), **kw)
        else:
            y = fcp.fft_conv(x, w, b, **kw)
            ref = F.conv2d(x.double(), w.double(), b.double(), **kw)
    assert y.shape == ref.shape, (xs, ws, kw)
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < TOL, (err, tr, xs, ws, kw)


@pytest.mark.parametrize("seed", range(32))
def test_random_1d_and_3d_problems_match_torch(seed):
    """1-d lines up to the four-step layouts (64 x N2 column kernels, fused axis inside the split) and 3-d volumes (row
    kernels with lane groups, plane kernels), random arguments."""
    rng = np.random.RandomState(2000 + seed)
    nd = 1 if seed % 2 == 0 else 3
    tr = bool(rng.randint(2))
    groups = int(rng.choice([1, 1, 2, 4]))
    ig, og = int(rng.choice([1, 2, 3, 8])), int(rng.choice([1, 2, 5, 8]))
    cin, cout = ig * groups, og * groups
    if nd == 1:
        k = (int(rng.choice([1, 3, 17, 64, 129, 1025])),)
        size = (int(rng.choice([900, 5000, 9000, 20000, 40000, 70000, 150000])) + int(rng.randint(0, 50)),)
    else:
        k = tuple(int(rng.randint(1, 8)) for _ in range(3))
        size = tuple(int(rng.randint(12, 90)) for _ in range(3))
    dil = tuple(int(rng.choice([1, 1, 2, 3])) for _ in range(nd))
    stride = tuple(int(rng.choice([1, 1, 2, 3])) for _ in range(nd))
    size = tuple(max(s, (kk - 1) * d + 1) for s, kk, d in zip(size, k, dil))
    kw = dict(stride=stride, dilation=dil, groups=groups)
    if tr:
        kw["padding"] = tuple(int(rng.randint(0, min(4, (kk - 1) * d + 1))) for kk, d in zip(k, dil))
        kw["output_padding"] = tuple(int(rng.randint(0, max(s_, d))) for s_, d in zip(stride, dil))
        ws = (cin, og) + k
    else:
        kw["padding"] = tuple(int(rng.randint(0, 5)) for _ in range(nd))
        ws = (cout, ig) + k
    B = int(rng.randint(1, 4))
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, cin, *size, device="cuda", generator=g)
    w = torch.randn(*ws, device="cuda", generator=g)
    b = torch.randn(cout, device="cuda", generator=g)
    with torch.no_grad():
        if tr:
            y = fcp.fft_conv_transpose(x, w, b, **kw)
            ref = getattr(F, f"conv_transpose{nd}d")(x.double(), w.double(), b.double(), **kw)
        else:
            y = fcp.fft_conv(x, w, b, **kw)
            ref = getattr(F, f"conv{nd}d")(x.double(), w.double(), b.double(), **kw)
    assert y.shape == ref.shape, (size, ws, kw)
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < TOL, (err, tr, size, ws, kw)


# ------------------------------------------------------------------------------------------- backward
@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_backward_matches_torch_forward_conv(ndim):
    """reference tests/test_functional.py:72-117: weight / bias gradients of y.sum(); input gradient as well."""
    torch.manual_seed(9)
    conv = getattr(F, f"conv{ndim}d")
    for cin, cout, groups, k, p, s, d, size in [(2, 2, 1, 2, 0, 1, 1, 7), (3, 3, 3, 3, 1, 2, 2, 8), (2, 4, 2, 3, 1, 1, 2, 8), (3, 2, 1, 2, 1, 2, 1, 7)]:
        x0 = torch.randn(2, cin, *([size] * ndim), device="cuda", requires_grad=True)
        w0 = torch.randn(cout, cin // groups, *([k] * ndim), device="cuda", requires_grad=True)
        b0 = torch.randn(cout, device="cuda", requires_grad=True)
        x1, w1, b1 = (t.detach().clone().requires_grad_() for t in (x0, w0, b0))
        kw = dict(padding=p, stride=s, dilation=d, groups=groups)
        y0 = fcp.fft_conv(x0, w0, bias=b0, **kw)
        y1 = conv(x1, w1, bias=b1, **kw)
        y0.sum().backward()
        y1.sum().backward()
        for a, r in ((y0, y1), (w0.grad, w1.grad), (b0.grad, b1.grad), (x0.grad, x1.grad)):
            e = (a - r).abs()
            assert e.mean().item() < 5e-5 and e.max().item() < 1e-4


@pytest.mark.parametrize("ndim", [1, 2, 3])
def test_backward_matches_torch_transposed_conv(ndim):
    """reference tests/test_functional_transpose.py:73-124."""
    torch.manual_seed(10)
    conv = getattr(F, f"conv_transpose{ndim}d")
    for cin, cout, groups, k, p, op, s, d, size in [(2, 2, 1, 2, 0, 0, 1, 1, 7), (3, 3, 3, 3, 1, 1, 3, 3, 8), (2, 4, 2, 3, 1, 0, 2, 2, 8), (2, 3, 1, 2, 1, 1, 2, 2, 7)]:
        x0 = torch.randn(2, cin, *([size] * ndim), device="cuda", requires_grad=True)
        w0 = torch.randn(cin, cout // groups, *([k] * ndim), device="cuda", requires_grad=True)
        b0 = torch.randn(cout, device="cuda", requires_grad=True)
        x1, w1, b1 = (t.detach().clone().requires_grad_() for t in (x0, w0, b0))
        kw = dict(padding=p, output_padding=op, stride=s, dilation=d, groups=groups)
        y0 = fcp.fft_conv_transpose(x0, w0, bias=b0, **kw)
        y1 = conv(x1, w1, bias=b1, **kw)
        y0.sum().backward()
        y1.sum().backward()
        for a, r in ((y0, y1), (w0.grad, w1.grad), (b0.grad, b1.grad), (x0.grad, x1.grad)):
            e = (a - r).abs()
            assert e.mean().item() < 5e-5 and e.max().item() < 1e-4


def test_backward_through_segmented_plans():
    """Gradients at sizes where the forward call and its adjoint convolutions run as overlap-save segments."""
    torch.manual_seed(12)
    x0 = torch.randn(2, 4, 600, 700, device="cuda", requires_grad=True)
    w0 = torch.randn(6, 2, 7, 5, device="cuda", requires_grad=True)
    b0 = torch.randn(6, device="cuda", requires_grad=True)
    x1, w1, b1 = (t.detach().clone().requires_grad_() for t in (x0, w0, b0))
    gy = torch.randn(2, 6, 594, 696, device="cuda")
    y0 = fcp.fft_conv(x0, w0, bias=b0, groups=2)
    y1 = F.conv2d(x1, w1, bias=b1, groups=2)
    (y0 * gy).sum().backward()
    (y1 * gy).sum().backward()
    entry = Fn.get_plan(False, 2, 4, 6, 2, (600, 700), (7, 5), (1, 1), (0, 0), (1, 1), (0, 0), "constant")
    assert entry.plan.info.segments > 1
    for a, r in ((y0, y1), (w0.grad, w1.grad), (b0.grad, b1.grad), (x0.grad, x1.grad)):
        assert rel_err(a.detach().cpu().numpy(), r.detach().cpu().numpy()) < TOL


def test_module_backward_trains():
    torch.manual_seed(11)
    m = fcp.FFTConv2d(3, 4, 5, padding=2, padding_mode="reflect").cuda()
    r = torch.nn.Conv2d(3, 4, 5, padding=2, padding_mode="reflect").cuda()
    r.load_state_dict(m.state_dict())
    x = torch.randn(2, 3, 16, 16, device="cuda")
    m(x).square().mean().backward()
    r(x).square().mean().backward()
    assert rel_err(m.weight.grad.cpu().numpy(), r.weight.grad.cpu().numpy()) < TOL
    assert rel_err(m.bias.grad.cpu().numpy(), r.bias.grad.cpu().numpy()) < TOL


# ------------------------------------------------------------------------------------------------ round 2 additions
def _reference_fft_conv():
    """The unmodified reference (baseline/_ref travels to the GPU box with the snapshot); None when it is absent."""
    import os
    import sys
    import warnings

    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "fft_conv_pytorch")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    warnings.filterwarnings("ignore")
    from fft_conv_pytorch.functional import fft_conv as ref_fft_conv

    return ref_fft_conv


@pytest.mark.parametrize("name,xs,ws", [("c1", (1, 8, 32768), (8, 8, 1025)), ("c2", (8, 8, 512, 512), (8, 8, 65, 65)),
                                        ("c3", (4, 8, 64, 64, 64), (8, 8, 17, 17, 17))])
def test_baseline_c1_c2_c3_against_the_reference_itself(name, xs, ws):
    """North star: 'max relative error <= 1e-4 against torch's fft_conv and also against direct F.conv': the reference's
    own fft_conv evaluated in float64 on the same GPU on identical inputs."""
    ref_fft_conv = _reference_fft_conv()
    if ref_fft_conv is None:
        pytest.skip("baseline/_ref is not present")
    x, w, b = _seeded(xs, ws, ws[0])
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    with torch.no_grad():
        y = fcp.fft_conv(xd, wd, bd)
        ref = ref_fft_conv(xd.double(), wd.double(), bd.double())
    assert tuple(y.shape) == tuple(ref.shape)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < TOL


def test_baseline_c4_full_tensor_slice():
    """c4: every output of the first 8 output channels against cuDNN's direct convolution (fp32, TF32 off)."""
    x, w, b = _seeded((16, 256, 65536), (256, 256, 4097), 256)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    with torch.no_grad():
        y = fcp.fft_conv(xd, wd, bd)
        ref = F.conv1d(xd, wd[:8].contiguous(), bd[:8].contiguous())
    d = (y[:, :8].double() - ref.double()).abs().max().item() / ref.double().abs().max().item()
    assert d < TOL, d


def test_baseline_c5_full_tensor_one_item():
    """c5: every output of one batch item against cuDNN's direct transposed convolution (fp32, TF32 off)."""
    x, w, b = _seeded((4, 64, 1024, 1024), (64, 16, 31, 31), 64)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    with torch.no_grad():
        y = fcp.fft_conv_transpose(xd, wd, bd, stride=2, dilation=2, groups=4)
        ref = F.conv_transpose2d(xd[1:2], wd, bd, stride=2, dilation=2, groups=4)
    d = (y[1:2].double() - ref.double()).abs().max().item() / ref.double().abs().max().item()
    assert d < TOL, d


_PAIR_CASES = [
    # x, w, transposed, kwargs: shapes the packed batch-pair kernels cover (full groups of 8 or 16 channels)
    ((8, 8, 512, 512), (8, 8, 65, 65), False, {}),                                                   # BASELINE c2
    ((3, 8, 200, 180), (8, 8, 5, 7), False, dict(padding=(1, 2))),                                   # odd batch, padding
    ((1, 16, 300, 260), (16, 8, 9, 3), False, dict(groups=2, stride=(2, 1))),                        # single item, groups of 8
    ((2, 32, 560, 560), (32, 16, 9, 9), True, dict(stride=2, dilation=2, groups=2, padding=2)),      # 16 per group, segments, lattice
    ((5, 8, 700, 520), (8, 8, 31, 17), False, {}),                                                   # row and column segments
    ((2, 8, 200, 1700), (8, 8, 5, 301), False, {}),                                                  # 2048-point rows
    ((4, 16, 256, 256), (16, 16, 9, 9), False, dict(padding=4)),                                     # 16 per group, two items per CTA
    ((16, 8, 256, 256), (8, 8, 31, 31), False, {}),                                                  # y stage with 64-point sub-problems
    ((10, 8, 300, 600), (8, 8, 9, 5), False, {}),                                                    # y stage + row segments, item blocks
    ((3, 16, 700, 200), (16, 8, 5, 5), False, dict(groups=2)),                                       # y stage, N = 1024 when unsegmented
]


def test_y_stage_program_is_what_the_pair_flag_runs_at_c2():
    from fft_conv_pytorch_b200 import _lib as L

    e = Fn.get_plan(False, 8, 8, 8, 1, (512, 512), (65, 65), (1, 1), (0, 0), (1, 1), (0, 0), "constant", L.FC_FLAG_PAIR)
    d = e.plan.describe()
    assert "pair_r2c_N512_ys4" in d and "pair_fused64_N512" in d and "pair_c2r_N512_ys4" in d, d


def test_bias_only_rows_written_by_the_fused_kernel_match():
    """BASELINE c5 geometry at reduced size: stride = dilation = 2 row lattice, rows filled by the fused kernel or the last one."""
    from fft_conv_pytorch_b200 import _lib as L

    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 32, 300, 280, generator=g).cuda()
    w = torch.randn(32, 8, 7, 7, generator=g).cuda()
    b = torch.randn(32, generator=g).cuda()
    kw = dict(stride=2, dilation=2, groups=4, padding=1, output_padding=1)
    out = []
    try:
        for flags in (0, L.FC_FLAG_NO_ROW_FILL):
            Fn.set_default_flags(flags)
            Fn.clear_caches()
            with torch.no_grad():
                out.append(fcp.fft_conv_transpose(x, w, b, **kw))
    finally:
        Fn.set_default_flags(0)
        Fn.clear_caches()
    with torch.no_grad():
        ref = F.conv_transpose2d(x.double(), w.double(), b.double(), **kw)
    assert torch.equal(out[0], out[1])
    assert (out[0].double() - ref).abs().max().item() / ref.abs().max().item() < TOL


@pytest.mark.parametrize("xs,ws,tr,kw", _PAIR_CASES)
def test_pair_kernels_match_one_line_kernels_and_torch(xs, ws, tr, kw):
    from fft_conv_pytorch_b200 import _lib as L

    g = torch.Generator().manual_seed(11)
    x = torch.randn(*xs, generator=g).cuda()
    w = torch.randn(*ws, generator=g).cuda()
    cout = ws[1] * kw.get("groups", 1) if tr else ws[0]
    b = torch.randn(cout, generator=g).cuda()
    fn = fcp.fft_conv_transpose if tr else fcp.fft_conv
    out = {}
    try:
        # pair: batch-pair kernels, with the y-stage program (64 / 128-point fused kernel) where it applies; pair_noys: the
        # whole fused-axis transform inside the pair fused kernel; plain: the one-line-per-item kernels
        for name, flags in (("pair", L.FC_FLAG_PAIR), ("pair_noys", L.FC_FLAG_PAIR | L.FC_FLAG_NO_YSTAGE), ("plain", L.FC_FLAG_NO_PAIR)):
            Fn.set_default_flags(flags)
            Fn.clear_caches()
            with torch.no_grad():
                out[name] = fn(x, w, b, **kw)
            if name != "plain":
                d = Fn._plans[next(reversed(Fn._plans))].plan.describe()
                assert "pair_fused" in d, d
                assert name == "pair" or "pair_fused64" not in d
    finally:
        Fn.set_default_flags(0)
        Fn.clear_caches()
    nd = len(xs) - 2
    with torch.no_grad():
        ref = getattr(F, ("conv_transpose%dd" if tr else "conv%dd") % nd)(x.double(), w.double(), b.double(), **kw)
    scale = ref.abs().max().item()
    for name in ("pair", "pair_noys"):
        assert (out[name].double() - ref).abs().max().item() / scale < TOL
        assert (out[name] - out["plain"]).abs().max().item() / scale < 2e-6


def test_host_pipeline_chunks_keep_the_full_batch_spectrum_layout():
    """the batch chunks of the host pipeline run on the kernel spectrum of the full-batch plan, so both must
    choose the same spectrum layout (round 1: a batch > 32 stayed on the SIMT layout while its chunks picked the
    tensor-core one; now batches > 32 run the tensor-core GEMM in chunks of 32 and the layouts agree by construction)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(40, 32, 2048, generator=g)
    w = torch.randn(128, 32, 9, generator=g)
    b = torch.randn(128, generator=g)
    with torch.no_grad():
        y = fcp.fft_conv(x.pin_memory(), w.cuda(), b.cuda())
        ref = F.conv1d(x.cuda(), w.cuda(), b.cuda())
    assert not y.is_cuda
    assert rel_err(y.numpy(), ref.cpu().numpy()) < TOL


def test_two_devices_in_one_process():
    """the dynamic shared-memory opt-in and the SM count are per device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 8, 300, 300, generator=g)
    w = torch.randn(8, 8, 9, 9, generator=g)
    ref = F.conv2d(x.double(), w.double()).numpy()
    for d in (0, 1):
        with torch.no_grad():
            y = fcp.fft_conv(x.to(f"cuda:{d}"), w.to(f"cuda:{d}"))
        assert rel_err(y.cpu().numpy(), ref) < TOL


def test_tensor_core_contraction_in_batch_chunks():
    """Batches above 80 run the tcgen05 GEMM in even chunks of up to 80 batch items (here 2 x 45); up to 80 in one."""
    g = torch.Generator().manual_seed(6)
    for B, marker in ((90, "tc_gemm_3xtf32_b45"), (40, "tc_gemm_3xtf32")):
        x = torch.randn(B, 32, 2048, generator=g).cuda()
        w = torch.randn(128, 32, 9, generator=g).cuda()
        b = torch.randn(128, generator=g).cuda()
        e = Fn.get_plan(False, B, 32, 128, 1, (2048,), (9,), (1,), (0,), (1,), (0,), "constant")
        d = e.plan.describe()
        assert int(e.plan.info.tensor_core) == 1 and marker in d, d
        with torch.no_grad():
            y = fcp.fft_conv(x, w, b)
            ref = F.conv1d(x.double(), w.double(), b.double())
        assert (y.double() - ref).abs().max().item() / ref.abs().max().item() < TOL


def test_transforms_next_to_the_tensor_core_gemm_write_its_operands():
    """Four-step 1-d plans with one GEMM chunk: the contiguous complex passes on either side of the GEMM write the Bt blobs /
    gather from the product themselves (no relayout kernels); same result as the program with the relayout kernels."""
    from fft_conv_pytorch_b200 import _lib as L

    g = torch.Generator().manual_seed(8)
    # (transform lengths 64 x 256 unsegmented with a padded batch, 64 x 512, and 16384-point windows with two groups)
    for xs, ws, kw, base in (((3, 64, 16384), (128, 64, 1000), {}, 0),
                             ((5, 32, 32768), (256, 32, 33), {}, L.FC_FLAG_NO_SEGMENT),
                             ((2, 64, 70001), (256, 32, 4100), dict(groups=2), 0)):
        x = torch.randn(*xs, generator=g).cuda()
        w = torch.randn(*ws, generator=g).cuda()
        b = torch.randn(ws[0], generator=g).cuda()
        out = {}
        try:
            for name, flags in (("fused", base), ("relayout", base | L.FC_FLAG_NO_FAST_C2C)):
                Fn.set_default_flags(flags)
                Fn.clear_caches()
                with torch.no_grad():
                    out[name] = fcp.fft_conv(x, w, b, **kw).clone()
                d = Fn._plans[next(reversed(Fn._plans))].plan.describe()
                assert "tc_gemm_3xtf32" in d, d
                assert ("tc_c2c_fwd" in d and "tc_c2c_inv" in d and "tc_relayout" not in d) == (name == "fused"), d
        finally:
            Fn.set_default_flags(0)
            Fn.clear_caches()
        with torch.no_grad():
            ref = F.conv1d(x.double(), w.double(), b.double(), **kw)
        assert (out["fused"].double() - ref).abs().max().item() / ref.abs().max().item() < TOL
        assert (out["fused"] - out["relayout"]).abs().max().item() / ref.abs().max().item() < 2e-5


# ---- streaming K1 / K4 (csrc/fc_stream.cuh: bulk-copy loads, tensor-map transposing copies) against the register-path
# kernels: the arithmetic is the same, so the outputs must be bit-identical; and against F.conv2d in float64.
_STREAM_CASES = [
    ((2, 8, 512, 512), (8, 8, 65, 65), {}),                       # BASELINE c2 geometry, two batch items: whole 32 KB tiles
    ((2, 8, 448, 448), (8, 8, 65, 65), dict(padding=32)),         # zero padding: per-row copies of 448 floats at offset 32
    ((3, 4, 200, 448), (4, 4, 17, 33), dict(padding=(8, 16))),    # 216 rows: partial last tile
    ((2, 8, 250, 500), (8, 8, 9, 13), {}),                        # 500-float rows (16-byte multiple), 250 rows
    ((2, 8, 250, 498), (8, 8, 9, 13), {}),                        # 498-float rows: K1 stays on the register path, K4 streams
    ((1, 8, 100, 512), (16, 8, 5, 5), {}),
]


@pytest.mark.parametrize("xs,ws,kw", _STREAM_CASES)
def test_streaming_row_kernels_match_register_path_bitwise(xs, ws, kw):
    from fft_conv_pytorch_b200 import _lib as L

    g = torch.Generator().manual_seed(17)
    x = torch.randn(*xs, generator=g).cuda()
    w = torch.randn(*ws, generator=g).cuda()
    b = torch.randn(ws[0], generator=g).cuda()
    out = {}
    try:
        # default: K4 streams; both: K1 streams too; none: register-path K1 / K4
        for name, flags in (("default", 0), ("both", L.FC_FLAG_STREAM_R2C), ("none", L.FC_FLAG_NO_STREAM)):
            Fn.set_default_flags(flags)
            Fn.clear_caches()
            with torch.no_grad():
                out[name] = fcp.fft_conv(x, w, b, **kw).clone()
    finally:
        Fn.set_default_flags(0)
        Fn.clear_caches()
    with torch.no_grad():
        ref = F.conv2d(x.double(), w.double(), b.double(), **kw)
    assert torch.equal(out["default"], out["none"])
    assert torch.equal(out["both"], out["none"])
    assert (out["default"].double() - ref).abs().max().item() / ref.abs().max().item() < TOL


# ---- 1-d overlap-save with the segments as extra batch items (fc_plan.cpp "batch segments")
_BSEG_GPU_CASES = [
    ((2, 24, 9000), (4, 24, 33), dict(padding=16)),
    ((1, 40, 9001), (4, 20, 40), dict(padding=7, stride=3, groups=2)),
    ((2, 8, 70001), (8, 8, 4100), dict(stride=2)),                 # 16384-point windows, fused axis kernel
    ((1, 8, 70000), (8, 8, 1025), {}),                             # ... 5 windows per line
    ((3, 64, 40000), (128, 64, 1000), dict(padding=500)),          # tensor-core contraction over (batch, window) items
    ((2, 128, 66000), (256, 128, 3000), {}),                       # ... with 256 output channels
    ((4, 32, 33000), (32, 32, 64), {}),                            # a line just above a power of two
]


@pytest.mark.parametrize("xs,ws,kw", _BSEG_GPU_CASES)
def test_batch_segments_match_one_transform_plan_and_torch(xs, ws, kw):
    from fft_conv_pytorch_b200 import _lib as L

    g = torch.Generator().manual_seed(23)
    x = torch.randn(*xs, generator=g).cuda()
    w = torch.randn(*ws, generator=g).cuda()
    b = torch.randn(ws[0], generator=g).cuda()
    out = {}
    try:
        for name, flags in (("seg", 0), ("one", L.FC_FLAG_NO_SEGMENT)):
            Fn.set_default_flags(flags)
            Fn.clear_caches()
            with torch.no_grad():
                out[name] = fcp.fft_conv(x, w, b, **kw).clone()
            d = Fn._plans[next(reversed(Fn._plans))].plan.describe()
            assert ("batch segments" in d) == (name == "seg"), d
    finally:
        Fn.set_default_flags(0)
        Fn.clear_caches()
    with torch.no_grad():
        ref = F.conv1d(x.double(), w.double(), b.double(), **kw)
    assert out["seg"].shape == ref.shape
    assert (out["seg"].double() - ref).abs().max().item() / ref.abs().max().item() < TOL
    assert (out["seg"] - out["one"]).abs().max().item() / ref.abs().max().item() < 2e-5


def test_baseline_c4_runs_as_five_windows():
    e = Fn.get_plan(False, 16, 256, 256, 1, (65536,), (4097,), (1,), (0,), (1,), (0,), "constant", 0)
    d = e.plan.describe()
    assert "5 windows of 16384 points" in d and "tc_gemm_3xtf32" in d, d
    assert e.plan.out_size == (61440,)


def test_batch_segments_through_every_call_path():
    """1-d plans that run as batch segments: CPU tensors (one-shot and chunked host pipeline), several GEMM chunks,
    a captured CUDA graph, the module, and the backward pass — all against torch's direct convolution."""
    from fft_conv_pytorch_b200.graphs import GraphedConv

    g = torch.Generator().manual_seed(31)
    # host buffers in and out: B = 1 goes through fc_conv_host, B = 7 through the chunked pipeline (chunk plans share the spectrum)
    for B in (1, 7):
        x = torch.randn(B, 40, 34000, generator=g)
        w = torch.randn(4, 20, 40, generator=g)
        b = torch.randn(4, generator=g)
        with torch.no_grad():
            y = fcp.fft_conv(x, w, b, padding=20, groups=2)
            ref = F.conv1d(x.double(), w.double(), b.double(), padding=20, groups=2)
        assert not y.is_cuda and rel_err(y.numpy(), ref.numpy()) < TOL
    e = Fn.get_plan(False, 7, 40, 4, 2, (34000,), (40,), (1,), (20,), (1,), (0,), "constant")
    assert int(e.plan.info.segments) > 1
    # 8 channels: the small-call rule depends on the batch, so the host pipeline pins the full-batch choice for its chunks
    for B, segs in ((6, 3), (16, 1)):
        x = torch.randn(B, 8, 32768, generator=g)
        w = torch.randn(8, 8, 1025, generator=g)
        with torch.no_grad():
            y = fcp.fft_conv(x, w, None)
            ref = F.conv1d(x.cuda(), w.cuda())  # (cuDNN fp32, TF32 off)
        assert int(Fn.get_plan(False, B, 8, 8, 1, (32768,), (1025,), (1,), (0,), (1,), (0,), "constant").plan.info.segments) == segs
        assert rel_err(y.numpy(), ref.cpu().numpy()) < TOL
    # (batch, window) items beyond one tensor-core GEMM chunk of 80
    x = torch.randn(12, 64, 50000, generator=g).cuda()
    w = torch.randn(128, 64, 1500, generator=g).cuda()
    b = torch.randn(128, generator=g).cuda()
    e = Fn.get_plan(False, 12, 64, 128, 1, (50000,), (1500,), (1,), (0,), (1,), (0,), "constant")
    d = e.plan.describe()
    assert "batch segments" in d and "tc_gemm_3xtf32_b" in d, d
    with torch.no_grad():
        y = fcp.fft_conv(x, w, b)
        ref = F.conv1d(x[:, :, :20000], w, b)  # (cuDNN fp32, TF32 off: conftest) the first outputs of every line
    assert (y[:, :, : ref.shape[-1]].double() - ref.double()).abs().max().item() / ref.double().abs().max().item() < TOL
    # module under a captured graph
    m = fcp.FFTConv1d(24, 8, 65, padding=32).cuda()
    xs = torch.randn(3, 24, 33000, generator=g).cuda()
    gc = GraphedConv(m, xs)
    with torch.no_grad():
        y = gc().clone()
        ref = F.conv1d(xs.double(), m.weight.double(), m.bias.double(), padding=32)
    assert rel_err(y.cpu().numpy(), ref.cpu().numpy()) < TOL
    # backward (the adjoint convolutions plan themselves)
    x0 = torch.randn(2, 20, 20000, generator=g).cuda().requires_grad_()
    w0 = torch.randn(6, 20, 33, generator=g).cuda().requires_grad_()
    b0 = torch.randn(6, generator=g).cuda().requires_grad_()
    x1, w1, b1 = (t.detach().clone().requires_grad_() for t in (x0, w0, b0))
    gy = torch.randn(2, 6, 19968, generator=g).cuda()
    (fcp.fft_conv(x0, w0, bias=b0) * gy).sum().backward()
    (F.conv1d(x1, w1, bias=b1) * gy).sum().backward()
    for a, r in ((w0.grad, w1.grad), (b0.grad, b1.grad), (x0.grad, x1.grad)):
        assert rel_err(a.detach().cpu().numpy(), r.detach().cpu().numpy()) < TOL


@pytest.mark.parametrize("xs,ws,kw", [((32, 32, 8192), (64, 32, 129), {}), ((8, 8, 4000), (8, 8, 65), dict(padding=32)),
                                      ((5, 16, 2048), (16, 8, 33), dict(groups=2, stride=2)), ((3, 128, 4096), (128, 128, 100), {}),
                                      ((2, 6, 3000), (4, 6, 40), dict(padding=20, padding_mode="circular"))])
def test_short_1d_lines_on_the_four_step_layout(xs, ws, kw):
    """1-d lines of 2048 ... 8192 points: column + warp-engine kernels (default) against the generic one-pass kernels."""
    from fft_conv_pytorch_b200 import _lib as L

    g = torch.Generator().manual_seed(41)
    x = torch.randn(*xs, generator=g).cuda()
    w = torch.randn(*ws, generator=g).cuda()
    b = torch.randn(ws[0], generator=g).cuda()
    out = {}
    try:
        for name, flags in (("split", 0), ("one", L.FC_FLAG_NO_SHORT_SPLIT)):
            Fn.set_default_flags(flags)
            Fn.clear_caches()
            with torch.no_grad():
                out[name] = fcp.fft_conv(x, w, b, **kw).clone()
            d = Fn._plans[next(reversed(Fn._plans))].plan.describe()
            assert ("structure=2" in d) == (name == "split"), d
    finally:
        Fn.set_default_flags(0)
        Fn.clear_caches()
    kw2 = dict(kw)
    mode = kw2.pop("padding_mode", "constant")
    xt = x.double()
    if mode != "constant":
        p = kw2.pop("padding")
        xt = F.pad(xt, (p, p), mode=mode)
    with torch.no_grad():
        ref = F.conv1d(xt, w.double(), b.double(), **kw2)
    assert (out["split"].double() - ref).abs().max().item() / ref.abs().max().item() < TOL
    assert (out["split"] - out["one"]).abs().max().item() / ref.abs().max().item() < 2e-5


@pytest.mark.parametrize("xs,ws,kw,tr", [((2, 2, 100, 100, 70), (3, 2, 5, 3, 9), {}, False), ((1, 4, 128, 128, 128), (4, 4, 9, 9, 9), {}, False),
                                         ((1, 3, 40, 100, 90), (2, 3, 3, 3, 5), dict(padding=(1, 2, 2)), False),
                                         ((1, 2, 60, 60, 30), (2, 2, 3, 3, 3), dict(stride=2, padding=1, output_padding=1), True)])
def test_3d_programs_with_128_point_plane_axes(xs, ws, kw, tr):
    g = torch.Generator().manual_seed(43)
    x = torch.randn(*xs, generator=g).cuda()
    w = torch.randn(*ws, generator=g).cuda()
    b = torch.randn(ws[1] if tr else ws[0], generator=g).cuda()
    with torch.no_grad():
        y = (fcp.fft_conv_transpose if tr else fcp.fft_conv)(x, w, b, **kw)
        ref = (F.conv_transpose3d if tr else F.conv3d)(x.double(), w.double(), b.double(), **kw)
    d = Fn._plans[next(reversed(Fn._plans))].plan.describe()
    assert ("plane_inv_128" in d if tr else "plane_fwd_" in d and "128" in d), d  # (a zero-stuffing gather keeps the forward passes apart)
    assert (y.double() - ref).abs().max().item() / ref.abs().max().item() < TOL


@pytest.mark.parametrize("seed", range(48))
def test_random_1d_problems_across_the_round_2_paths(seed):
    """Random 1-d forward problems over the ranges where the plan chooses between the line kernels (512 / 1024 points), the
    four-step layout from 2048 points, batch segments (windows as batch items), the tiled / register-tile / tensor-core
    contractions: lengths 300 ... 120000, 1 ... 96 channels per group, batches up to 40, every padding mode."""
    rng = np.random.RandomState(5000 + seed)
    groups = int(rng.choice([1, 1, 2, 3]))
    ig = int(rng.choice([1, 3, 8, 16, 20, 48, 64, 96]))
    og = int(rng.choice([1, 4, 8, 13, 32, 64, 128]))
    cin, cout = ig * groups, og * groups
    k = int(rng.choice([1, 2, 9, 33, 100, 513, 2049]))
    dil = int(rng.choice([1, 1, 2, 3]))
    stride = int(rng.choice([1, 1, 1, 2, 3]))
    L_ = int(rng.choice([300, 500, 1000, 2000, 4000, 8100, 16500, 33000, 70000, 120000])) + int(rng.randint(0, 40))
    L_ = max(L_, (k - 1) * dil + 1)
    B = int(rng.choice([1, 2, 5, 9, 40]))
    while B * (cin + cout) * L_ > 6e7:  # keep a case under ~0.25 GB
        B = max(1, B // 2)
        if B == 1:
            break
    if B * (cin + cout) * L_ > 6e7:
        L_ = max(int(6e7 / (cin + cout)), (k - 1) * dil + 1)
    mode = str(rng.choice(["constant", "constant", "reflect", "replicate", "circular"]))
    pad = int(rng.randint(0, min(40, L_ - 1)))
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, cin, L_, device="cuda", generator=g)
    w = torch.randn(cout, ig, k, device="cuda", generator=g)
    b = torch.randn(cout, device="cuda", generator=g)
    with torch.no_grad():
        y = fcp.fft_conv(x, w, b, stride=stride, padding=pad, dilation=dil, groups=groups, padding_mode=mode)
        xp = x.double() if mode == "constant" or pad == 0 else F.pad(x.double(), (pad, pad), mode=mode)
        ref = F.conv1d(xp, w.double(), b.double(), stride=stride, padding=pad if mode == "constant" else 0, dilation=dil, groups=groups)
    assert y.shape == ref.shape, (B, cin, cout, L_, k, stride, pad, dil, groups, mode)
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < TOL, (err, B, cin, cout, L_, k, stride, pad, dil, groups, mode)
