#!/usr/bin/env python3
"""Time the tensor-core contraction alone at the BASELINE c4 channel shape (16 x 256 -> 256) on `bins` frequency bins."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fft_conv_pytorch_b200 import _lib as L

bins = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
B, cin, cout, groups = 16, 256, 256, 1
lib = L.load()
torch.manual_seed(0)
a = torch.randn(B, cin, bins, dtype=torch.complex64, device="cuda")
b = torch.randn(cout, cin, bins, dtype=torch.complex64, device="cuda")
btc = torch.empty_like(torch.view_as_real(b))
y = torch.empty(B, cout, bins, dtype=torch.complex64, device="cuda")
scratch = torch.empty(int(lib.fc_tc_scratch_bytes(B, cin, cout, groups, bins)), dtype=torch.uint8, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
L.check(lib, lib.fc_tc_prepare_kernel(P(torch.view_as_real(b)), P(btc), cin, cout, groups, bins, st), "prepare")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(6):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib, lib.fc_tc_complex_matmul(P(torch.view_as_real(a)), P(btc), P(torch.view_as_real(y)), P(scratch), B, cin, cout, groups, bins, st), "tc")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ref = torch.einsum("bif,oif->bof", a[:, :, :64].to(torch.complex128), b[:, :, :64].to(torch.complex128))
err = (y[:, :, :64].to(torch.complex128) - ref).abs().max().item() / ref.abs().max().item()
flops = 8.0 * B * cin * cout * bins
print(f"bins {bins}: relayout+gemm+relayout best {min(ts[1:]):.3f} ms; rel err {err:.2e}; "
      f"{flops / min(ts[1:]) / 1e9:.1f} fp32-equivalent TFLOP/s ({3 * flops / min(ts[1:]) / 1e9:.1f} issued as TF32)")
