#!/usr/bin/env python3
"""Kernel-size sweep on one B200: time and peak memory of this library's fft_conv / fft_conv_transpose
against torch's direct convolutions (cuDNN, TF32 off) and the unmodified reference's fft_conv on the same GPU, at the
three sweeps of the reference's README figure (reference doc/scripts/generate_benchmark_plot.py:125-160: 1-d 32768,
2-d 512^2, 3-d 64^3; batch 2, 8 -> 8 channels). matplotlib is not in this image, so the result is a JSON file and a
markdown table (gpurun_out/kernel_size_sweep.{json,md}) instead of a PNG.

Device time: CUDA events, best and median of `iters` calls after 2 warm-ups; ours with the kernel spectrum cached
(steady state) and, in `ours_cold_ms`, including the kernel-spectrum transform. Memory: peak allocated bytes above the
inputs during one call."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torch
import torch.nn.functional as F

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import functional as Fn

SWEEPS = [
    dict(ndim=1, input_size=32768, kernel_sizes=[1] + list(range(256, 4096, 512))),
    dict(ndim=2, input_size=512, kernel_sizes=[1] + list(range(4, 49, 6))),
    dict(ndim=3, input_size=64, kernel_sizes=[1, 2, 4, 6, 8]),
]
B, CIN, COUT = 2, 8, 8


def timed(fn, iters):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def peak_mem(fn):
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    base = torch.cuda.memory_allocated()
    torch.cuda.reset_peak_memory_stats()
    y = fn()
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    del y
    return peak


def main():
    iters = int(os.environ.get("SWEEP_ITERS", "16"))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        from fft_conv_pytorch import functional as ref  # the unmodified reference, if it travelled with the snapshot
    except Exception:
        ref = None
    rows = []
    for sw in SWEEPS:
        nd = sw["ndim"]
        g = torch.Generator().manual_seed(0)
        x = torch.randn(B, CIN, *([sw["input_size"]] * nd), generator=g).cuda()
        for k in sw["kernel_sizes"]:
            w = torch.randn(COUT, CIN, *([k] * nd), generator=g).cuda()
            wt = torch.randn(CIN, COUT, *([k] * nd), generator=g).cuda()
            b = torch.randn(COUT, generator=g).cuda()
            direct = getattr(F, f"conv{nd}d")
            direct_t = getattr(F, f"conv_transpose{nd}d")
            row = dict(ndim=nd, input_size=sw["input_size"], kernel_size=k)
            with torch.no_grad():
                row["ours_ms"], row["ours_ms_median"] = timed(lambda: fcp.fft_conv(x, w, b), iters)
                row["ours_mem_mb"] = peak_mem(lambda: fcp.fft_conv(x, w, b)) / 1e6

                def cold():
                    Fn.clear_caches(plans=False)
                    return fcp.fft_conv(x, w, b)

                row["ours_cold_ms"], _ = timed(cold, max(4, iters // 4))
                row["ours_transpose_ms"], _ = timed(lambda: fcp.fft_conv_transpose(x, wt, b), iters)
                try:
                    row["direct_ms"], row["direct_ms_median"] = timed(lambda: direct(x, w, b), iters)
                    row["direct_mem_mb"] = peak_mem(lambda: direct(x, w, b)) / 1e6
                    row["direct_transpose_ms"], _ = timed(lambda: direct_t(x, wt, b), iters)
                    err = (fcp.fft_conv(x, w, b) - direct(x, w, b)).abs().max() / direct(x, w, b).abs().max()
                    row["rel_err_vs_direct"] = float(err)
                except Exception as e:  # cuDNN has no algorithm / runs out of memory for some large kernels
                    row["direct_error"] = str(e)[:120]
                if ref is not None:
                    try:
                        row["reference_gpu_ms"], _ = timed(lambda: ref.fft_conv(x, w, b), iters)
                        row["reference_gpu_mem_mb"] = peak_mem(lambda: ref.fft_conv(x, w, b)) / 1e6
                    except Exception as e:
                        row["reference_gpu_error"] = str(e)[:120]
            # forward + backward (the reference's own benchmark creates its inputs with requires_grad=True,
            # generate_benchmark_plot.py:27): ours through autograd.py, torch's direct convolution, the reference on the GPU
            def fwd_bwd(fn):
                xg = x.detach().requires_grad_(True)
                wg = w.detach().requires_grad_(True)
                bg = b.detach().requires_grad_(True)

                def run():
                    xg.grad = wg.grad = bg.grad = None
                    fn(xg, wg, bg).sum().backward()

                return run

            try:
                row["ours_fwd_bwd_ms"], _ = timed(fwd_bwd(fcp.fft_conv), max(4, iters // 2))
                row["direct_fwd_bwd_ms"], _ = timed(fwd_bwd(direct), max(4, iters // 2))
                if ref is not None:
                    row["reference_gpu_fwd_bwd_ms"], _ = timed(fwd_bwd(ref.fft_conv), max(4, iters // 2))
            except Exception as e:  # noqa: BLE001
                row["fwd_bwd_error"] = str(e)[:120]
            rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "kernel_size_sweep.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "kernel_size_sweep.md"), "w") as f:
        f.write("| ndim | input | kernel | ours ms (cached spectrum) | ours ms (cold) | ours transposed ms | direct (cuDNN) ms | direct transposed ms | "
                "reference fft_conv on GPU ms | ours MB | direct MB | reference MB | rel err vs direct | ours fwd+bwd ms | direct fwd+bwd ms | "
                "reference fwd+bwd ms |\n|" + "---|" * 16 + "\n")
        fmt = lambda v, p=3: "—" if v is None else f"{v:.{p}f}"
        for r in rows:
            f.write(f"| {r['ndim']} | {r['input_size']} | {r['kernel_size']} | {fmt(r.get('ours_ms'))} | {fmt(r.get('ours_cold_ms'))} | "
                    f"{fmt(r.get('ours_transpose_ms'))} | {fmt(r.get('direct_ms'))} | {fmt(r.get('direct_transpose_ms'))} | {fmt(r.get('reference_gpu_ms'))} | "
                    f"{fmt(r.get('ours_mem_mb'), 1)} | {fmt(r.get('direct_mem_mb'), 1)} | {fmt(r.get('reference_gpu_mem_mb'), 1)} | "
                    f"{r.get('rel_err_vs_direct', float('nan')):.1e} | {fmt(r.get('ours_fwd_bwd_ms'))} | {fmt(r.get('direct_fwd_bwd_ms'))} | "
                    f"{fmt(r.get('reference_gpu_fwd_bwd_ms'))} |\n")


if __name__ == "__main__":
    main()
