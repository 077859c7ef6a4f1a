#!/usr/bin/env python3
"""Device time of every BASELINE.json config with this library (CUDA events, cached kernel spectrum, L2 flushed
between calls) and, where it fits, of the unmodified reference on the same GPU (cuFFT + cuBLAS through torch)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import torch

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import functional as Fn

CFG = {
    "c1": dict(x=(1, 8, 32768), w=(8, 8, 1025), tr=False, kw={}),
    "c2": dict(x=(8, 8, 512, 512), w=(8, 8, 65, 65), tr=False, kw={}),
    "c3": dict(x=(4, 8, 64, 64, 64), w=(8, 8, 17, 17, 17), tr=False, kw={}),
    "c4": dict(x=(16, 256, 65536), w=(256, 256, 4097), tr=False, kw={}),
    # not BASELINE configs: common image sizes, to watch the short-line kernels
    "img128": dict(x=(32, 8, 128, 128), w=(8, 8, 15, 15), tr=False, kw={}),
    "img256": dict(x=(16, 8, 256, 256), w=(8, 8, 31, 31), tr=False, kw={}),
    "c5_shard": dict(x=(4, 64, 1024, 1024), w=(64, 16, 31, 31), tr=True, kw=dict(stride=2, dilation=2, groups=4)),
    # the whole c5 batch on one GPU (8.6 GB in, 36.4 GB out): fits since the overlap-save program needs no full-size spectra
    "c5_full": dict(x=(32, 64, 1024, 1024), w=(64, 16, 31, 31), tr=True, kw=dict(stride=2, dilation=2, groups=4)),
}


def timeit(fn, n, flush):
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def main():
    argv = sys.argv[1:]
    tag = ""
    if argv and argv[0].startswith("--flags="):  # FC_FLAG_* bits OR-ed into every plan (A/B runs), e.g. --flags=256: no batch pairs
        Fn.set_default_flags(int(argv[0].split("=")[1]))
        tag = "_flags" + argv[0].split("=")[1]
        argv = argv[1:]
    only = argv or [k for k in CFG if k != "c5_full"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {}
    for name in only:
        c = CFG[name]
        g = torch.Generator().manual_seed(0)
        x = torch.randn(*c["x"], generator=g).cuda()
        w = torch.randn(*c["w"], generator=g).cuda()
        cout = c["w"][1] * c["kw"].get("groups", 1) if c["tr"] else c["w"][0]
        b = torch.randn(cout, generator=g).cuda()
        fn = fcp.fft_conv_transpose if c["tr"] else fcp.fft_conv
        with torch.no_grad():
            t0 = time.perf_counter()
            y = fn(x, w, b, **c["kw"])
            torch.cuda.synchronize()
            cold = time.perf_counter() - t0
            for _ in range(2):
                fn(x, w, b, **c["kw"])
            best, med = timeit(lambda: fn(x, w, b, **c["kw"]), 7, flush)
        n = y.numel()
        r = dict(out_shape=list(y.shape), ours_ms_best=best, ours_ms_median=med, ours_gsamples=n / best / 1e6, first_call_s=cold,
                 peak_gib=torch.cuda.max_memory_allocated() / 2**30)
        del y
        # per-kernel breakdown
        try:
            import ctypes
            from fft_conv_pytorch_b200 import _lib as L
            nd = len(c["x"]) - 2
            tup = lambda v, d: tuple(v) if hasattr(v, "__iter__") else (v,) * nd
            kw = c["kw"]
            entry = Fn.get_plan(c["tr"], c["x"][0], c["x"][1], cout, kw.get("groups", 1), tuple(c["x"][2:]), tuple(c["w"][2:]),
                                tup(kw.get("stride", 1), nd), tup(kw.get("padding", 0), nd), tup(kw.get("dilation", 1), nd),
                                tup(kw.get("output_padding", 0), nd), "constant")
            plan = entry.plan
            lib = plan.lib
            kspec = Fn.kernel_spectrum(entry, w, x.device)
            const = entry.const_for(x.device)
            ws = torch.empty(int(plan.info.workspace_bytes), dtype=torch.uint8, device="cuda")
            yb = torch.empty((c["x"][0], cout) + plan.out_size, device="cuda")
            nl = int(plan.info.n_launches)
            ms = (ctypes.c_float * nl)()
            n_out = ctypes.c_int(0)
            P = lambda t: ctypes.c_void_p(t.data_ptr())
            st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            acc = [0.0] * nl
            reps = 3
            for _ in range(reps):
                flush.zero_()
                L.check(lib, lib.fc_conv_profiled(plan.handle, P(const), P(x), P(kspec), P(b), P(yb), P(ws), st, ms, nl, ctypes.byref(n_out)), "prof")
                for j in range(nl):
                    acc[j] += ms[j] / reps
            ks = []
            for j in range(nl):
                nm = ctypes.create_string_buffer(64)
                ab = ctypes.c_int64(0)
                lib.fc_plan_launch_info(plan.handle, j, nm, 64, ctypes.byref(ab))
                ks.append(dict(kernel=nm.value.decode(), ms=acc[j], gbs=ab.value / acc[j] / 1e6 if acc[j] > 0 else None))
            r["kernels"] = ks
            r["fft_size"] = list(plan.fft_size)
            r["workspace_gib"] = plan.info.workspace_bytes / 2**30
            r["kspec_gib"] = plan.info.kspec_bytes / 2**30
            del ws, yb, kspec
        except Exception as e:  # noqa: BLE001
            r["kernels_error"] = repr(e)[:200]
        Fn.clear_caches()
        torch.cuda.empty_cache()
        if os.environ.get("FFTCONV_SKIP_REF"):
            out[name] = r
            print(name, json.dumps(r), flush=True)
            continue
        # reference on the same GPU
        try:
            from fft_conv_pytorch.functional import fft_conv as rf, fft_conv_transpose as rft
            import warnings
            warnings.filterwarnings("ignore")
            rfn = rft if c["tr"] else rf
            with torch.no_grad():
                rfn(x, w, b, **c["kw"])
                torch.cuda.synchronize()
                best, med = timeit(lambda: rfn(x, w, b, **c["kw"]), 3, flush)
            r["ref_gpu_ms_best"] = best
            r["speedup_vs_ref_gpu"] = best / r["ours_ms_best"]
        except Exception as e:  # noqa: BLE001
            r["ref_gpu_error"] = repr(e)[:160]
        torch.cuda.empty_cache()
        out[name] = r
        print(name, json.dumps(r), flush=True)
        del x, w, b
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"time_configs{tag}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
