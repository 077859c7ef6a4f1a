#!/usr/bin/env python3
"""Baseline probe: time and profile the REFERENCE fft_conv_pytorch on this box.

Measurement only. It imports the unmodified reference package from
baseline/_ref/ (a copy of /root/reference/fft_conv_pytorch, git-ignored) and
reports, per BASELINE.json config:
  * CUDA-event wall time of the reference's fft_conv / fft_conv_transpose,
  * peak allocated memory,
  * the CUDA kernels it launches (torch profiler, grouped by kernel name),
  * its error against fp64 fft_conv and against direct F.conv (TF32 off/on),
  * the same call on the host CPU cores of the same box.
Results go to gpurun_out/ref_probe.json and gpurun_out/ref_probe.txt.
"""
import json
import os
import sys
import time
import traceback
import warnings

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref"))

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from fft_conv_pytorch.functional import fft_conv, fft_conv_transpose  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "gpurun_out")
os.makedirs(OUT, exist_ok=True)
LOG = open(os.path.join(OUT, "ref_probe.txt"), "w")
RESULTS = {"env": {}, "configs": []}


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def save():
    with open(os.path.join(OUT, "ref_probe.json"), "w") as f:
        json.dump(RESULTS, f, indent=1)


def make(cfg, device, dtype=torch.float32, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(*cfg["x"], generator=g, dtype=torch.float32)
    w = torch.randn(*cfg["w"], generator=g, dtype=torch.float32)
    b = torch.randn(cfg["cout"], generator=g, dtype=torch.float32)
    return x.to(device, dtype), w.to(device, dtype), b.to(device, dtype)


def ref_call(cfg, x, w, b):
    if cfg["transpose"]:
        return fft_conv_transpose(x, w, b, **cfg["kw"])
    return fft_conv(x, w, b, **cfg["kw"])


def direct_call(cfg, x, w, b):
    n = x.ndim - 2
    fn = getattr(F, ("conv_transpose%dd" if cfg["transpose"] else "conv%dd") % n)
    return fn(x, w, b, **cfg["kw"])


def cuda_time(fn, iters, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    return {"best_s": ts[0], "median_s": ts[len(ts) // 2], "iters": iters}


def cpu_time(fn, iters, warm):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    ts.sort()
    return {"best_s": ts[0], "median_s": ts[len(ts) // 2], "iters": iters}


def err(a, ref):
    a = a.double()
    ref = ref.double()
    e = (a - ref).abs()
    return {
        "max_abs": e.max().item(),
        "mean_abs": e.mean().item(),
        "max_over_refmax": (e.max() / ref.abs().max()).item(),
        "max_over_refrms": (e.max() / ref.pow(2).mean().sqrt()).item(),
    }


def kernel_table(fn):
    from torch.profiler import ProfilerActivity, profile

    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as p:
        fn()
        torch.cuda.synchronize()
    rows = []
    for e in p.key_averages():
        dt = getattr(e, "self_device_time_total", None)
        if dt is None:
            dt = getattr(e, "self_cuda_time_total", 0)
        if dt and dt > 0 and not e.key.startswith("aten::"):
            rows.append({"kernel": e.key[:160], "us": dt, "calls": e.count})
    rows.sort(key=lambda r: -r["us"])
    return rows[:25]


CONFIGS = [
    dict(name="cfg1_1d", x=(1, 8, 32768), w=(8, 8, 1025), cout=8, transpose=False, kw={}, cpu=True, f64=True, direct=True),
    dict(name="cfg2_2d", x=(8, 8, 512, 512), w=(8, 8, 65, 65), cout=8, transpose=False, kw={}, cpu=True, f64=True, direct=True),
    dict(name="cfg3_3d", x=(4, 8, 64, 64, 64), w=(8, 8, 17, 17, 17), cout=8, transpose=False, kw={}, cpu=True, f64=True, direct=True),
    dict(name="cfg4_1d_wide", x=(16, 256, 65536), w=(256, 256, 4097), cout=256, transpose=False, kw={}, cpu=False, f64=False, direct=True),
    dict(name="cfg5_2d_transpose_shard_b4", x=(4, 64, 1024, 1024), w=(64, 16, 31, 31), cout=64, transpose=True,
         kw=dict(stride=2, dilation=2, groups=4), cpu=False, f64=False, direct=True),
    dict(name="cfg5_2d_transpose_b1", x=(1, 64, 1024, 1024), w=(64, 16, 31, 31), cout=64, transpose=True,
         kw=dict(stride=2, dilation=2, groups=4), cpu=False, f64=True, direct=False),
    dict(name="cfg5_2d_transpose_full_b32", x=(32, 64, 1024, 1024), w=(64, 16, 31, 31), cout=64, transpose=True,
         kw=dict(stride=2, dilation=2, groups=4), cpu=False, f64=False, direct=False),
]


def run_config(cfg):
    r = {"name": cfg["name"], "x": cfg["x"], "w": cfg["w"], "kw": cfg["kw"], "transpose": cfg["transpose"]}
    RESULTS["configs"].append(r)
    log("=" * 100)
    log(cfg["name"], cfg["x"], cfg["w"], cfg["kw"])
    dev = "cuda"
    x, w, b = make(cfg, dev)
    torch.cuda.synchronize()
    with torch.no_grad():
        try:
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats()
            base = torch.cuda.memory_allocated()
            y = ref_call(cfg, x, w, b)
            torch.cuda.synchronize()
            r["out_shape"] = list(y.shape)
            r["out_samples"] = y.numel()
            r["peak_alloc_gib"] = torch.cuda.max_memory_allocated() / 2**30
            r["inputs_gib"] = base / 2**30
            big = y.numel() > 2e9
            t = cuda_time(lambda: ref_call(cfg, x, w, b), iters=3 if big else 10, warm=1 if big else 3)
            r["ref_gpu"] = t
            r["ref_gpu_gsamples_per_s"] = y.numel() / t["best_s"] / 1e9
            log("  ref GPU: best %.3f ms median %.3f ms  -> %.3f Gsamples/s, out %s, peak alloc %.2f GiB (inputs %.2f GiB)"
                % (t["best_s"] * 1e3, t["median_s"] * 1e3, r["ref_gpu_gsamples_per_s"], tuple(y.shape), r["peak_alloc_gib"], r["inputs_gib"]))
            save()
        except torch.cuda.OutOfMemoryError as e:
            r["ref_gpu"] = "OOM"
            r["peak_alloc_gib"] = torch.cuda.max_memory_allocated() / 2**30
            log("  ref GPU: OOM (peak alloc before failure %.1f GiB): %s" % (r["peak_alloc_gib"], str(e)[:200]))
            save()
            return
        # kernel breakdown
        try:
            rows = kernel_table(lambda: ref_call(cfg, x, w, b))
            r["kernels"] = rows
            tot = sum(k["us"] for k in rows) or 1
            for k in rows[:14]:
                log("    %9.1f us %5.1f%% x%-3d %s" % (k["us"], 100 * k["us"] / tot, k["calls"], k["kernel"][:110]))
        except Exception as e:  # noqa: BLE001
            log("  profiler failed:", repr(e)[:200])
        save()
        # accuracy
        try:
            if cfg["f64"]:
                y64 = ref_call(cfg, x.double(), w.double(), b.double())
                r["err_ref_fp32_vs_ref_fp64"] = err(y, y64)
                log("  ref fp32 vs ref fp64:", r["err_ref_fp32_vs_ref_fp64"])
            else:
                y64 = None
            if cfg["direct"]:
                for tf32 in (False, True):
                    torch.backends.cudnn.allow_tf32 = tf32
                    torch.backends.cuda.matmul.allow_tf32 = tf32
                    t0 = time.perf_counter()
                    yd = direct_call(cfg, x, w, b)
                    torch.cuda.synchronize()
                    first = time.perf_counter() - t0
                    td = cuda_time(lambda: direct_call(cfg, x, w, b), iters=2, warm=0) if first < 20 else {"best_s": first, "median_s": first, "iters": 1}
                    key = "direct_tf32_%s" % ("on" if tf32 else "off")
                    r[key] = {"time": td, "err_ref_fft_vs_direct": err(y, yd)}
                    if y64 is not None:
                        r[key]["err_direct_vs_fp64"] = err(yd, y64)
                    log("  direct F.conv TF32=%s: best %.3f ms; ref-fft vs direct: %s%s"
                        % (tf32, td["best_s"] * 1e3, r[key]["err_ref_fft_vs_direct"],
                           ("; direct vs fp64: %s" % r[key]["err_direct_vs_fp64"]) if y64 is not None else ""))
                    del yd
                torch.backends.cudnn.allow_tf32 = False
                torch.backends.cuda.matmul.allow_tf32 = False
            del y64
        except Exception as e:  # noqa: BLE001
            log("  accuracy section failed:", repr(e)[:300])
        save()
        del y
        torch.cuda.empty_cache()
        # host CPU cores of this same box
        if cfg["cpu"]:
            xc, wc, bc = x.cpu(), w.cpu(), b.cpu()
            tc = cpu_time(lambda: ref_call(cfg, xc, wc, bc), iters=5, warm=2)
            r["ref_cpu"] = tc
            r["ref_cpu_gsamples_per_s"] = r["out_samples"] / tc["best_s"] / 1e9
            log("  ref CPU (%d threads): best %.3f ms median %.3f ms -> %.5f Gsamples/s"
                % (torch.get_num_threads(), tc["best_s"] * 1e3, tc["median_s"] * 1e3, r["ref_cpu_gsamples_per_s"]))
        save()


def main():
    RESULTS["env"] = {
        "torch": torch.__version__,
        "cuda": torch.version.cuda,
        "gpu": torch.cuda.get_device_name(0) if torch.cuda.is_available() else None,
        "n_gpu": torch.cuda.device_count(),
        "cpu_count": os.cpu_count(),
        "torch_threads": torch.get_num_threads(),
        "cufft_plan_cache_max": torch.backends.cuda.cufft_plan_cache.max_size if torch.cuda.is_available() else None,
        "cudnn_allow_tf32_default": torch.backends.cudnn.allow_tf32,
        "matmul_allow_tf32_default": torch.backends.cuda.matmul.allow_tf32,
    }
    try:
        with open("/proc/cpuinfo") as f:
            models = [line.split(":", 1)[1].strip() for line in f if line.startswith("model name")]
        RESULTS["env"]["cpu_model"] = models[0] if models else None
    except OSError:
        pass
    log(json.dumps(RESULTS["env"]))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    only = sys.argv[1:]
    for cfg in CONFIGS:
        if only and cfg["name"] not in only:
            continue
        try:
            run_config(cfg)
        except Exception:  # noqa: BLE001
            log("  CONFIG FAILED:", traceback.format_exc()[-800:])
            torch.cuda.empty_cache()
        save()
    log("done")


if __name__ == "__main__":
    main()
