#!/usr/bin/env python3
"""Benchmark of the fft_conv hot path (BASELINE.json metric: output Gsamples/s; achieved HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1..c5] [--quick]

Headline (`value`, `e2e`, `roofline`): BASELINE configs[1] (c2: FFTConv2d fp32, input (8,8,512,512), 8->8 channels,
65x65 kernel, bias), kernel spectrum cached (steady state), one step = one pass of the hot path over one
batch of synthetic input. With N > 1 (torchrun, one rank per GPU) every rank convolves its own batch of the same
shape with weights broadcast once from rank 0 (weak scaling of independent replicas, no data-path collective);
`value` is the whole-job throughput: samples of all ranks / max-over-ranks device time.

Beside the headline the one JSON line carries
  configs        every BASELINE shape (c1, c2, c3, c4, c5 per-GPU shard B=4): ms, Gsamples/s, fraction of the HBM roofline
                 of the whole pipeline (SURVEY A_pipe) and of the dominant kernel, and the unmodified reference on the
                 same GPU (torch.fft = cuFFT, einsum = cuBLAS) where it runs
  gpu_reference  the unmodified reference (baseline/_ref) on the same GPU at the headline config
  strong         strong scaling of the sharded configs: BASELINE c5 with its global batch 32 split B/N per
                 rank (N = 1 runs the whole batch on one GPU), c2 and c3 split by batch; device time = max over ranks
  cpu_baseline   the unmodified reference on the host cores (N = 1 only)

`--impl reference` times the reference's own CPU implementation of the path (baseline/_ref when present, else the
oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c1": dict(nd=1, x=(1, 8, 32768), w=(8, 8, 1025), tr=False, kw={}, desc="1D fft_conv fp32 (1,8,32768) k1025"),
    "c2": dict(nd=2, x=(8, 8, 512, 512), w=(8, 8, 65, 65), tr=False, kw={}, desc="2D FFTConv2d fp32 (8,8,512,512) 8->8 k65x65 bias"),
    "c3": dict(nd=3, x=(4, 8, 64, 64, 64), w=(8, 8, 17, 17, 17), tr=False, kw={}, desc="3D FFTConv3d fp32 (4,8,64,64,64) 8->8 k17^3 bias"),
    "c4": dict(nd=1, x=(16, 256, 65536), w=(256, 256, 4097), tr=False, kw={}, desc="1D wide-channel fp32 (16,256,65536) 256->256 k4097"),
    # BASELINE c5 is (32,64,1024,1024) sharded over 8 GPUs: the per-GPU shard is B = 4; `strong` runs the global batch
    "c5": dict(nd=2, x=(4, 64, 1024, 1024), w=(64, 16, 31, 31), tr=True, kw=dict(stride=2, dilation=2, groups=4),
               desc="2D fft_conv_transpose fp32 per-GPU shard (4,64,1024,1024) 64->64 k31x31 s2 d2 g4 (global batch 32 over 8 GPUs)"),
}
C5_GLOBAL_BATCH = 32


def cout_of(cfg):
    return cfg["w"][1] * cfg["kw"].get("groups", 1) if cfg["tr"] else cfg["w"][0]


def out_shape(cfg, batch=None):
    """Output shape by the reference's shape algebra (functional.py:79 / :144-154), zero padding."""
    kw = cfg["kw"]
    s, d = kw.get("stride", 1), kw.get("dilation", 1)
    sp = []
    for L, K in zip(cfg["x"][2:], cfg["w"][2:]):
        sp.append((L - 1) * s + d * (K - 1) + 1 if cfg["tr"] else (L - d * (K - 1) - 1) // s + 1)
    return (cfg["x"][0] if batch is None else batch, cout_of(cfg)) + tuple(sp)


def out_samples(cfg, batch=None):
    n = 1
    for v in out_shape(cfg, batch):
        n *= v
    return n


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_fn(cfg):
    """Returns (callable(batch) -> seconds per call, kind, cores). Reference from baseline/_ref if it travelled,
    else the oracle port (numpy/scipy pocketfft with all cores)."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    w = torch.randn(*cfg["w"], generator=g)
    b = torch.randn(cout_of(cfg), generator=g)
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    kind = "port"
    fn = None
    if os.path.isdir(os.path.join(ref_dir, "fft_conv_pytorch")):
        try:
            sys.path.insert(0, ref_dir)
            import warnings

            warnings.filterwarnings("ignore")
            from fft_conv_pytorch.functional import fft_conv as ref_fft_conv, fft_conv_transpose as ref_fft_conv_transpose  # unmodified

            rf = ref_fft_conv_transpose if cfg["tr"] else ref_fft_conv

            def fn(x):
                with torch.no_grad():
                    return rf(x, w, b, **cfg["kw"])

            kind = "reference"
        except Exception:
            fn = None
    if fn is None:
        from oracle import fftconv_oracle as O

        wn, bn = w.numpy(), b.numpy()
        of = O.fft_conv_transpose if cfg["tr"] else O.fft_conv

        def fn(x):
            return of(x.numpy(), wn, bn, **cfg["kw"])

    def run(batch):
        x = torch.randn(batch, *cfg["x"][1:], generator=g)
        t0 = time.perf_counter()
        fn(x)
        return time.perf_counter() - t0

    return run, kind, cores


def bench_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, kind, cores = cpu_reference_fn(cfg)
    full_b = cfg["x"][0]
    t1 = run(1)  # probe cost of one sample
    t1 = min(t1, run(1))
    budget = 150.0
    batch = max(1, min(full_b, int(budget / max(args.steps + args.warmup, 1) / max(t1, 1e-6))))
    for _ in range(args.warmup):
        run(batch)
    times = [run(batch) for _ in range(args.steps)]
    per_sample = out_samples(cfg) / full_b
    total = sum(times)
    value = per_sample * batch * args.steps / total / 1e9
    line = {
        "impl": "reference", "metric": "fft_conv output Gsamples/s", "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "batch_per_step": batch},
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} calls of the full kernel/channel shape at batch {batch} of {full_b}"},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [int(s[0]) for s in self.samples if s[0].isdigit()]
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- our arm (GPU)
class Ctx:
    """Process-wide state of one bench run (device, ranks, L2 flush buffer, measured peak)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.flush_rd = torch.zeros(256 << 18, dtype=torch.int32, device=self.dev) if args.flush == "write+read" else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        self.peak = float(peaks.get("hbm_gbs", 6650.0))
        self.peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    def flush_l2(self):
        """Evict everything of ours from L2 before a timed step: write 256 MiB; optionally also read a second 256 MiB
        buffer so that the dirty lines of the memset are written back before the timed region starts (measured on
        B200: no difference)."""
        self.flush.zero_()
        if self.flush_rd is not None:
            self.flush_rd.sum()

    def barrier(self):
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, vals):
        import torch
        import torch.distributed as dist

        if self.world == 1:
            return list(vals)
        t = torch.tensor(list(vals), device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()


def plan_entry(cfg, batch):
    from fft_conv_pytorch_b200 import functional as Fn

    nd, kw = cfg["nd"], cfg["kw"]
    tup = lambda v: tuple(v) if hasattr(v, "__iter__") else (v,) * nd
    return Fn.get_plan(cfg["tr"], batch, cfg["x"][1], cout_of(cfg), kw.get("groups", 1), tuple(cfg["x"][2:]), tuple(cfg["w"][2:]),
                       tup(kw.get("stride", 1)), tup(kw.get("padding", 0)), tup(kw.get("dilation", 1)), tup(kw.get("output_padding", 0)), "constant")


def kernel_breakdown(ctx, cfg, x, w, b, reps):
    """CUDA events around every launch of one call (fc_conv_profiled), L2 flushed before each call."""
    import torch

    from fft_conv_pytorch_b200 import _lib as L
    from fft_conv_pytorch_b200 import functional as Fn

    entry = plan_entry(cfg, x.shape[0])
    plan = entry.plan
    lib = plan.lib
    kspec = Fn.kernel_spectrum(entry, w, ctx.dev)
    const = entry.const_for(ctx.dev)
    ws = torch.empty(int(plan.info.workspace_bytes), dtype=torch.uint8, device=ctx.dev)
    yb = torch.empty((x.shape[0], cout_of(cfg)) + plan.out_size, device=ctx.dev)
    nl = int(plan.info.n_launches)
    acc = [0.0] * nl
    ms = (ctypes.c_float * nl)()
    n_out = ctypes.c_int(0)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream(ctx.dev).cuda_stream)
    for _ in range(reps):
        ctx.flush_l2()
        L.check(lib, lib.fc_conv_profiled(plan.handle, P(const), P(x), P(kspec), P(b), P(yb), P(ws), stream, ms, nl, ctypes.byref(n_out)),
                "fc_conv_profiled")
        for j in range(n_out.value):
            acc[j] += ms[j]
    out = []
    for j in range(nl):
        name = ctypes.create_string_buffer(64)
        ab = ctypes.c_int64(0)
        lib.fc_plan_launch_info(plan.handle, j, name, 64, ctypes.byref(ab))
        out.append({"kernel": name.value.decode(), "ms": acc[j] / reps, "algo_bytes": ab.value})
    info = plan.info
    a_pipe = int(info.algo_bytes_s1 + info.algo_bytes_s3 + info.algo_bytes_s4)
    del ws, yb
    return out, a_pipe, info


def gpu_reference_ms(ctx, cfg, x, w, b, steps):
    """The unmodified reference (baseline/_ref) on the same GPU: torch.fft (cuFFT) + einsum (cuBLAS). None if it is not
    there or cannot run the shape (cuFFT rejects c5's extent 2168; c4/c5 may not fit)."""
    import torch

    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "fft_conv_pytorch")):
        return None, "baseline/_ref absent"
    try:
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        import warnings

        warnings.filterwarnings("ignore")
        from fft_conv_pytorch.functional import fft_conv as rf, fft_conv_transpose as rft

        fn = rft if cfg["tr"] else rf
        with torch.no_grad():
            fn(x, w, b, **cfg["kw"])
            torch.cuda.synchronize(ctx.dev)
            ts = []
            for _ in range(steps):
                ctx.flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                y = fn(x, w, b, **cfg["kw"])
                e1.record()
                torch.cuda.synchronize(ctx.dev)
                ts.append(e0.elapsed_time(e1))
                del y
        return min(ts), None
    except Exception as e:  # noqa: BLE001
        return None, repr(e)[:120]
    finally:
        torch.cuda.empty_cache()


def time_config(ctx, name, cfg, batch, steps, warmup, with_ref=True, breakdown=True, seed=0):
    """Device time of one config through the public functional API (eager launches, cached kernel spectrum, L2
    flushed between steps, per-step CUDA events). batch = this rank's share (0: this rank idles)."""
    import torch

    import fft_conv_pytorch_b200 as fcp
    from fft_conv_pytorch_b200 import functional as Fn

    res = {"name": name, "batch": batch}
    if batch > 0:
        g = torch.Generator(device=ctx.dev).manual_seed(seed + 17 * ctx.rank)
        x = torch.randn(batch, *cfg["x"][1:], device=ctx.dev, generator=g)
        gw = torch.Generator().manual_seed(7)
        w = torch.randn(*cfg["w"], generator=gw).to(ctx.dev)
        b = torch.randn(cout_of(cfg), generator=gw).to(ctx.dev)
        fn = fcp.fft_conv_transpose if cfg["tr"] else fcp.fft_conv
        with torch.no_grad():
            for _ in range(warmup):
                y = fn(x, w, b, **cfg["kw"])
                del y
    ctx.barrier()
    ts = []
    if batch > 0:
        with torch.no_grad():
            for _ in range(steps):
                ctx.flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                y = fn(x, w, b, **cfg["kw"])
                e1.record()
                torch.cuda.synchronize(ctx.dev)
                ts.append(e0.elapsed_time(e1))
                del y
    ctx.barrier()
    res["ms_sum"] = sum(ts)
    res["ms_best"] = min(ts) if ts else 0.0
    res["ms"] = statistics.median(ts) if ts else 0.0
    if batch > 0:
        res["peak_gib"] = torch.cuda.max_memory_allocated(ctx.dev) / 2**30
        if breakdown and ctx.rank == 0:
            ks, a_pipe, info = kernel_breakdown(ctx, cfg, x, w, b, 3)
            dom = max(ks, key=lambda d: d["ms"])
            res.update({
                "gsamples": out_samples(cfg, batch) / (res["ms"] * 1e-3) / 1e9,
                "a_pipe_frac": a_pipe / (res["ms"] * 1e-3) / 1e9 / ctx.peak,
                "dominant_kernel": dom["kernel"], "frac": dom["algo_bytes"] / (dom["ms"] * 1e-3) / 1e9 / ctx.peak,
                "kernels": [{"k": k["kernel"], "us": round(k["ms"] * 1e3, 1), "gbs": round(k["algo_bytes"] / (k["ms"] * 1e-3) / 1e9) if k["ms"] > 0 else None}
                            for k in ks],
                "fft_size": list(info.fft_size[: info.ndim]), "segments": int(info.segments), "tensor_core": int(info.tensor_core),
            })
        if with_ref and ctx.rank == 0:
            Fn.clear_caches()
            torch.cuda.empty_cache()
            ms, why = gpu_reference_ms(ctx, cfg, x, w, b, 3)
            res["ref_gpu_ms"] = ms
            if why:
                res["ref_gpu_note"] = why
        del x, w, b
    Fn.clear_caches()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(ctx.dev)
    return res


def copy_ceiling_ms(ctx, h2d_bytes, d2h_bytes, steps):
    """What the box delivers for the host<->device traffic of one e2e step alone (no kernels): the pinned upload and
    the pinned download on two streams at once, as the host pipeline issues them; max over ranks."""
    import torch

    xh = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    yh = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    xd = torch.empty(h2d_bytes, dtype=torch.uint8, device=ctx.dev)
    yd = torch.empty(d2h_bytes, dtype=torch.uint8, device=ctx.dev)
    s1, s2 = torch.cuda.Stream(device=ctx.dev), torch.cuda.Stream(device=ctx.dev)

    def once():
        with torch.cuda.stream(s1):
            xd.copy_(xh, non_blocking=True)
        with torch.cuda.stream(s2):
            yh.copy_(yd, non_blocking=True)
        s1.synchronize()
        s2.synchronize()

    for _ in range(3):
        once()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    dt = time.perf_counter() - t0
    ctx.barrier()
    return 1e3 * ctx.max_over_ranks([dt])[0] / steps


def bench_ours(args, cfg_name):
    import torch
    import torch.distributed as dist

    import fft_conv_pytorch_b200 as fcp
    from fft_conv_pytorch_b200 import dist as fdist
    from fft_conv_pytorch_b200 import functional as Fn

    ctx = Ctx(args)
    if args.plan_flags:
        Fn.set_default_flags(args.plan_flags)
    cfg = CONFIGS[cfg_name]
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    warmup = max(args.warmup, 3)  # timing rule: at least 3 untimed steps (reported as `warmup`; `warmup_requested` is the flag)
    nd = cfg["nd"]
    module_api = not cfg["tr"] and not cfg["kw"]
    torch.manual_seed(1234 + rank)
    if module_api:  # the public nn.Module of the reference API
        mod = getattr(fcp, f"FFTConv{nd}d")(cfg["x"][1], cout_of(cfg), cfg["w"][2:]).to(dev)
        if world > 1:  # one-time weight broadcast over NVLink; no collective on the data path
            fdist.broadcast_parameters(mod, src=0)
        call_fn = mod
        weight, bias = mod.weight, mod.bias
    else:
        weight = torch.randn(*cfg["w"], device=dev)
        bias = torch.randn(cout_of(cfg), device=dev)
        if world > 1:
            dist.broadcast(weight, src=0)
            dist.broadcast(bias, src=0)
        f = fcp.fft_conv_transpose if cfg["tr"] else fcp.fft_conv
        call_fn = lambda x: f(x, weight, bias, **cfg["kw"])
    big = out_samples(cfg) * 4 > (2 << 30)
    n_rot = 1 if big else 3  # distinct resident inputs
    xs = [torch.randn(*cfg["x"], device=dev) for _ in range(n_rot)]
    samples = out_samples(cfg)

    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        # steady-state serving loop: the call is replayed as a CUDA graph, one graph per resident input
        calls = [fcp.graphed(call_fn, x) for x in xs] if not args.no_graph else [(lambda x=x: call_fn(x)) for x in xs]
        for i in range(warmup):
            calls[i % n_rot]()
        ctx.barrier()
        # ---- device-resident throughput: per-step CUDA events, L2 flushed between steps
        per_call = int(plan_entry(cfg, cfg["x"][0]).plan.info.n_launches)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        ctx.barrier()
        for i in range(args.steps):
            ctx.flush_l2()
            ev[i][0].record()
            y = calls[i % n_rot]()
            ev[i][1].record()
        ctx.barrier()
        gpu_launches = per_call * args.steps
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)
        y_shape = tuple(y.shape)

        # ---- end to end through the public API with host buffers (pinned): H2D + kernels + D2H every step
        e2e_steps = args.steps if not big else min(args.steps, 5)
        xh = [torch.randn(*cfg["x"]).pin_memory() for _ in range(2 if not big else 1)]
        yh = None
        for i in range(3):  # same call pattern as the timed loop, so no first-use pinned allocation lands in the timed region
            yh = call_fn(xh[i % len(xh)])
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            yh = call_fn(xh[i % len(xh)])  # returns a pinned CPU tensor after synchronising its stream
        torch.cuda.synchronize(dev)
        e2e_s = (time.perf_counter() - t0) * args.steps / e2e_steps
        h2d = xh[0].numel() * 4
        d2h = yh.numel() * 4
        del xh, yh
        ceiling_ms = copy_ceiling_ms(ctx, h2d, d2h, min(args.steps, 20))

        # ---- per-kernel breakdown for the roofline of the dominant kernel (CUDA events around every launch)
        breakdown = a_pipe = info = None
        if rank == 0:
            breakdown, a_pipe, info = kernel_breakdown(ctx, cfg, xs[0], weight, bias, max(min(args.steps, 50), 5))
        ctx.barrier()

        # ---- the unmodified reference on the same GPU (cuFFT + cuBLAS): the bar on this hardware
        gpu_ref = None
        if rank == 0 and not args.quick:
            ms, why = gpu_reference_ms(ctx, cfg, xs[0], weight.detach(), bias.detach(), 5)
            gpu_ref = {"ms_per_step": ms, "value": samples / (ms * 1e-3) / 1e9 if ms else None, "unit": "Gsamples/s",
                       "what": "unmodified reference (baseline/_ref) on the same GPU: torch.fft (cuFFT) + einsum (cuBLAS), best of 5, L2 flushed"}
            if why:
                gpu_ref["note"] = why
    clocks = sampler.stop() if rank == 0 else None
    dev_ms, e2e_s = ctx.max_over_ranks([dev_ms, e2e_s])
    del calls, xs, y
    Fn.clear_caches()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)

    # ---- every BASELINE shape (rank 0's GPU; N = 1 runs) and the strong-scaling records (all ranks)
    cfg_rows, strong, alternatives = [], {}, []
    if not args.quick:
        if world == 1:
            for nm in ("c1", "c2", "c3", "c4", "c5"):
                c = CONFIGS[nm]
                try:
                    r = time_config(ctx, nm, c, c["x"][0], steps=5, warmup=2, with_ref=True)
                except Exception as e:  # noqa: BLE001  (e.g. out of memory on a smaller part)
                    r = {"name": nm, "error": repr(e)[:160]}
                    Fn.clear_caches()
                    torch.cuda.empty_cache()
                r.pop("ms_sum", None)
                cfg_rows.append(r)
        # the other programs the library has for the headline shape (opt-in plan flags), same shape, same method
        if world == 1:
            for label, fl in (("packed batch pairs + y stage: K1p/K4p run one radix-4 stage of the fused axis, 128-point fused kernel with "
                               "bulk-copied kernel-spectrum chunks (FC_FLAG_PAIR)", 512),
                              ("packed batch pairs, whole fused-axis transform in the fused kernel (FC_FLAG_PAIR | FC_FLAG_NO_YSTAGE)", 512 | 1024)):
                try:
                    Fn.set_default_flags(args.plan_flags | fl)
                    r = time_config(ctx, "c2", cfg, cfg["x"][0], steps=5, warmup=2, with_ref=False)
                    r.pop("ms_sum", None)
                    r["program"] = label
                    alternatives.append(r)
                except Exception as e:  # noqa: BLE001
                    alternatives.append({"program": label, "error": repr(e)[:160]})
                finally:
                    Fn.set_default_flags(args.plan_flags)
                    Fn.clear_caches()
        # strong scaling: fixed global batch split over the ranks; N = 1 is the T1 of the efficiency T1 / (N * T_N)
        for nm, gb in (("c5", C5_GLOBAL_BATCH), ("c2", CONFIGS["c2"]["x"][0]), ("c3", CONFIGS["c3"]["x"][0])):
            c = CONFIGS[nm]
            a0, a1 = fdist.shard_range(gb, rank, world)
            try:
                r = time_config(ctx, nm, c, a1 - a0, steps=5, warmup=2, with_ref=False, breakdown=False, seed=100)
                t_med, t_best = ctx.max_over_ranks([r["ms"], r["ms_best"]])
                strong[nm] = {"global_batch": gb, "per_rank_batch": (gb + world - 1) // world, "ms_per_step": t_med, "ms_best": t_best,
                              "gsamples": out_samples(c, gb) / (t_med * 1e-3) / 1e9, "peak_gib": r.get("peak_gib")}
            except Exception as e:  # noqa: BLE001
                strong[nm] = {"global_batch": gb, "error": repr(e)[:160]}
                Fn.clear_caches()
                torch.cuda.empty_cache()

    if rank == 0:
        dom = max(breakdown, key=lambda d: d["ms"])
        achieved = dom["algo_bytes"] / (dom["ms"] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any (not measured in this run)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(dom["kernel"])
            traffic_src = tj.get("_source", "profiles/traffic.json (ncu --set full capture, committed)") if traffic is not None else None
        except Exception:
            pass
        value = samples * world * args.steps / (dev_ms * 1e-3) / 1e9
        e2e_value = samples * world * args.steps / e2e_s / 1e9
        line = {
            "metric": "fft_conv output Gsamples/s", "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "warmup_requested": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "per_gpu_batch": cfg["x"][0], "kernel_spectrum": "cached", "parallelism": f"batch-sharded x{world}",
                       "l2": ("flushed between steps (256 MiB memset, then a 256 MiB read of a second buffer so the memset's dirty lines are "
                              "written back before the timed region), per-step CUDA events") if args.flush == "write+read"
                       else "flushed between steps (256 MiB memset), per-step CUDA events",
                       "launch": "eager" if args.no_graph else "cuda-graph replay", "fft_size": list(info.fft_size[: info.ndim]),
                       "fused": int(info.fused), "out_shape": list(y_shape)},
            "e2e": {"value": e2e_value, "unit": "Gsamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps, "copy_ceiling_ms": ceiling_ms,
                    "copy_ceiling_note": "the step's pinned H2D + D2H alone (no kernels), both directions at once, max over ranks"},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": achieved, "peak": ctx.peak, "unit": "GB/s", "frac": achieved / ctx.peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": ctx.peak_src,
                         "algo_bytes_per_launch": dom["algo_bytes"], "ms_per_launch": dom["ms"]},
            "pipeline": {"a_pipe_bytes": a_pipe, "a_pipe_gbs": a_pipe / (dev_ms / args.steps * 1e-3) / 1e9,
                         "frac_of_peak": a_pipe / (dev_ms / args.steps * 1e-3) / 1e9 / ctx.peak, "kernels": breakdown},
        }
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        if cfg_rows:
            line["configs"] = cfg_rows
        if alternatives:
            line["alternative_programs"] = alternatives
        if strong:
            line["strong"] = {"note": "fixed global batch split over the ranks, device time = max over ranks; efficiency = T1 / (N * T_N) "
                                      "against the N = 1 line of the same key", **strong}
        if not args.no_cpu_baseline and world == 1:
            run, kind, cores = cpu_reference_fn(cfg)
            full_b = cfg["x"][0]
            t1 = min(run(1), run(1))
            batch = max(1, min(full_b, int(20.0 / 4 / max(t1, 1e-6))))
            run(batch)
            ts = [run(batch) for _ in range(3)]
            v = samples / full_b * batch / min(ts) / 1e9
            line["cpu_baseline"] = {"value": v, "unit": "Gsamples/s", "cores": cores, "kind": kind,
                                    "sample": f"best of 3 calls at batch {batch} of {full_b}, same kernel/channel shape"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline only: skip the per-config table, the strong-scaling records and the GPU reference")
    ap.add_argument("--flush", default="write", choices=["write", "write+read"],
                    help="L2 flush between timed steps: a 256 MiB memset; write+read adds a read of a second buffer so no dirty flush lines remain (measured: same result)")
    ap.add_argument("--plan-flags", type=int, default=0, help="FC_FLAG_* bits OR-ed into every plan (A/B runs: 256 = no batch-pair kernels)")
    ap.add_argument("--no-graph", action="store_true", help="queue the kernels from Python every step instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        bench_reference(args, CONFIGS[args.config])
    else:
        bench_ours(args, args.config)


if __name__ == "__main__":
    main()
