#!/usr/bin/env python3
"""Benchmark of the fft_conv hot path (BASELINE.json metric: output Gsamples/s; achieved HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c3]

A step is one pass of the hot path over one batch of synthetic input of the BASELINE config (default c2:
FFTConv2d fp32, input (8,8,512,512), 8->8 channels, 65x65 kernel, bias), with the kernel spectrum cached (steady
state). With N > 1 (torchrun, one rank per GPU) every rank convolves its own batch of the same shape
with weights broadcast once from rank 0 (weak scaling, no data-path collective); `value` is the whole-job
throughput: samples of all ranks / max-over-ranks device time.

One JSON line is printed by rank 0. `--impl reference` times the reference's own CPU implementation of the path
(baseline/_ref when present, else the oracle port) on the host cores.
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
    # name: (module ndim, x shape, cin, cout, kernel)
    "c1": dict(nd=1, x=(1, 8, 32768), cin=8, cout=8, k=1025, desc="1D fft_conv fp32 (1,8,32768) k1025"),
    "c2": dict(nd=2, x=(8, 8, 512, 512), cin=8, cout=8, k=65, desc="2D FFTConv2d fp32 (8,8,512,512) 8->8 k65x65 bias"),
    "c3": dict(nd=3, x=(4, 8, 64, 64, 64), cin=8, cout=8, k=17, desc="3D FFTConv3d fp32 (4,8,64,64,64) 8->8 k17^3 bias"),
}


def out_samples(cfg):
    sp = [s - cfg["k"] + 1 for s in cfg["x"][2:]]
    n = cfg["x"][0] * cfg["cout"]
    for s in sp:
        n *= s
    return n


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_reference_fn(cfg):
    """Returns (callable(batch) -> seconds per call, kind, cores). Reference from baseline/_ref if it travelled,
    else the oracle port (numpy/scipy pocketfft with all cores)."""
    import numpy as np
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(0)
    nd = cfg["nd"]
    w = torch.randn(cfg["cout"], cfg["cin"], *([cfg["k"]] * nd), generator=g)
    b = torch.randn(cfg["cout"], generator=g)
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    kind = "port"
    fn = None
    if os.path.isdir(os.path.join(ref_dir, "fft_conv_pytorch")):
        try:
            sys.path.insert(0, ref_dir)
            import warnings

            warnings.filterwarnings("ignore")
            from fft_conv_pytorch.functional import fft_conv as ref_fft_conv  # the unmodified reference

            def fn(x):
                with torch.no_grad():
                    return ref_fft_conv(x, w, b)

            kind = "reference"
        except Exception:
            fn = None
    if fn is None:
        from oracle import fftconv_oracle as O

        wn, bn = w.numpy(), b.numpy()

        def fn(x):
            return O.fft_conv(x.numpy(), wn, bn, workers=cores)

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
def bench_ours(args, cfg):
    import torch
    import torch.distributed as dist

    import fft_conv_pytorch_b200 as fcp
    from fft_conv_pytorch_b200 import functional as Fn
    from fft_conv_pytorch_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    nd = cfg["nd"]
    torch.manual_seed(1234 + rank)
    mod = getattr(fcp, f"FFTConv{nd}d")(cfg["cin"], cfg["cout"], cfg["k"]).to(dev)
    if world > 1:  # one-time weight broadcast over NVLink; no collective on the data path
        from fft_conv_pytorch_b200 import dist as fdist

        fdist.broadcast_parameters(mod, src=0)
    n_rot = 3  # distinct resident inputs
    xs = [torch.randn(*cfg["x"], device=dev) for _ in range(n_rot)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    flush_rd = torch.zeros(256 << 18, dtype=torch.int32, device=dev)  # second 256 MiB buffer, only ever read

    def flush_l2():
        """Evict everything of ours from L2 before a timed step: write 256 MiB; optionally also read a second 256 MiB
        buffer so that the dirty lines the memset leaves in L2 are written back before the timed region starts
        (measured on B200: 0.1355 vs 0.1362 ms/step, no real difference)."""
        flush.zero_()
        if args.flush == "write+read":
            flush_rd.sum()
    samples = out_samples(cfg)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    with torch.no_grad():
        # steady-state serving loop: the call (3 kernels) is replayed as a CUDA graph, one graph per resident input
        calls = [fcp.graphed(mod, x) for x in xs] if not args.no_graph else [(lambda x=x: mod(x)) for x in xs]
        for i in range(max(args.warmup, 3)):
            calls[i % n_rot]()
        barrier()
        # ---- device-resident throughput: per-step CUDA events, L2 flushed between steps
        per_call = int(Fn.get_plan(False, cfg["x"][0], cfg["cin"], cfg["cout"], 1, tuple(cfg["x"][2:]), (cfg["k"],) * nd, (1,) * nd, (0,) * nd,
                                   (1,) * nd, (0,) * nd, "constant").plan.info.n_launches)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        for i in range(args.steps):
            flush_l2()
            ev[i][0].record()
            y = calls[i % n_rot]()
            ev[i][1].record()
        barrier()
        gpu_launches = per_call * args.steps
        dev_ms = sum(a.elapsed_time(b) for a, b in ev)

        # ---- end to end through the public API with host buffers (pinned): H2D + kernels + D2H every step
        xh = [torch.randn(*cfg["x"]).pin_memory() for _ in range(2)]
        yh = None
        for i in range(max(args.warmup, 3)):  # same call pattern as the timed loop (the previous result is still referenced while
            yh = mod(xh[i % 2])               # the next one is allocated), so no first-use pinned allocation lands in the timed region
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            yh = mod(xh[i % 2])  # returns a pinned CPU tensor after synchronising its stream
        torch.cuda.synchronize(dev)
        e2e_s = time.perf_counter() - t0
        h2d = xh[0].numel() * 4
        d2h = yh.numel() * 4

        # ---- per-kernel breakdown for the roofline of the dominant kernel (CUDA events around every launch)
        breakdown = None
        if rank == 0:
            entry = Fn.get_plan(False, cfg["x"][0], cfg["cin"], cfg["cout"], 1, tuple(cfg["x"][2:]), (cfg["k"],) * nd, (1,) * nd, (0,) * nd,
                                (1,) * nd, (0,) * nd, "constant")
            plan = entry.plan
            lib = plan.lib
            kspec = Fn.kernel_spectrum(entry, mod.weight, dev)
            const = entry.const_for(dev)
            ws = torch.empty(int(plan.info.workspace_bytes), dtype=torch.uint8, device=dev)
            yb = torch.empty_like(y)
            nl = int(plan.info.n_launches)
            acc = [0.0] * nl
            ms = (ctypes.c_float * nl)()
            n_out = ctypes.c_int(0)
            P = lambda t: ctypes.c_void_p(t.data_ptr())
            stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            reps = max(args.steps, 5)
            for i in range(reps):
                flush_l2()
                L.check(lib, lib.fc_conv_profiled(plan.handle, P(const), P(xs[i % n_rot]), P(kspec), P(mod.bias), P(yb), P(ws), stream, ms, nl,
                                                  ctypes.byref(n_out)), "fc_conv_profiled")
                for j in range(n_out.value):
                    acc[j] += ms[j]
            breakdown = []
            for j in range(nl):
                name = ctypes.create_string_buffer(64)
                ab = ctypes.c_int64(0)
                lib.fc_plan_launch_info(plan.handle, j, name, 64, ctypes.byref(ab))
                breakdown.append({"kernel": name.value.decode(), "ms": acc[j] / reps, "algo_bytes": ab.value})

    clocks = sampler.stop() if rank == 0 else None
    # max over ranks
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = t[0].item(), t[1].item()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        dom = max(breakdown, key=lambda d: d["ms"])
        achieved = dom["algo_bytes"] / (dom["ms"] * 1e-3) / 1e9
        traffic = None
        try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu capture, if any
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom["kernel"])
        except Exception:
            pass
        value = samples * world * args.steps / (dev_ms * 1e-3) / 1e9
        e2e_value = samples * world * args.steps / e2e_s / 1e9
        info = Fn.get_plan(False, cfg["x"][0], cfg["cin"], cfg["cout"], 1, tuple(cfg["x"][2:]), (cfg["k"],) * nd, (1,) * nd, (0,) * nd, (1,) * nd,
                           (0,) * nd, "constant").plan.info
        a_pipe = info.algo_bytes_s1 + info.algo_bytes_s3 + info.algo_bytes_s4
        line = {
            "metric": "fft_conv output Gsamples/s", "value": value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["desc"], "per_gpu_batch": cfg["x"][0], "kernel_spectrum": "cached", "parallelism": f"batch-sharded x{world}",
                       "l2": ("flushed between steps (256 MiB memset, then a 256 MiB read of a second buffer so the memset's dirty lines are "
                              "written back before the timed region), per-step CUDA events") if args.flush == "write+read"
                       else "flushed between steps (256 MiB memset), per-step CUDA events",
                       "launch": "eager" if args.no_graph else "cuda-graph replay", "fft_size": list(info.fft_size[: info.ndim]),
                       "fused": int(info.fused)},
            "e2e": {"value": e2e_value, "unit": "Gsamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s / args.steps},
            "gpu_launches": gpu_launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "algo_bytes_per_launch": dom["algo_bytes"], "ms_per_launch": dom["ms"]},
            "pipeline": {"a_pipe_bytes": a_pipe, "a_pipe_gbs": a_pipe / (dev_ms / args.steps * 1e-3) / 1e9,
                         "frac_of_peak": a_pipe / (dev_ms / args.steps * 1e-3) / 1e9 / peak, "kernels": breakdown},
        }
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
    ap.add_argument("--flush", default="write", choices=["write", "write+read"],
                    help="L2 flush between timed steps: a 256 MiB memset; write+read adds a read of a second buffer so no dirty flush lines remain (measured: same result)")
    ap.add_argument("--no-graph", action="store_true", help="queue the kernels from Python every step instead of replaying a CUDA graph")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        bench_reference(args, cfg)
    else:
        bench_ours(args, cfg)


if __name__ == "__main__":
    main()
