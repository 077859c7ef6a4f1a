#!/usr/bin/env python3
"""BASELINE c5 sharded over the GPUs of one box (torchrun, one rank per GPU): fft_conv_transpose, input (32,64,1024,1024),
kernel (64,16,31,31), stride 2, dilation 2, groups 4; rank r convolves its B/world samples. The only communication is
the one-time parameter broadcast (NCCL); the script also times shipping the kernel spectrum instead of rebuilding it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/c5_sharded.py [--steps K]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import fft_conv_pytorch_b200 as fcp
from fft_conv_pytorch_b200 import dist as fdist
from fft_conv_pytorch_b200 import functional as Fn
from tests.helpers import spot_check


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    kw = dict(stride=2, dilation=2, groups=4)
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 16, 31, 31, generator=g).to(dev) if rank == 0 else torch.empty(64, 16, 31, 31, device=dev)
    b = torch.randn(64, generator=g).to(dev) if rank == 0 else torch.empty(64, device=dev)
    a0, a1 = fdist.shard_range(args.batch, rank, world)
    xg = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(a1 - a0, 64, 1024, 1024, generator=xg).to(dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def sync():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # one-time: parameters over NCCL; the kernel spectrum is built on rank 0 and broadcast over NVLink (258 MB with
    # overlap-save segments), then served from every rank's cache
    entry = Fn.get_plan(True, a1 - a0, 64, 64, 4, (1024, 1024), (31, 31), (2, 2), (0, 0), (2, 2), (0, 0), "constant")
    entry.const_for(dev)
    sync()
    e0, e1, e2 = ev(), ev(), ev()
    e0.record()
    if world > 1:
        dist.broadcast(w, src=0)
        dist.broadcast(b, src=0)
    e1.record()
    kspec = fdist.broadcast_kernel_spectrum(entry, w, dev, src=0)
    e2.record()
    sync()
    t_bcast_w, t_kspec_shared = e0.elapsed_time(e1), e1.elapsed_time(e2)
    t_bcast_k = None
    if world > 1:
        tmp = torch.empty_like(kspec)
        sync()
        e0.record()
        dist.broadcast(tmp, src=0)
        e1.record()
        sync()
        t_bcast_k = e0.elapsed_time(e1)
        del tmp
    sync()
    e0.record()
    Fn.kernel_spectrum(entry, w, dev, use_cache=False)  # for comparison: every rank rebuilding its own
    e1.record()
    sync()
    t_k2 = e0.elapsed_time(e1)
    with torch.no_grad():
        y = fcp.fft_conv_transpose(x, w, b, **kw)
    sync()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with torch.no_grad():
        for _ in range(3):
            y = fcp.fft_conv_transpose(x, w, b, **kw)
        sync()
        tot = 0.0
        for _ in range(args.steps):
            flush.zero_()
            e0.record()
            y = fcp.fft_conv_transpose(x, w, b, **kw)
            e1.record()
            torch.cuda.synchronize(dev)
            tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    err, n = spot_check(x.cpu().numpy(), w.cpu().numpy(), b.cpu().numpy(), y, n=40, transposed=True, **kw)
    errt = torch.tensor([err], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(errt, op=dist.ReduceOp.MAX)
    if rank == 0:
        samples = args.batch * 64 * 2107 * 2107
        print(json.dumps({"config": "c5 fft_conv_transpose (32,64,1024,1024) k31 s2 d2 g4", "n_gpus": world, "batch_per_gpu": a1 - a0,
                          "ms_per_step_max_over_ranks": ms, "gsamples_per_s": samples / ms / 1e6, "spot_check_rel_err_max_over_ranks": errt.item(),
                          "fft_size": list(entry.plan.fft_size), "segments": int(entry.plan.info.segments), "kspec_mib": kspec.numel() * 4 / 2**20,
                          "one_time_ms": {"weights_broadcast_nccl": t_bcast_w, "kernel_spectrum_build_on_rank0_plus_broadcast": t_kspec_shared,
                                          "kernel_spectrum_broadcast_nccl_alone": t_bcast_k, "kernel_spectrum_rebuild_on_every_rank": t_k2}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
