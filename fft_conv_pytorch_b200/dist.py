"""Multi-GPU use of the path: one process per GPU, batch (or output-channel) partition, no data-path collective.

An output element (b, o, ·) depends only on x[b, group(o) channels] and w[o] (reference functional.py:12-16), so the
work shards by batch and by output channel with no reduction. The only communication is a one-time
broadcast of the parameters (``torch.distributed``: NCCL over NVLink on GPUs, gloo in the CPU tests); every rank
then builds its own kernel spectrum locally, which is cheaper than shipping the spectrum (K << N).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) share of `total` items for `rank`; the first `total % world` ranks get one extra."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """Broadcast weight/bias from `src` so every rank convolves with the same kernel."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    with torch.no_grad():
        for p in module.parameters():
            dist.broadcast(p.detach(), src=src)
            # The collective writes the storage without touching the version counter the spectrum cache checks
            # (and `p.data` has a counter of its own, so an edit through it would not either): an in-place op on a
            # detached alias shares the parameter's counter and invalidates a spectrum cached before the broadcast.
            p.detach().add_(0)


def broadcast_kernel_spectrum(entry, kernel: torch.Tensor, device: torch.device, src: int = 0) -> torch.Tensor:
    """Build the cached kernel spectrum of `kernel` for plan entry `entry` (``functional.get_plan``) on rank `src` only and
    ship it to the other ranks (NCCL broadcast over NVLink), which install it in their caches. Worth it when the spectrum
    is small: with overlap-save segments BASELINE c5's is 258 MB and the broadcast takes 0.45 ms on B200 NVLink against
    3.2 ms for rebuilding it on every rank; an unsegmented multi-GB spectrum is cheaper to rebuild (K << N)."""
    from . import functional as Fn

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return Fn.kernel_spectrum(entry, kernel, device)
    if dist.get_rank() == src:
        kspec = Fn.kernel_spectrum(entry, kernel, device)
    else:
        kspec = torch.empty(int(entry.plan.info.kspec_bytes) // 4, dtype=torch.float32, device=device)
    dist.broadcast(kspec, src=src)
    if dist.get_rank() != src:
        Fn.install_kernel_spectrum(entry, kernel, device, kspec)
    return kspec


def shard_batch(x: torch.Tensor, rank: int = None, world: int = None) -> torch.Tensor:
    """This rank's slice of a replicated batch."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    a, b = shard_range(x.shape[0], rank, world)
    return x[a:b]


def shard_out_channels(weight: torch.Tensor, bias, groups: int, rank: int = None, world: int = None, transposed: bool = False):
    """Output-channel partition on group boundaries (for batch-1 problems): returns (weight_shard, bias_shard,
    groups_shard, input-channel slice). Requires `groups % world == 0` or `groups == 1`."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    cout = weight.shape[1] * groups if transposed else weight.shape[0]
    if groups == 1:
        a, b = shard_range(cout, rank, world)
        w = weight[:, a:b] if transposed else weight[a:b]
        return w.contiguous(), None if bias is None else bias[a:b].contiguous(), 1, slice(None)
    if groups % world:
        raise ValueError(f"groups ({groups}) must be divisible by the world size ({world}) to shard output channels")
    ga, gb = shard_range(groups, rank, world)
    og = cout // groups
    cin = weight.shape[0] if transposed else weight.shape[1] * groups
    ig = cin // groups
    w = weight[ga * ig:gb * ig] if transposed else weight[ga * og:gb * og]
    bsh = None if bias is None else bias[ga * og:gb * og].contiguous()
    return w.contiguous(), bsh, gb - ga, slice(ga * ig, gb * ig)


def all_gather_batch(y: torch.Tensor, sizes: List[int] = None) -> torch.Tensor:
    """Optional: assemble the batch-sharded result on every rank (small configs only; large outputs stay sharded).
    Ragged shards are padded to the largest one for the collective and trimmed afterwards."""
    world = dist.get_world_size()
    if sizes is None:
        sizes = [y.shape[0]] * world
    big = max(sizes)
    pad = y.contiguous()
    if y.shape[0] < big:
        pad = torch.cat([pad, pad.new_zeros((big - y.shape[0],) + tuple(y.shape[1:]))], 0)
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:s] for o, s in zip(outs, sizes)], 0)
