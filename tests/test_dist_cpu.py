"""world_size-2 gloo test of the multi-GPU host logic (partition + one-time parameter broadcast, no data-path
collective). The device op is replaced by the oracle here (no GPU in the CPU suite); the `-m gpu`
suite and bench.py run the same helpers over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fft_conv_pytorch_b200 import dist as fdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    from oracle import fftconv_oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)  # ranks start with different parameters
        conv = torch.nn.Conv2d(4, 6, 5, groups=2)
        fdist.broadcast_parameters(conv, src=0)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(5, 4, 20, 16, generator=g)  # replicated input, odd batch -> ragged shards
        xs = fdist.shard_batch(x)
        w, b = conv.weight.detach().numpy(), conv.bias.detach().numpy()
        ys = torch.from_numpy(O.fft_conv(xs.numpy(), w, b, groups=2))
        sizes = [fdist.shard_range(5, r, world)[1] - fdist.shard_range(5, r, world)[0] for r in range(world)]
        y = fdist.all_gather_batch(ys, sizes)
        # output-channel partition on group boundaries, full batch per rank
        wsh, bsh, gsh, cin_sl = fdist.shard_out_channels(conv.weight.detach(), conv.bias.detach(), 2)
        yo = torch.from_numpy(O.fft_conv(x[:, cin_sl].numpy(), wsh.numpy(), bsh.numpy(), groups=gsh))
        outs = [torch.empty_like(yo) for _ in range(world)]
        dist.all_gather(outs, yo)
        q.put((rank, y.numpy(), torch.cat(outs, 1).numpy(), w, b))
    finally:
        dist.destroy_process_group()


def test_batch_and_channel_partition_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda r: r[0])
    # same parameters everywhere after the broadcast
    assert np.array_equal(res[0][3], res[1][3]) and np.array_equal(res[0][4], res[1][4])
    g = torch.Generator().manual_seed(7)
    x = torch.randn(5, 4, 20, 16, generator=g)
    ref = torch.nn.functional.conv2d(x, torch.from_numpy(res[0][3]), torch.from_numpy(res[0][4]), groups=2).numpy()
    for r in res:
        assert r[1].shape == ref.shape and np.abs(r[1] - ref).max() < 1e-4  # batch-sharded, gathered
        assert r[2].shape == ref.shape and np.abs(r[2] - ref).max() < 1e-4  # channel-sharded, gathered


def _kspec_worker(rank, world, port, q):
    from fft_conv_pytorch_b200 import functional as Fn

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # host plan only (no GPU here): rank 0 "owns" a spectrum, the others receive and cache it
        entry = Fn.get_plan(True, 2, 8, 8, 2, (520, 300), (5, 5), (2, 2), (0, 0), (2, 2), (0, 0), "constant")
        w = torch.randn(8, 4, 5, 5)
        cpu = torch.device("cpu")
        n = int(entry.plan.info.kspec_bytes) // 4
        if rank == 0:
            Fn.install_kernel_spectrum(entry, w, cpu, torch.arange(n, dtype=torch.float32))
        k = fdist.broadcast_kernel_spectrum(entry, w, cpu, src=0)
        hit = Fn.kernel_spectrum(entry, w, cpu)  # served from the cache: no device work
        ok = hit.data_ptr() == k.data_ptr() and bool((k == torch.arange(n, dtype=torch.float32)).all())
        w.add_(1.0)  # a weight update invalidates the received spectrum like a local one
        stale = Fn._kspec_cache.get((id(entry.plan), id(w), -1))
        q.put((rank, ok, stale is not None and stale[1] != w._version, int(entry.plan.info.segments)))
    finally:
        dist.destroy_process_group()


def test_kernel_spectrum_broadcast_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_kspec_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, version_moved, segments in res:
        assert ok and version_moved and segments > 1


@pytest.mark.parametrize("total,world", [(8, 2), (5, 2), (32, 8), (3, 4), (1, 8)])
def test_shard_range_covers_everything(total, world):
    spans = [fdist.shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    for a, b in zip(spans, spans[1:]):
        assert a[1] == b[0]
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1
