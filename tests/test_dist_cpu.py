"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, shard-invariant noise,
and lock-step data-parallel replicas after the flat-gradient all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ddpm_oracle as O
from oracle import philox as PX
from tinydiffusionmodels_b200.dist import allreduce_mean_, shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 64, 1000, 262144):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_sharded_noise_equals_unsharded():
    """The per-step noise of sample i depends on (seed, i, t) only: any sharding reproduces it."""
    full = PX.randn(10, 784, 42, 0, 500, PX.DOMAIN_REVERSE)
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi = shard_range(10, r, world)
            parts.append(PX.randn(hi - lo, 784, 42, lo, 500, PX.DOMAIN_REVERSE))
        assert np.array_equal(np.concatenate(parts), full)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                       # identical initial replicas
        p = torch.randn(1000)
        m, v = torch.zeros(1000), torch.zeros(1000)
        g_all = [torch.Generator().manual_seed(100 + r) for r in range(world)]
        for k in range(1, 4):
            grads = [torch.randn(1000, generator=g_all[r]) for r in range(world)]   # every rank can replay all
            mine = grads[rank].clone()
            allreduce_mean_(mine)
            assert torch.allclose(mine, torch.stack(grads).mean(0), atol=1e-6)
            p, m, v = O.adamw_step(p, mine, m, v, k)
        gathered = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(gathered, p)
        if rank == 0:
            out.put(all(torch.equal(gathered[0], x) for x in gathered))
    finally:
        dist.destroy_process_group()


def test_data_parallel_replicas_stay_in_lock_step():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
