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


# ---- data-parallel sharding of the input pipeline (tinydiffusionmodels_b200/data.py) -----------------------------
def test_rank_batch_slices_partition_every_global_batch():
    from tinydiffusionmodels_b200.data import rank_batch_slices

    for n, bs in ((60000, 128), (1000, 32), (97, 8), (5, 4)):
        for world in (1, 2, 3, 8):
            per_rank = [rank_batch_slices(n, bs, r, world) for r in range(world)]
            steps = len(per_rank[0])
            assert all(len(p) == steps for p in per_rank)                  # every rank takes every step
            if steps == 0:
                assert n < world
                continue
            covered = []
            for k in range(steps):
                spans = [per_rank[r][k] for r in range(world)]
                assert all(hi > lo for lo, hi in spans)                    # nobody enters a step empty-handed
                assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))  # contiguous, disjoint
                assert max(hi - lo for lo, hi in spans) <= bs
                assert len({hi - lo for lo, hi in spans}) == 1             # equal local batches: equal-weight mean, disjoint noise keys
                covered.append((spans[0][0], spans[-1][1]))
            assert covered[0][0] == 0 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
            assert n - covered[-1][1] < world                              # only a < world remainder is dropped
    assert rank_batch_slices(60000, 128) == [(i, min(i + 128, 60000)) for i in range(0, 60000, 128)]   # = DataLoader
    with pytest.raises(ValueError):
        rank_batch_slices(10, 0)
    with pytest.raises(ValueError):
        rank_batch_slices(10, 2, rank=2, world=2)


def _data_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tinydiffusionmodels_b200 import data as D
        from tinydiffusionmodels_b200 import ops

        n = 1003
        g = torch.Generator().manual_seed(5)
        images = torch.randint(0, 256, (n, 4, 4), generator=g, dtype=torch.uint8)
        # host logic of DeviceImages.batches on CPU tensors: the device kernel is stood in for by the oracle
        ds = object.__new__(D.DeviceImages)
        ds.images, ds.device, ds.mean, ds.std = images, torch.device("cpu"), 0.5, 0.5
        ds.permutation = lambda seed, epoch: torch.randperm(n, generator=torch.Generator().manual_seed(100 * rank + epoch))
        ops.normalize_u8 = lambda im, idx, mean, std, check_index=True: torch.cat(
            [O.normalize_u8(im, idx, mean, std).reshape(len(idx), -1), idx.view(-1, 1).float()], 1)
        got = list(ds.batches(16, seed=1, epoch=0))        # rank / world come from the process group
        mine = torch.cat(got)
        sizes = torch.tensor([len(got), mine.shape[0]])
        all_sizes = [torch.empty_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes)
        pad = torch.zeros(n, mine.shape[1])
        pad[: mine.shape[0]] = mine
        gathered = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad)
        if rank == 0:
            rows = torch.cat([gathered[r][: int(all_sizes[r][1])] for r in range(world)])
            idx = rows[:, -1].long()
            ok = (len({int(s[0]) for s in all_sizes}) == 1                                    # same number of steps
                  and int(idx.min()) >= 0 and int(idx.max()) < n
                  and len(set(idx.tolist())) == len(idx) and n - len(idx) < world            # disjoint, near-complete
                  and torch.equal(rows[:, :-1], O.normalize_u8(images, idx).reshape(len(idx), -1)))
            out.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_input_pipeline_shards_across_two_ranks():
    """world_size-2 gloo: both ranks slice rank 0's permutation (their own differ on purpose), see disjoint images,
    take the same number of steps and together cover the epoch."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_data_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=10) is True
