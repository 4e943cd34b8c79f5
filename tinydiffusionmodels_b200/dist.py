"""Multi-GPU plumbing for the two ways the hot path scales.

* Sampling shards by *independent units*: rank r draws global samples [lo, hi); the Philox noise is
  keyed on the global sample index, so the union over ranks is bit-identical for any world size.
  No collective on the data path.
* Training is data-parallel: one sum all-reduce (NCCL over NVLink; gloo in CPU tests) of the flat
  181,473-element fp32 gradient buffer per step, averaged by the fused AdamW's grad_scale.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split of n units over `world` ranks; the first n % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def init_from_env(backend: str | None = None, device=None):
    """torchrun-style init (RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT). Returns (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def allreduce_mean_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean of the flat gradient buffer across ranks (what UNetTrainer does as
    sum all-reduce + grad_scale=1/world folded into AdamW)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, group=group)
        flat_grad.mul_(1.0 / dist.get_world_size(group))
    return flat_grad
