"""Device-resident MNIST input pipeline.

The reference feeds training from ``DataLoader(datasets.MNIST(..., transform=ToTensor+Normalize), shuffle=True,
num_workers=4, pin_memory=True)`` (src/mnist.py:139-147): four host workers decode and normalise 28x28 images one
by one and every batch crosses PCIe as fp32.  The whole training set is 60,000 x 784 bytes = 47 MB as uint8, so here
it lives in HBM once; an epoch is a device-side permutation and one gather + normalise kernel per batch
(``tdm_u8_gather_normalize``, bit-identical to the reference transform), with no host work and no host sync.
"""
from __future__ import annotations

from typing import Iterator

import torch

from . import _lib, ops


class DeviceImages:
    """uint8 images (N, H, W) resident on a CUDA device, served as shuffled, normalised fp32 batches.

    Batch semantics follow the reference DataLoader: a fresh permutation per epoch, ``batch_size`` images per
    batch, the last partial batch kept (``drop_last=False``).
    """

    def __init__(self, images: torch.Tensor, device, mean: float = 0.5, std: float = 0.5):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TdmError("DeviceImages needs a CUDA device (no CPU fallback)")
        if images.dtype != torch.uint8:
            raise ValueError("DeviceImages expects uint8 images")
        if images.dim() == 4 and images.shape[1] == 1:
            images = images[:, 0]
        if images.dim() != 3 or (images.shape[1] * images.shape[2]) % 4 != 0:
            raise ValueError("DeviceImages expects (N, H, W) images with H*W a multiple of 4")
        self.images = images.to(self.device).contiguous()
        self.device = self.images.device   # resolved index
        self.mean, self.std = float(mean), float(std)

    def __len__(self) -> int:
        return int(self.images.shape[0])

    def num_batches(self, batch_size: int) -> int:
        return (len(self) + batch_size - 1) // batch_size

    def permutation(self, seed: int, epoch: int) -> torch.Tensor:
        """The epoch's visiting order: ``torch.randperm`` on the device, seeded by (seed, epoch)."""
        g = torch.Generator(device=self.device)
        g.manual_seed((int(seed) * 1_000_003 + int(epoch)) & 0x7FFF_FFFF_FFFF_FFFF)
        return torch.randperm(len(self), device=self.device, generator=g)

    def batches(self, batch_size: int, *, seed: int = 0, epoch: int = 0, shuffle: bool = True,
                max_batches: int | None = None, rank: int | None = None, world: int | None = None
                ) -> Iterator[torch.Tensor]:
        """Yield this rank's (b, 1, H, W) fp32 batches of one epoch (``batch_size`` images per rank and step).

        Data parallel: every rank slices the same permutation (rank 0's, broadcast once per epoch) and takes its own
        part of each global batch of ``world * batch_size`` images (``rank_batch_slices``), so the ranks see disjoint
        data and take the same number of steps; no image data ever crosses a GPU boundary.  ``rank`` / ``world`` default to the
        initialised ``torch.distributed`` group, else to a single process.
        """
        if rank is None or world is None:
            rank, world = _rank_world()
        order = self.permutation(seed, epoch) if shuffle else None
        if order is not None and world > 1:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                dist.broadcast(order, src=0)   # once per epoch, 8 B per image: every rank slices rank 0's order
        for i, (lo, hi) in enumerate(rank_batch_slices(len(self), batch_size, rank, world)):
            if max_batches is not None and i >= max_batches:
                return
            idx = torch.arange(lo, hi, device=self.device) if order is None else order[lo:hi]
            yield ops.normalize_u8(self.images, idx, self.mean, self.std, check_index=False)


def rank_batch_slices(n: int, batch_size: int, rank: int = 0, world: int = 1) -> list[tuple[int, int]]:
    """Positions [lo, hi) of the epoch's visiting order that ``rank`` takes at each step.

    Step k covers the global batch [k*G, min((k+1)*G, n)) with G = world * batch_size, split contiguously and
    EVENLY over the ranks.  The last, partial global batch is kept like the reference's ``drop_last=False``
    (src/mnist.py:146), trimmed to a multiple of ``world`` (at most world-1 images of an epoch are skipped): every
    rank must take part in every step's gradient exchange with the same local batch, because the optimizer
    averages the ranks' mean gradients with equal weights and the training noise is keyed on rank*b + row.
    """
    from .dist import shard_range

    if batch_size < 1:
        raise ValueError("batch_size must be positive")
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    out = []
    g = batch_size * world
    for start in range(0, n, g):
        size = min(g, n - start)
        size -= size % world
        if size == 0:
            break
        lo, hi = shard_range(size, rank, world)
        out.append((start + lo, start + hi))
    return out


def _rank_world() -> tuple[int, int]:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def mnist_on_device(device, root: str = "./data", train: bool = True, download: bool = True) -> DeviceImages:
    """The reference's dataset (src/mnist.py:139-145), decoded once by torchvision and kept on the device."""
    from torchvision import datasets

    ds = datasets.MNIST(root, train=train, download=download)
    return DeviceImages(ds.data, device)   # (60000, 28, 28) uint8
