"""GPU parity of the input pipeline and the output step against the oracle and the
torchvision-recorded golden vectors: bit-exact, through the C ABI."""
import math
from pathlib import Path

import pytest
import torch

from oracle import ddpm_oracle as O
from tinydiffusionmodels_b200 import ops
from tinydiffusionmodels_b200.data import DeviceImages

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gio():
    return torch.load(GOLD / "io_golden.pt", weights_only=False)


def test_normalize_u8_golden(cuda, gio):
    imgs = gio["images"].to(cuda)
    assert torch.equal(ops.normalize_u8(imgs).cpu(), gio["normalized"])
    assert torch.equal(ops.normalize_u8(imgs, gio["index"].to(cuda)).cpu(), gio["normalized"][gio["index"]])
    assert torch.equal(ops.normalize_u8(imgs.unsqueeze(1)).cpu(), gio["normalized"])          # (N,1,H,W) accepted


@pytest.mark.parametrize("n,h,w,mean,std", [(1, 28, 28, 0.5, 0.5), (257, 28, 28, 0.5, 0.5), (33, 4, 4, 0.1307, 0.3081),
                                            (5, 32, 36, 0.0, 1.0), (3, 40, 52, 0.25, 2.0)])
def test_normalize_u8_shapes_and_constants(cuda, n, h, w, mean, std):
    g = torch.Generator().manual_seed(n)
    imgs = torch.randint(0, 256, (n, h, w), generator=g, dtype=torch.uint8)
    idx = torch.randint(0, n, (2 * n + 1,), generator=g)                                        # repeats allowed
    got = ops.normalize_u8(imgs.to(cuda), idx.to(cuda), mean, std).cpu()
    assert torch.equal(got, O.normalize_u8(imgs, idx, mean, std))


def test_normalize_u8_edges(cuda):
    imgs = torch.zeros(3, 28, 28, dtype=torch.uint8, device=cuda)
    assert ops.normalize_u8(imgs, torch.empty(0, dtype=torch.int64, device=cuda)).shape == (0, 1, 28, 28)
    with pytest.raises(IndexError):
        ops.normalize_u8(imgs, torch.tensor([3], device=cuda))
    with pytest.raises(ValueError):
        ops.normalize_u8(imgs.float())
    with pytest.raises(Exception):
        ops.normalize_u8(torch.zeros(2, 3, 5, dtype=torch.uint8, device=cuda))                 # H*W % 4 != 0


def test_full_mnist_size_epoch(cuda):
    """60,000 images: every image is served exactly once per epoch, the last batch is partial (drop_last=False),
    two epochs differ, and the whole epoch equals the oracle transform of the permuted set (checksum of all
    batches against torch on the same uint8 data)."""
    g = torch.Generator().manual_seed(3)
    imgs = torch.randint(0, 256, (60000, 28, 28), generator=g, dtype=torch.uint8)
    ds = DeviceImages(imgs, cuda)
    assert len(ds) == 60000 and ds.num_batches(128) == 469
    order = ds.permutation(seed=11, epoch=0)
    assert torch.equal(torch.sort(order).values, torch.arange(60000, device=cuda))
    assert not torch.equal(order, ds.permutation(seed=11, epoch=1))
    batches = list(ds.batches(128, seed=11, epoch=0))
    assert len(batches) == 469 and batches[-1].shape == (96, 1, 28, 28) and batches[0].shape == (128, 1, 28, 28)
    got = torch.cat(batches).cpu()
    assert torch.equal(got, O.normalize_u8(imgs, order.cpu()))
    unshuffled = torch.cat(list(ds.batches(4096, shuffle=False, max_batches=2))).cpu()
    assert torch.equal(unshuffled, O.normalize_u8(imgs[:8192]))


@pytest.mark.parametrize("n", [1, 4, 7, 25, 30])
def test_image_grid_u8_golden(cuda, gio, n):
    g = gio["grids"][n]
    grid = ops.image_grid_u8(g["x"].to(cuda), nrow=int(math.sqrt(n)), from_signed=True)
    assert grid.dtype == torch.uint8 and torch.equal(grid.cpu(), g["grid"])
    # the two-step form (unit-range kernel, then the grid of a [0,1] tensor) gives the same bytes
    assert torch.equal(ops.image_grid_u8(ops.to_unit_range(g["x"].to(cuda)), nrow=int(math.sqrt(n))).cpu(), g["grid"])


@pytest.mark.parametrize("n,h,w,nrow,pad", [(10, 28, 28, 8, 2), (9, 14, 20, 3, 0), (64, 28, 28, 8, 2), (5, 8, 8, 1, 3),
                                            (1024, 28, 28, 32, 2)])
def test_image_grid_u8_shapes(cuda, n, h, w, nrow, pad):
    g = torch.Generator().manual_seed(n)
    x = torch.rand(n, 1, h, w, generator=g) * 1.2 - 0.1            # a little outside [0,1]: the clamp matters
    got = ops.image_grid_u8(x.to(cuda), nrow=nrow, padding=pad).cpu()
    assert torch.equal(got, O.image_grid_u8(x, nrow=nrow, padding=pad))


def test_sample_cli_writes_the_grid_the_reference_would(cuda, tmp_path, monkeypatch):
    """End of the sampling path: the PNG on disk decodes to the oracle's grid of the very x_0 the loop produced."""
    import io

    import numpy as np
    from PIL import Image

    from tinydiffusionmodels_b200 import mnist

    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    x = torch.randn(25, 1, 28, 28, device=cuda) * 0.7
    path = mnist._save_grid(x, str(tmp_path), "samples.png", from_signed=True)
    arr = np.array(Image.open(io.BytesIO(open(path, "rb").read())))
    assert np.array_equal(arr, O.image_grid_u8(O.to_unit_range(x.cpu()), nrow=5).numpy())
