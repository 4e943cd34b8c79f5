"""Input transform / output step: pin the oracle to torchvision's own outputs recorded
in tests/golden/io_golden.pt (made by tests/golden/make_golden_io.py), check the host-side pieces, and the
checkpoint round trip with the reference's on-disk format.  CPU only."""
import io
import math
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def gio():
    return torch.load(GOLD / "io_golden.pt", weights_only=False)


def test_normalize_u8_matches_torchvision_transform(gio):
    out = O.normalize_u8(gio["images"])
    assert out.shape == (64, 1, 28, 28) and torch.equal(out, gio["normalized"])
    assert torch.equal(O.normalize_u8(gio["images"], gio["index"]), gio["normalized"][gio["index"]])
    assert out.min() == -1.0 and out.max() == 1.0            # 0 -> -1, 255 -> +1 exactly


@pytest.mark.parametrize("n", [1, 4, 7, 25, 30])
def test_image_grid_u8_matches_save_image(gio, n):
    g = gio["grids"][n]
    grid = O.image_grid_u8(O.to_unit_range(g["x"]), nrow=int(math.sqrt(n)))
    assert grid.dtype == torch.uint8 and torch.equal(grid, g["grid"])


def test_live_torchvision_agrees_when_installed(gio):
    tv = pytest.importorskip("torchvision")
    x01 = O.to_unit_range(gio["grids"][25]["x"])
    live = tv.utils.make_grid(x01, nrow=5).mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8)
    assert torch.equal(live, O.image_grid_u8(x01, nrow=5))


def test_png_bytes_equal_the_reference_file(gio):
    """The in-memory PIL encode of the uint8 grid reproduces save_image's file byte for byte."""
    pytest.importorskip("PIL")
    from tinydiffusionmodels_b200.mnist import _encode_png

    g = gio["grids"][25]
    data = _encode_png(g["grid"].numpy())
    if data != bytes(g["png"].numpy()):       # a different zlib/PIL build may compress differently: compare pixels
        from PIL import Image

        assert np.array_equal(np.array(Image.open(io.BytesIO(data))), g["grid"].numpy())


def test_grid_shape_is_host_only():
    from tinydiffusionmodels_b200 import ops

    assert ops.image_grid_shape(25, 28, 28, 5) == (152, 152)
    assert ops.image_grid_shape(30, 28, 28, 5) == (182, 152)
    assert ops.image_grid_shape(7, 28, 28, 2) == (122, 62)
    assert ops.image_grid_shape(1, 28, 28, 1) == (28, 28)
    assert ops.image_grid_shape(3, 28, 28, 8, padding=0) == (28, 84)


def test_device_pipeline_refuses_cpu():
    from tinydiffusionmodels_b200 import _lib, ops
    from tinydiffusionmodels_b200.data import DeviceImages

    imgs = torch.zeros(4, 28, 28, dtype=torch.uint8)
    with pytest.raises(_lib.TdmError):
        ops.normalize_u8(imgs)
    with pytest.raises(_lib.TdmError):
        ops.image_grid_u8(torch.zeros(4, 1, 28, 28))
    with pytest.raises(_lib.TdmError):
        DeviceImages(imgs, "cpu")


def test_checkpoint_round_trip_with_reference_format(tmp_path):
    """src/mnist.py:165 writes torch.save(model.state_dict()); src/utils.py load_checkpoint reads it back.  A file
    holding the reference's keys/shapes loads into our module and what we save has the same keys, shapes, dtypes
    and values (mnist_golden.pt carries a state_dict produced by the real reference class)."""
    from tinydiffusionmodels_b200.mnist import SimpleUNet
    from tinydiffusionmodels_b200.utils import load_checkpoint, save_checkpoint

    ref_sd = torch.load(GOLD / "mnist_golden.pt", weights_only=False)["state_dict"]
    src = tmp_path / "ref.pth"
    torch.save(ref_sd, src)
    m = SimpleUNet()
    m.load_state_dict(load_checkpoint(str(src), "cpu"))
    dst = tmp_path / "ours.pth"
    save_checkpoint(m.state_dict(), str(dst))
    back = torch.load(dst, map_location="cpu", weights_only=True)
    assert list(back.keys()) == list(ref_sd.keys())
    for k, v in ref_sd.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and torch.equal(back[k], v), k


def test_text_checkpoint_round_trip_with_reference_format(tmp_path):
    """src/shakespeare.py:311-341 writes {'diffusion_model', 'rounding_fn', ['embedding_fn'], 'epoch', ...}; the
    state_dicts in text_golden.pt come from the real reference classes.  Such a file loads into our modules and
    the same dictionary layout written from our modules carries identical keys, shapes, dtypes and values."""
    from tinydiffusionmodels_b200.shakespeare import LearnedEmbedding, LearnedRounding, TinyTransformer
    from tinydiffusionmodels_b200.utils import load_checkpoint, save_checkpoint

    gt = torch.load(GOLD / "text_golden.pt", weights_only=False)
    ref_ckpt = {"diffusion_model": gt["model_sd"], "rounding_fn": gt["rounding_sd"], "embedding_fn": gt["emb_sd"],
                "epoch": 3, "final_training": True}
    src = tmp_path / "text_ref.pth"
    torch.save(ref_ckpt, src)
    ck = load_checkpoint(str(src), "cpu")
    model, rnd, emb = TinyTransformer(gt["dim"]), LearnedRounding(gt["dim"], gt["V"]), LearnedEmbedding(gt["V"], gt["dim"])
    model.load_state_dict(ck["diffusion_model"])
    rnd.load_state_dict(ck["rounding_fn"])
    emb.load_state_dict(ck["embedding_fn"])
    dst = tmp_path / "text_ours.pth"
    save_checkpoint({"diffusion_model": model.state_dict(), "rounding_fn": rnd.state_dict(),
                     "embedding_fn": emb.state_dict(), "epoch": 3, "final_training": True}, str(dst))
    back = torch.load(dst, map_location="cpu", weights_only=True)   # plain tensors only: nothing of the module is pickled
    assert list(back) == list(ref_ckpt) and back["epoch"] == 3 and back["final_training"] is True
    for part in ("diffusion_model", "rounding_fn", "embedding_fn"):
        assert list(back[part]) == list(ref_ckpt[part])
        for k, v in ref_ckpt[part].items():
            assert back[part][k].dtype == v.dtype and back[part][k].shape == v.shape and torch.equal(back[part][k], v), k
