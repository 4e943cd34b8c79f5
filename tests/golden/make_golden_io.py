"""Golden vectors for the input transform and the output step.

    python tests/golden/make_golden_io.py          (authoring container; writes tests/golden/io_golden.pt)

The pixel work at src/mnist.py:141-144 (ToTensor + Normalize inside datasets.MNIST) and :116-119 / :196-199
(utils.save_image) belongs to torchvision, a third-party dependency the reference leaves unpinned; this script
calls torchvision (the version in this image, recorded in the fixture) exactly the way the reference does and
stores inputs + outputs.  tests/test_oracle.py pins oracle.normalize_u8 / oracle.image_grid_u8 to them.
"""
import io
import math
import os
import tempfile
from pathlib import Path

import numpy as np
import torch
import torchvision
from PIL import Image
from torchvision import transforms, utils

HERE = Path(__file__).resolve().parent


def reference_transform(images_u8: torch.Tensor) -> torch.Tensor:
    """What datasets.MNIST.__getitem__ + the reference's transform return for each image."""
    tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5,), (0.5,))])   # src/mnist.py:141-144
    return torch.stack([tf(Image.fromarray(img.numpy(), mode="L")) for img in images_u8])


def reference_save_image(x: torch.Tensor) -> tuple[np.ndarray, bytes]:
    """src/mnist.py:194-199: unit-range map, save_image to a temp file; returns (decoded uint8 HWC, PNG bytes)."""
    x01 = (x.clamp(-1, 1) + 1) / 2
    with tempfile.NamedTemporaryFile(suffix=".png", delete=False) as tmp:
        name = tmp.name
    try:
        utils.save_image(x01, name, nrow=int(math.sqrt(x.shape[0])))
        data = open(name, "rb").read()
    finally:
        os.unlink(name)
    return np.array(Image.open(io.BytesIO(data))), data


def main():
    g = torch.Generator().manual_seed(7)
    images = torch.randint(0, 256, (64, 28, 28), generator=g, dtype=torch.uint8)
    images[0].view(-1)[:256] = torch.arange(256, dtype=torch.uint8)      # every pixel value occurs
    index = torch.randperm(64, generator=g)[:40]
    out = {"torchvision": torchvision.__version__, "torch": torch.__version__,
           "images": images, "index": index, "normalized": reference_transform(images)}
    grids = {}
    for n in (1, 4, 7, 25, 30):
        x = torch.randn(n, 1, 28, 28, generator=g) * 0.8
        x.view(-1)[:4] = torch.tensor([-1.0, 1.0, 0.0, 1.0 - 2.0 ** -24])
        arr, png = reference_save_image(x)
        grids[n] = {"x": x, "grid": torch.from_numpy(arr.copy())}
        if n == 25:
            grids[n]["png"] = torch.frombuffer(bytearray(png), dtype=torch.uint8).clone()
    out["grids"] = grids
    torch.save(out, HERE / "io_golden.pt")
    print("wrote", HERE / "io_golden.pt", {n: tuple(v["grid"].shape) for n, v in grids.items()})


if __name__ == "__main__":
    main()
