"""The reference's command lines run end to end on the CUDA path, offline.

`python -m src.mnist --train/--sample` and `python -m src.shakespeare --sample/--guided_sample` with the additive
`--synthetic` flags (no dataset / hub downloads).  In-process calls of the same `main(argv)` the `-m` entry points use.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def in_tmp(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    return tmp_path


def test_mnist_train_then_sample_cli(cuda, in_tmp):
    from tinydiffusionmodels_b200 import mnist
    ck = str(in_tmp / "ckpt.pth")
    mnist.main(["--train", "--synthetic", "--steps_per_epoch", "6", "--epochs", "1", "--batch_size", "64", "--ckpt", ck])
    assert os.path.exists(ck)
    sd = torch.load(ck, map_location="cpu")
    assert "rb1.conv1.weight" in sd and sd["rb4.conv2.weight"].shape == (32, 32, 3, 3)   # the reference's checkpoint ABI
    mnist.main(["--sample", "--ckpt", ck, "--n_samples", "9"])
    assert (in_tmp / "samples" / "samples.png").exists()


@pytest.mark.parametrize("extra", [[], ["--use_learned_embeddings", "--embed_dim", "256"], ["--use_cosine_fallback"]])
def test_shakespeare_sample_cli_default_and_deployed_width(cuda, in_tmp, extra):
    from tinydiffusionmodels_b200 import shakespeare
    shakespeare.main(["--sample", "--synthetic", "--vocab_size", "4096", "--n", "3", *extra])
    assert (in_tmp / "samples" / "sample_2.txt").exists()


def test_shakespeare_guided_cli(cuda, in_tmp):
    from tinydiffusionmodels_b200 import shakespeare
    shakespeare.main(["--guided_sample", "--synthetic", "--vocab_size", "4096", "--n", "2", "--alpha", "0.3"])
    assert (in_tmp / "samples" / "guided_sample_1.txt").exists()


@pytest.mark.parametrize("extra", [["--use_learned_embeddings", "--embed_dim", "256"], []])
def test_shakespeare_train_then_sample_cli(cuda, in_tmp, extra):
    """`python -m src.shakespeare --train` (row f2) with learned 256-wide embeddings and with the frozen pre-trained
    table at the base LM's width (2048, the reference CLI's default), then `--sample` from the checkpoint it wrote."""
    from tinydiffusionmodels_b200 import shakespeare
    ck = str(in_tmp / "text_ckpt.pth")
    shakespeare.main(["--train", "--synthetic", "--vocab_size", "1024", "--epochs", "1", "--batch_size", "4", "--warmup_steps", "2",
                      "--seed", "3", "--ckpt", ck, *extra])
    saved = torch.load(ck, map_location="cpu")
    assert saved["final_training"] is True and ("embedding_fn" in saved) == bool(extra)
    assert all(torch.isfinite(v).all() for v in saved["diffusion_model"].values())
    shakespeare.main(["--sample", "--synthetic", "--vocab_size", "1024", "--n", "2", "--ckpt", ck, *extra])
    assert (in_tmp / "samples" / "sample_1.txt").exists()
