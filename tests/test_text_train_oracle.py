"""Pins oracle/text_train_oracle.py (the CPU restatement of one Shakespeare training step) to the reference's own
outputs recorded in tests/golden/text_train_golden.pt, and its dropout / timestep / noise generators to their stated
properties.  CPU only."""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import text_train_oracle as TO

GOLD = Path(__file__).resolve().parent / "golden" / "text_train_golden.pt"
sys.path.insert(0, str(GOLD.parent))
from make_golden_text_train import sample_index  # noqa: E402

TAB = O.make_tables()


def _check(digest, full, rtol, what):
    flat = full.detach().reshape(-1)
    got = flat[sample_index(flat.numel())]
    scale = max(digest["norm"] / max(flat.numel(), 1) ** 0.5, 1e-12)
    assert abs(flat.double().norm().item() - digest["norm"]) <= rtol * max(digest["norm"], 1e-12), what
    assert float((got - digest["sample"]).abs().max()) <= 50 * rtol * scale + 1e-9, what


def test_losses_and_gradients_equal_the_reference_step():
    g = torch.load(GOLD)
    losses, grads = TO.text_losses_and_grads(g["init"]["model"], g["init"]["rounding"]["decoder.weight"],
                                             g["init"]["rounding"]["decoder.bias"],
                                             g["init"]["embedding"]["embeddings.weight"], g["ids"], g["t"], g["noise"],
                                             TAB, rounding_weight=g["rw"], dropout=0.0)
    assert torch.allclose(torch.stack(losses), g["losses"], rtol=2e-6, atol=0)
    for k, d in g["grads"]["model"].items():
        _check(d, grads["model." + k], 2e-5, k)
    _check(g["grads"]["rounding"]["decoder.weight"], grads["decoder.weight"], 2e-5, "decoder.weight")
    _check(g["grads"]["rounding"]["decoder.bias"], grads["decoder.bias"], 2e-5, "decoder.bias")
    _check(g["grads"]["embedding"]["embeddings.weight"], grads["embeddings.weight"], 2e-5, "embeddings.weight")


def test_adamw_update_equals_torch_optim_on_the_reference_step():
    g = torch.load(GOLD)
    _, grads = TO.text_losses_and_grads(g["init"]["model"], g["init"]["rounding"]["decoder.weight"],
                                        g["init"]["rounding"]["decoder.bias"],
                                        g["init"]["embedding"]["embeddings.weight"], g["ids"], g["t"], g["noise"], TAB,
                                        rounding_weight=g["rw"], dropout=0.0)
    named = {("model." + k): v for k, v in g["init"]["model"].items()}
    named["decoder.weight"] = g["init"]["rounding"]["decoder.weight"]
    named["decoder.bias"] = g["init"]["rounding"]["decoder.bias"]
    named["embeddings.weight"] = g["init"]["embedding"]["embeddings.weight"]
    after = {**{"model." + k: v for k, v in g["after"]["model"].items()},
             "decoder.weight": g["after"]["rounding"]["decoder.weight"],
             "decoder.bias": g["after"]["rounding"]["decoder.bias"],
             "embeddings.weight": g["after"]["embedding"]["embeddings.weight"]}
    for k, d in after.items():
        p, _, _ = O.adamw_step(named[k], grads[k], torch.zeros_like(named[k]), torch.zeros_like(named[k]), 1,
                               lr=g["lr"], wd=g["wd"])
        # the first AdamW step moves every element by lr * g / (|g| + eps): where the gradient is rounding noise around
        # zero (the key bias of softmax attention) that quotient amplifies 1e-10 differences, hence 2 % of lr absolute
        flat = p.reshape(-1)
        assert float((flat[sample_index(flat.numel())] - d["sample"]).abs().max()) <= 2e-2 * g["lr"], k
        assert abs(flat.double().norm().item() - d["norm"]) <= 1e-5 * max(d["norm"], 1e-12), k


def test_dropout_masks_are_a_pure_function_with_the_stated_rate():
    a = TO.keep_mask((4, 64, 256), 0.1, seed=7, step=3, site=18)
    b = TO.keep_mask((4, 64, 256), 0.1, seed=7, step=3, site=18)
    assert torch.equal(a, b)
    assert set(a.unique().tolist()) == {0.0, float(np.float32(1.0) / (np.float32(1.0) - np.float32(0.1)))}
    assert abs(float((a == 0).float().mean()) - 0.1) < 5e-3
    for other in (TO.keep_mask((4, 64, 256), 0.1, 7, 4, 18), TO.keep_mask((4, 64, 256), 0.1, 7, 3, 19),
                  TO.keep_mask((4, 64, 256), 0.1, 8, 3, 18)):
        assert not torch.equal(a, other)       # step, site and seed all enter the counter / key
    assert torch.equal(TO.keep_mask((2, 8), 0.0, 1, 1, 1), torch.ones(2, 8))


def test_timesteps_are_uniform_over_the_schedule():
    t = TO.draw_t(20000, seed=5, step=1)
    assert t.min() >= 0 and t.max() <= 999
    hist = torch.bincount(t // 100, minlength=10).float() / 20000
    assert float((hist - 0.1).abs().max()) < 0.01
    assert torch.equal(TO.draw_t(8, 5, 1, sample_offset=4), TO.draw_t(12, 5, 1)[4:])   # keyed on the global sequence index


def test_dropout_restatement_is_the_mask_free_one_times_the_masks():
    """With p > 0 the train-mode forward differs from the mask-free one, reduces to it at p = 0 and is differentiable
    through the masks (expected value of a dropped activation equals the activation)."""
    g = torch.load(GOLD)
    sd = g["init"]["model"]
    x = torch.randn(2, 64, 256, generator=torch.Generator().manual_seed(0))
    t = torch.tensor([10, 900])
    a = TO.transformer_forward_train(sd, x, t, 0.0, 1, 1)
    assert torch.allclose(a, O.transformer_forward(sd, x, t), atol=1e-6)
    b = TO.transformer_forward_train(sd, x, t, 0.1, 1, 1)
    assert not torch.allclose(a, b, atol=1e-3)
    assert torch.allclose(a, g["pred"], atol=10) and a.shape == g["pred"].shape


@pytest.mark.skipif(not Path("/root/reference/src/shakespeare.py").exists(), reason="reference not present")
def test_live_reference_train_mode_forward_without_dropout():
    sys.path.insert(0, str(GOLD.parent))
    from make_golden import import_reference

    _, ref = import_reference()
    torch.manual_seed(1)
    m = ref.TinyTransformer(256, depth=2, dropout=0.0).train()
    x = torch.randn(2, 64, 256)
    t = torch.tensor([3, 700])
    want = m(x, t).detach()
    got = TO.transformer_forward_train({k: v for k, v in m.state_dict().items()}, x, t, 0.0, 0, 0)
    assert torch.allclose(got, want, atol=2e-5)
