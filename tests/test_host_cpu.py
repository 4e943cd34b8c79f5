"""Host-side logic that needs no GPU: schedule tables, state_dict ABI, flat-parameter views, errors."""
from pathlib import Path

import pytest
import torch

from oracle import ddpm_oracle as O
from tinydiffusionmodels_b200 import ops
from tinydiffusionmodels_b200._lib import TdmError
from tinydiffusionmodels_b200.schedule import make_schedule
from tinydiffusionmodels_b200.unet_engine import PARAM_COUNT, PARAM_SPEC, flatten_state_dict, unflatten

GOLD = Path(__file__).resolve().parent / "golden"


def test_schedule_tables_match_oracle_and_reference():
    s = make_schedule()
    gm = torch.load(GOLD / "mnist_golden.pt", weights_only=False)
    for k, v in O.make_tables().items():
        assert torch.equal(getattr(s, k), v), k
        assert torch.equal(getattr(s, k), gm["tables"][k]), k


def test_module_level_tables_are_exposed_like_the_reference():
    import src.mnist as m
    import src.shakespeare as t
    for mod, n in ((m, "timesteps"), (t, "T")):
        assert getattr(mod, n) == 1000
        for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            assert getattr(mod, k).shape == (1000,)


def test_state_dict_abi_equals_reference():
    from src.mnist import SimpleUNet
    from src.shakespeare import LearnedEmbedding, LearnedRounding, TinyTransformer
    gm = torch.load(GOLD / "mnist_golden.pt", weights_only=False)
    gt = torch.load(GOLD / "text_golden.pt", weights_only=False)
    ours = SimpleUNet().state_dict()
    assert [(k, tuple(v.shape)) for k, v in ours.items()] == [(k, tuple(v.shape)) for k, v in gm["state_dict"].items()]
    assert [(k, tuple(v.shape)) for k, v in ours.items()] == PARAM_SPEC
    tt = TinyTransformer(gt["dim"]).state_dict()
    assert [(k, tuple(v.shape)) for k, v in tt.items()] == [(k, tuple(v.shape)) for k, v in gt["model_sd"].items()]
    assert list(LearnedRounding(gt["dim"], gt["V"]).state_dict()) == list(gt["rounding_sd"])
    assert list(LearnedEmbedding(gt["V"], gt["dim"]).state_dict()) == list(gt["emb_sd"])
    # a reference checkpoint loads
    m = SimpleUNet()
    m.load_state_dict(gm["state_dict"])


def test_flat_parameter_views_track_the_module():
    from src.mnist import SimpleUNet
    m = SimpleUNet()
    flat = m.flat_params()
    assert flat.numel() == PARAM_COUNT == 181_473
    with torch.no_grad():
        m.rb2.conv1.bias.fill_(3.0)
    assert torch.equal(unflatten(flat)["rb2.conv1.bias"], torch.full((64,), 3.0))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    assert torch.equal(flatten_state_dict(sd), flat)
    g = m.flat_grads()
    assert all(p.grad is not None and p.grad.data_ptr() >= g.data_ptr() for p in m.parameters())
    with pytest.raises(KeyError):
        flatten_state_dict({})


def test_no_cpu_fallback_anywhere():
    from src.mnist import SimpleUNet
    from src.shakespeare import TinyTransformer
    with pytest.raises(TdmError):
        ops.q_sample(torch.zeros(1, 4), torch.zeros(1, dtype=torch.long), torch.zeros(1, 4))
    with pytest.raises(TdmError):
        with torch.no_grad():
            SimpleUNet()(torch.zeros(1, 1, 28, 28), torch.zeros(1, dtype=torch.long))
    with pytest.raises(TdmError):
        with torch.no_grad():
            TinyTransformer(256).eval()(torch.zeros(1, 64, 256), torch.zeros(1, dtype=torch.long))


def test_cli_parsers_keep_reference_flags(capsys):
    import src.mnist as m
    import src.shakespeare as t
    if torch.cuda.is_available():
        pytest.skip("CLI smoke without a GPU only")
    for mod in (m, t):
        with pytest.raises(TdmError):
            mod.main([])          # no CUDA device -> loud failure, never a silent CPU path
