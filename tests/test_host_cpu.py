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


def test_text_training_schedules_known_values():
    """Host-side schedules of src/shakespeare.py:159-172 (pure functions; the text training loop itself is out of
    scope): warm-up ramp, cosine decay, eta_min floor, linear rounding-weight decay."""
    from src.shakespeare import dynamic_rounding_weight_schedule, get_cosine_schedule_with_warmup

    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=2.0)
    sched = get_cosine_schedule_with_warmup(opt, num_warmup_steps=4, num_training_steps=12, eta_min=0.1)
    lrs = []
    for _ in range(13):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
    assert lrs[:5] == [0.0, 0.5, 1.0, 1.5, 2.0]
    assert lrs[8] == pytest.approx(2.0 * 0.5) and lrs[12] == pytest.approx(2.0 * 0.1)      # midpoint, floor
    assert all(a >= b for a, b in zip(lrs[4:], lrs[5:]))
    assert dynamic_rounding_weight_schedule(0, 10) == 1.0
    assert dynamic_rounding_weight_schedule(10, 10) == pytest.approx(0.1)
    assert dynamic_rounding_weight_schedule(5, 10, 2.0, 1.0) == 1.5


def test_eval_mode_restores_the_previous_mode():
    from src.mnist import eval_mode

    m = torch.nn.Dropout()
    m.train()
    with eval_mode(m):
        assert not m.training
    assert m.training
    m.eval()
    with pytest.raises(RuntimeError):
        with eval_mode(m):
            raise RuntimeError("boom")
    assert not m.training


@pytest.mark.reference
def test_public_surface_matches_the_live_reference():
    """Every public function / class / constant the reference's two modules define exists here under the same name,
    with the same positional parameters and defaults (ours may add keyword parameters after them).  Not mirrored
    on purpose: the text training loop and its data preparation."""
    import inspect
    import math

    from tests.golden.make_golden import import_reference

    ref_m, ref_t = import_reference()
    import src.mnist as our_m
    import src.shakespeare as our_t

    not_mirrored = {"shakespeare": {"train", "load_text_dataset", "tokenize_corpus"}, "mnist": set()}
    third_party = ("torch", "tqdm", "transformers", "datasets", "math", "os", "pathlib", "typing", "google",
                   "argparse", "contextlib", "torchvision")
    for name, ref, ours in (("mnist", ref_m, our_m), ("shakespeare", ref_t, our_t)):
        for n, obj in vars(ref).items():
            if n.startswith("_") or inspect.ismodule(obj) or n in not_mirrored[name]:
                continue
            if (inspect.isfunction(obj) or inspect.isclass(obj)) and obj.__module__.split(".")[0] in third_party:
                continue
            assert hasattr(ours, n), f"src.{name}.{n} is missing"
            if inspect.isfunction(obj):
                want = [(p.name, p.default) for p in inspect.signature(obj).parameters.values()]
                got = [(p.name, p.default) for p in inspect.signature(getattr(ours, n)).parameters.values()]
                assert got[:len(want)] == want, f"src.{name}.{n}: {got} vs {want}"
                assert all(d is not inspect.Parameter.empty for _, d in got[len(want):]), n   # extras are optional
            if inspect.isclass(obj):
                for meth in ("__init__", "forward"):
                    want = [(p.name, p.default) for p in inspect.signature(getattr(obj, meth)).parameters.values()]
                    got = [(p.name, p.default) for p in inspect.signature(getattr(getattr(ours, n), meth)).parameters.values()]
                    assert got[:len(want)] == want, f"src.{name}.{n}.{meth}: {got} vs {want}"
    # the two schedule helpers agree exactly with the reference over a whole run
    for warm, total, eta in ((0, 10, 0), (5, 50, 0.0), (100, 1000, 0.05), (7, 7, 0)):
        o1 = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
        o2 = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
        s1 = ref_t.get_cosine_schedule_with_warmup(o1, warm, total, eta)
        s2 = our_t.get_cosine_schedule_with_warmup(o2, warm, total, eta)
        for _ in range(total + 3):
            assert o1.param_groups[0]["lr"] == o2.param_groups[0]["lr"]
            o1.step(); o2.step(); s1.step(); s2.step()
    for e in range(0, 11):
        assert ref_t.dynamic_rounding_weight_schedule(e, 10, 0.7, 0.05) == our_t.dynamic_rounding_weight_schedule(e, 10, 0.7, 0.05)
    assert math.isclose(our_t.dynamic_rounding_weight_schedule(3, 10), 0.73)


def test_training_learning_rates_equal_what_lambda_lr_would_set():
    """shakespeare.train writes lr * cosine_warmup_factor(k) to the device for optimiser step k; the reference steps a
    LambdaLR after every optim.step() (src/shakespeare.py:199-200, 250) - the same sequence."""
    import torch

    from src.shakespeare import cosine_warmup_factor, get_cosine_schedule_with_warmup

    lr, warm, total = 3e-4, 5, 23
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
    sched = get_cosine_schedule_with_warmup(opt, warm, total)
    for k in range(total):
        assert opt.param_groups[0]["lr"] == pytest.approx(lr * cosine_warmup_factor(k, warm, total), rel=1e-12, abs=0)
        opt.step()
        sched.step()


def test_text_training_refuses_to_run_without_cuda():
    import torch

    from src.shakespeare import LearnedEmbedding, LearnedRounding, TinyTransformer
    from tinydiffusionmodels_b200._lib import TdmError
    from tinydiffusionmodels_b200.text_train import TextTrainer

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m, r, e = TinyTransformer(256, depth=1), LearnedRounding(256, 64), LearnedEmbedding(64, 256)
    with pytest.raises(TdmError, match="no CPU fallback"):
        TextTrainer(m, r, e, "cpu", 2, 64)
