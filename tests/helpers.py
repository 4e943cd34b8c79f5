"""Shared helpers for the parity tests."""
import torch


def random_unet_state_dict(seed: int = 0) -> dict:
    """Random-init SimpleUNet weights with torch's default init distributions
    (conv/linear: U(+-1/sqrt(fan_in))), keyed like the reference state_dict."""
    from tinydiffusionmodels_b200.unet_engine import PARAM_SPEC

    g = torch.Generator().manual_seed(seed)
    sd = {}
    fan = {}
    for name, shape in PARAM_SPEC:
        if name.endswith("weight"):
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            fan[name.rsplit(".", 1)[0]] = fan_in
        bound = 1.0 / (fan[name.rsplit(".", 1)[0]] ** 0.5)
        sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
    return sd


def rel_rms(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-20))
