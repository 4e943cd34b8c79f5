"""Tensor-level wrappers over the elementwise C-ABI entry points.

These take torch CUDA tensors, validate what the C side cannot (dtype, device, contiguity) and
enqueue on the current stream.  No fallback: a CPU tensor is an error.
"""
from __future__ import annotations

import torch

from . import _lib
from .schedule import Schedule, schedule_on


def _need_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.TdmError("tinydiffusionmodels_b200 ops need CUDA tensors (no CPU fallback)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _t64(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.int64).contiguous()


def q_sample(x_start: torch.Tensor, t: torch.Tensor, noise: torch.Tensor | None = None,
             sched: Schedule | None = None, *, seed: int = 0, sample_offset: int = 0,
             stream_id: int = 0, return_noise: bool = False):
    """Forward diffusion (src/mnist.py:36-42; src/shakespeare.py:37-44). Any trailing shape.

    With ``noise=None`` the noise is drawn in-kernel (Philox) instead of ``randn_like``.
    """
    _need_cuda(x_start, t, noise)
    lib = _lib.load()
    sched = sched or schedule_on(x_start.device)
    x = _f32c(x_start)
    t = _t64(t)
    b = x.shape[0]
    inner = x.numel() // max(b, 1)
    out = torch.empty_like(x)
    st = _lib.stream_ptr(x.device)
    if noise is not None:
        n = _f32c(noise)
        _lib.check(lib.tdm_q_sample(x.data_ptr(), n.data_ptr(), t.data_ptr(),
                                    sched.sqrt_alphas_cumprod.data_ptr(),
                                    sched.sqrt_one_minus_alphas_cumprod.data_ptr(), out.data_ptr(),
                                    b, inner, sched.timesteps, st), "tdm_q_sample")
    else:
        n = torch.empty_like(x)
        _lib.check(lib.tdm_q_sample_philox(x.data_ptr(), t.data_ptr(),
                                           sched.sqrt_alphas_cumprod.data_ptr(),
                                           sched.sqrt_one_minus_alphas_cumprod.data_ptr(),
                                           n.data_ptr(), out.data_ptr(), b, inner, sched.timesteps,
                                           seed, sample_offset, stream_id, None, st), "tdm_q_sample_philox")
    return (out, n) if return_noise else out


def reverse_step(x: torch.Tensor, eps: torch.Tensor, t: torch.Tensor, z: torch.Tensor | None = None,
                 sched: Schedule | None = None, *, out: torch.Tensor | None = None, seed: int = 0,
                 sample_offset: int = 0, step_id: int = 0) -> torch.Tensor:
    """x_{t-1} from x_t and the predicted noise (the arithmetic of src/mnist.py:169-180)."""
    _need_cuda(x, eps, t, z)
    lib = _lib.load()
    sched = sched or schedule_on(x.device)
    x = _f32c(x)
    eps = _f32c(eps)
    t = _t64(t)
    z = None if z is None else _f32c(z)
    b = x.shape[0]
    inner = x.numel() // max(b, 1)
    out = torch.empty_like(x) if out is None else out
    _lib.check(lib.tdm_reverse_step(x.data_ptr(), eps.data_ptr(), _lib.ptr(z), t.data_ptr(),
                                    sched.betas.data_ptr(), sched.alphas.data_ptr(),
                                    sched.sqrt_one_minus_alphas_cumprod.data_ptr(), out.data_ptr(),
                                    b, inner, sched.timesteps, seed, sample_offset, step_id,
                                    _lib.stream_ptr(x.device)), "tdm_reverse_step")
    return out


def randn(shape, device, *, seed: int = 0, sample_offset: int = 0, stream_id: int = 0) -> torch.Tensor:
    """N(0,1) tensor from the library's Philox stream; row b depends only on (seed, offset+b)."""
    out = torch.empty(shape, device=device, dtype=torch.float32)
    _need_cuda(out)
    b = out.shape[0]
    inner = out.numel() // max(b, 1)
    _lib.check(_lib.load().tdm_randn_philox(out.data_ptr(), b, inner, seed, sample_offset, stream_id,
                                            _lib.stream_ptr(out.device)), "tdm_randn_philox")
    return out


def to_unit_range(x: torch.Tensor) -> torch.Tensor:
    """(clamp(x,-1,1)+1)/2 (src/mnist.py:194)."""
    _need_cuda(x)
    x = _f32c(x)
    out = torch.empty_like(x)
    _lib.check(_lib.load().tdm_to_unit_range(x.data_ptr(), out.data_ptr(), x.numel(),
                                             _lib.stream_ptr(x.device)), "tdm_to_unit_range")
    return out
