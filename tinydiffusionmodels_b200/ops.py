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


def normalize_u8(images: torch.Tensor, index: torch.Tensor | None = None, mean: float = 0.5,
                 std: float = 0.5, out: torch.Tensor | None = None, *, check_index: bool = True) -> torch.Tensor:
    """Gather rows of a device-resident uint8 image set and apply the reference's input transform,
    ToTensor + Normalize((mean,), (std,)) (src/mnist.py:141-144), bit-identically.

    ``images`` (N, H, W) or (N, 1, H, W) uint8 with H*W % 4 == 0; ``index`` int64 rows to take (``None`` =
    all rows in order; range-checked on the host unless ``check_index=False``, which avoids the device
    sync when the caller built the indices itself).  Returns (n, 1, H, W) fp32.
    """
    _need_cuda(images, index, out)
    if images.dtype != torch.uint8:
        raise ValueError("normalize_u8 expects a uint8 image tensor")
    if images.dim() == 4 and images.shape[1] == 1:
        images = images[:, 0]
    if images.dim() != 3:
        raise ValueError("normalize_u8 expects (N, H, W) or (N, 1, H, W) images")
    images = images.contiguous()
    total, h, w = images.shape
    if index is not None:
        index = _t64(index)
        if check_index and index.numel() and (int(index.min()) < 0 or int(index.max()) >= total):
            raise IndexError("normalize_u8: index out of range")
    n = total if index is None else index.numel()
    if out is None:
        out = torch.empty((n, 1, h, w), dtype=torch.float32, device=images.device)
    elif out.dtype != torch.float32 or out.numel() != n * h * w or not out.is_contiguous():
        raise ValueError("normalize_u8: out must be a contiguous fp32 tensor of n*H*W elements")
    _lib.check(_lib.load().tdm_u8_gather_normalize(images.data_ptr(), _lib.ptr(index), out.data_ptr(), n, h * w,
                                                   float(mean), float(std), _lib.stream_ptr(images.device)),
               "tdm_u8_gather_normalize")
    return out


def image_grid_shape(n: int, h: int, w: int, nrow: int, padding: int = 2) -> tuple[int, int]:
    """(height, width) of the grid torchvision's make_grid builds for n single-channel h x w images."""
    import ctypes

    gh, gw = ctypes.c_int(), ctypes.c_int()
    _lib.check(_lib.load().tdm_image_grid_shape(n, h, w, nrow, padding, ctypes.byref(gh), ctypes.byref(gw)),
               "tdm_image_grid_shape")
    return gh.value, gw.value


def image_grid_u8(x: torch.Tensor, nrow: int = 8, padding: int = 2, *, from_signed: bool = False) -> torch.Tensor:
    """The uint8 HWC array ``torchvision.utils.save_image(x, nrow=nrow, padding=padding)`` hands to PIL
    (src/mnist.py:116-119,194-199), built on the device.  ``x`` (n, 1, H, W) fp32 in [0, 1]; with
    ``from_signed`` the map (clamp(x,-1,1)+1)/2 of src/mnist.py:194 is applied first."""
    _need_cuda(x)
    if x.dim() != 4 or x.shape[1] != 1 or x.shape[0] < 1:
        raise ValueError("image_grid_u8 expects a non-empty (n, 1, H, W) tensor")
    x = _f32c(x)
    n, _, h, w = x.shape
    gh, gw = image_grid_shape(n, h, w, nrow, padding)
    grid = torch.empty((gh, gw, 3), dtype=torch.uint8, device=x.device)
    _lib.check(_lib.load().tdm_image_grid_u8(x.data_ptr(), grid.data_ptr(), n, h, w, nrow, padding,
                                             1 if from_signed else 0, _lib.stream_ptr(x.device)),
               "tdm_image_grid_u8")
    return grid
