"""ctypes binding of libtdm_b200.so (include/tdm_b200.h).

This is the only place Python touches the C ABI.  There is deliberately no fallback: if the
shared object is missing or a call fails, the caller gets an exception — the product path never
silently degrades to PyTorch or CPU code.
"""
from __future__ import annotations

import ctypes
import re
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libtdm_b200.so"
HEADER_PATH = _PKG.parent / "include" / "tdm_b200.h"

_P = c_void_p  # every device pointer crosses the ABI as an opaque address

# name -> (restype, [argtypes]) ; must mirror include/tdm_b200.h (tests/test_abi.py checks it)
SIGNATURES: dict[str, tuple] = {
    "tdm_version": (c_int, []),
    "tdm_last_error": (c_char_p, []),
    "tdm_launch_count": (c_int64, []),
    "tdm_q_sample": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, _P]),
    "tdm_q_sample_philox": (c_int, [_P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_uint64, c_uint64, c_uint32, _P, _P]),
    "tdm_reverse_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_uint64, c_uint64, c_uint32, _P]),
    "tdm_timestep_advance": (c_int, [_P, c_int64, c_int64, _P]),
    "tdm_randn_philox": (c_int, [_P, c_int64, c_int64, c_uint64, c_uint64, c_uint32, _P]),
    "tdm_to_unit_range": (c_int, [_P, _P, c_int64, _P]),
    "tdm_u8_gather_normalize": (c_int, [_P, _P, _P, c_int64, c_int64, c_float, c_float, _P]),
    "tdm_image_grid_shape": (c_int, [c_int64, c_int, c_int, c_int, c_int, _P, _P]),
    "tdm_image_grid_u8": (c_int, [_P, _P, c_int64, c_int, c_int, c_int, c_int, c_int, _P]),
    "tdm_unet_param_count": (c_int64, []),
    "tdm_unet_wpack_bytes": (c_int64, []),
    "tdm_unet_workspace_bytes": (c_int64, [c_int64, c_int]),
    "tdm_unet_debug_layout": (c_int, [c_int64, ctypes.POINTER(c_int64)]),
    "tdm_unet_pack_weights": (c_int, [_P, _P, _P]),
    "tdm_unet_pack_weights_host": (c_int, [_P, _P, _P, _P]),
    "tdm_unet_forget_host_params": (c_int, [_P]),
    "tdm_unet_set_fused": (c_int, [c_int]),
    "tdm_unet_forward": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, _P]),
    "tdm_unet_p_sample": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_uint64, c_uint64, c_uint32, _P]),
    "tdm_unet_forward_train": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, _P]),
    "tdm_unet_backward": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, _P]),
    "tdm_adamw_flat": (c_int, [_P, _P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, _P, _P]),
    "tdm_peer_buffer_bytes": (c_int64, [c_int64]),
    "tdm_peer_grad_offset": (c_int64, [c_int64, c_int]),
    "tdm_peer_alloc": (c_int, [c_int64, ctypes.POINTER(_P)]),
    "tdm_peer_free": (c_int, [_P]),
    "tdm_peer_export": (c_int, [_P, _P]),
    "tdm_peer_import": (c_int, [_P, ctypes.POINTER(_P)]),
    "tdm_peer_close": (c_int, [_P]),
    "tdm_adamw_flat_peer": (c_int, [_P, _P, _P, c_int64, c_float, c_float, c_float, c_float, c_float, c_float, _P,
                                    ctypes.POINTER(_P), c_int, c_int, _P]),
    "tdm_pack_linear": (c_int, [_P, c_int, c_int, c_int, _P, _P]),
    "tdm_pack_ffn_weights": (c_int, [_P, _P, _P, _P, _P]),
    "tdm_text_workspace_bytes": (c_int64, [c_int64, c_int, c_int]),
    "tdm_text_load_state": (c_int, [_P, _P, _P, _P, _P, c_int64, c_int64, c_int, c_int, _P]),
    "tdm_text_read": (c_int, [_P, c_int64, c_int, _P, c_int64, c_int, c_int, _P]),
    "tdm_text_forward": (c_int, [ctypes.POINTER(_P), c_int, _P, c_int64, _P, c_int64, c_int, c_int, _P]),
    "tdm_text_p_sample": (c_int, [ctypes.POINTER(_P), c_int, _P, c_int64, _P, _P, _P, _P, _P, c_int64, c_int, c_int, c_uint64, c_uint64, c_uint32, _P]),
    "tdm_round_workspace_bytes": (c_int64, [c_int64, c_int, c_int64]),
    "tdm_round_argmax": (c_int, [_P, c_int64, c_int, _P, c_int64, c_int64, _P, c_int, _P, c_int64, c_float, c_float, _P, _P, _P, c_int64, _P]),
    "tdm_linear_logits": (c_int, [_P, c_int64, c_int, _P, c_int64, c_int64, _P, c_int, _P, c_int64, _P, c_int64, _P]),
    "tdm_embedding_gather": (c_int, [_P, c_int64, c_int, _P, c_int64, _P, _P, _P]),
    "tdm_text_train_workspace_bytes": (c_int64, [c_int64, c_int, c_int, c_int, c_int64]),
    "tdm_text_train_wpack_bytes": (c_int64, [c_int, c_int, c_int64]),
    "tdm_text_train_pack": (c_int, [_P, ctypes.POINTER(c_int64), c_int, c_int, c_int64, _P, c_int64, _P]),
    "tdm_text_train_step": (c_int, [_P, _P, ctypes.POINTER(c_int64), _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_int,
                                    c_int, c_int, c_int64, c_float, _P, c_uint64, c_uint64, _P, _P, _P]),
    "tdm_adamw_flat_lr": (c_int, [_P, _P, _P, _P, c_int64, _P, c_double, c_double, c_float, c_float, c_float, _P, _P]),
    "tdm_text_train_debug_layout": (c_int, [c_int64, c_int, c_int, c_int, c_int64, ctypes.POINTER(c_int64)]),
    "tdm_unet_profile_p_sample": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64, c_uint64, ctypes.POINTER(c_float), _P]),
}


class TdmError(RuntimeError):
    """A libtdm_b200 entry point returned non-zero."""


_lib: ctypes.CDLL | None = None


def header_symbols() -> list[str]:
    """Every function name declared in include/tdm_b200.h."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tdm_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    """Load the shared library (building nothing). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TdmError(
            f"{LIB_PATH} is missing: run `python -m tinydiffusionmodels_b200.build` "
            "(or __graft_entry__.build()). There is no CPU/PyTorch fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().tdm_last_error().decode(errors="replace")
        raise TdmError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> int | None:
    """Device address of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    return int(load().tdm_launch_count())
