"""Device-side state of the MNIST UNet: flat fp32 parameters, packed bf16 weights, workspace.

The kernels consume one *flat* fp32 parameter vector in the reference's ``state_dict`` order
 — the same buffer the optimizer and the gradient all-reduce operate on — plus a
packed bf16 image of the convolution weights that ``tdm_unet_pack_weights`` derives from it.
"""
from __future__ import annotations

import torch

from . import _lib
from .schedule import Schedule, schedule_on

# (name, shape) in reference state_dict order (src/mnist.py:45-87)
def _param_spec() -> list[tuple[str, tuple[int, ...]]]:
    spec: list[tuple[str, tuple[int, ...]]] = []
    for name, cin, cout in (("rb1", 1, 32), ("rb2", 32, 64), ("rb3", 64, 64), ("rb4", 96, 32)):
        spec += [
            (f"{name}.conv1.weight", (cout, cin, 3, 3)), (f"{name}.conv1.bias", (cout,)),
            (f"{name}.conv2.weight", (cout, cout, 3, 3)), (f"{name}.conv2.bias", (cout,)),
            (f"{name}.time_emb.weight", (cout, 1)), (f"{name}.time_emb.bias", (cout,)),
        ]
        if cin != cout:
            spec += [(f"{name}.skip.weight", (cout, cin, 1, 1)), (f"{name}.skip.bias", (cout,))]
    spec += [("out.weight", (1, 32, 1, 1)), ("out.bias", (1,))]
    return spec


PARAM_SPEC = _param_spec()
PARAM_COUNT = sum(int(torch.Size(s).numel()) for _, s in PARAM_SPEC)
assert PARAM_COUNT == 181_473


def flatten_state_dict(sd: dict, device=None) -> torch.Tensor:
    """Reference-format state_dict -> flat fp32 vector in PARAM_SPEC order (validates shapes)."""
    parts = []
    for name, shape in PARAM_SPEC:
        if name not in sd:
            raise KeyError(f"state_dict is missing {name!r}")
        p = sd[name]
        if tuple(p.shape) != shape:
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(p.shape)}")
        parts.append(p.detach().reshape(-1).to(dtype=torch.float32, device=device))
    return torch.cat(parts)


def unflatten(flat: torch.Tensor) -> dict:
    """Views of the flat vector under the reference's state_dict keys."""
    out, off = {}, 0
    for name, shape in PARAM_SPEC:
        n = int(torch.Size(shape).numel())
        out[name] = flat[off:off + n].view(shape)
        off += n
    return out


class UNetEngine:
    """Owns the packed weights and the activation workspace for batches up to ``max_batch``."""

    def __init__(self, device, max_batch: int):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TdmError("UNetEngine needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        if self.lib.tdm_unet_param_count() != PARAM_COUNT:
            raise _lib.TdmError("library/param-spec mismatch")
        self.max_batch = int(max_batch)
        self.wpack = torch.zeros(self.lib.tdm_unet_wpack_bytes(), dtype=torch.uint8, device=self.device)
        self.ws_bytes = int(self.lib.tdm_unet_workspace_bytes(self.max_batch, 0))
        # zero-filled once: guard rows and the never-written pad positions of the concat buffer
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self._ws_batch = None  # the plane strides depend on the batch: re-zero when it changes
        self.sched: Schedule = schedule_on(self.device)
        self._packed_from = None
        self._loops: dict = {}   # captured reverse-step graphs of sample_loop, keyed by (batch, seed, offset, noise?)

    # -- weights ---------------------------------------------------------------------------
    def load_flat(self, flat: torch.Tensor, key=None) -> None:
        if flat.numel() != PARAM_COUNT or flat.dtype != torch.float32 or not flat.is_cuda:
            raise ValueError("flat params must be a CUDA fp32 vector of 181,473 elements")
        flat = flat.contiguous()
        # The host mirror lets the kernels take their per-channel epilogue vectors as launch arguments
        # (constant bank) instead of shared-memory loads; weights are static while sampling.
        host = flat.detach().to("cpu", copy=True).contiguous()   # synchronises: not graph-capturable
        _lib.check(self.lib.tdm_unet_pack_weights_host(flat.data_ptr(), host.data_ptr(), self.wpack.data_ptr(),
                                                       _lib.stream_ptr(self.device)),
                   "tdm_unet_pack_weights_host")
        self._packed_from = (flat.data_ptr(), flat._version) if key is None else key
        self._loops.clear()   # captured loops carry the old per-channel vectors in their launch arguments

    def __del__(self):
        try:
            self.lib.tdm_unet_forget_host_params(self.wpack.data_ptr())
        except Exception:   # interpreter teardown
            pass

    def load_state_dict(self, sd: dict) -> None:
        self.load_flat(flatten_state_dict(sd, self.device))

    def ensure_packed(self, flat: torch.Tensor, key=None) -> None:
        """Re-pack when ``key`` changed.  ``flat._version`` alone is NOT a usable key: the fused trainer and the
        CUDA-graph replays write the flat vector through raw pointers, and ``load_state_dict`` / ``torch.optim``
        write through the parameter *views*, whose version counters are their own - the caller (SimpleUNet.engine)
        folds those and its own weight-generation counter into ``key``."""
        want = (flat.data_ptr(), flat._version) if key is None else key
        if self._packed_from != want:
            self.load_flat(flat, key)

    # -- compute ---------------------------------------------------------------------------
    def _prep(self, batch: int) -> None:
        if batch > self.max_batch:
            raise ValueError(f"batch {batch} exceeds engine capacity {self.max_batch}")
        if self._ws_batch != batch:
            # a different batch lays the planes out differently; stale data would sit in pad slots
            if self._ws_batch is not None:
                self.ws.zero_()
            self._ws_batch = batch

    def forward(self, x: torch.Tensor, t: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """eps = SimpleUNet(x, t) (src/mnist.py:76-87). x (B,1,28,28) fp32, t (B,) int64."""
        b = x.shape[0]
        if b == 0:
            return torch.empty_like(x) if out is None else out
        self._prep(b)
        x = x.contiguous()
        t = t.to(torch.int64).contiguous()
        out = torch.empty_like(x) if out is None else out
        _lib.check(self.lib.tdm_unet_forward(self.wpack.data_ptr(), x.data_ptr(), t.data_ptr(),
                                             out.data_ptr(), self.ws.data_ptr(), self.ws_bytes, b,
                                             _lib.stream_ptr(self.device)), "tdm_unet_forward")
        return out

    def p_sample(self, x: torch.Tensor, t: torch.Tensor, z: torch.Tensor | None = None,
                 out: torch.Tensor | None = None, *, seed: int = 0, sample_offset: int = 0,
                 step_id: int = 0) -> torch.Tensor:
        """One fused reverse step (src/mnist.py:167-180); ``out`` may be ``x`` itself."""
        b = x.shape[0]
        if b == 0:
            return torch.empty_like(x) if out is None else out
        self._prep(b)
        x = x.contiguous()
        t = t.to(torch.int64).contiguous()
        out = torch.empty_like(x) if out is None else out
        s = self.sched
        _lib.check(self.lib.tdm_unet_p_sample(
            self.wpack.data_ptr(), x.data_ptr(), t.data_ptr(), _lib.ptr(z), s.betas.data_ptr(),
            s.alphas.data_ptr(), s.sqrt_one_minus_alphas_cumprod.data_ptr(), out.data_ptr(),
            self.ws.data_ptr(), self.ws_bytes, b, s.timesteps, seed, sample_offset, step_id,
            _lib.stream_ptr(self.device)), "tdm_unet_p_sample")
        return out

    KERNEL_NAMES = ("rb1_conv1", "rb1_conv2", "avgpool", "rb2_conv1", "rb2_conv2", "rb3_conv1",
                    "rb3_conv2", "rb4_conv1", "rb4_conv2_out_step")
    # algorithmic FLOPs per image of each launch (1x1 skips counted with the
    # conv1 kernel that computes them, the out conv with rb4.conv2)
    KERNEL_FLOPS_PER_IMAGE = (451_584, 14_450_688 + 50_176, 0, 7_225_344 + 802_816, 14_450_688, 14_450_688,
                              14_450_688, 43_352_064 + 4_816_896, 14_450_688 + 50_176)

    def fused(self) -> bool:
        """Whether sampling launches run the 28x28 residual blocks as single kernels (tdm_unet_set_fused)."""
        prev = int(self.lib.tdm_unet_set_fused(1))
        self.lib.tdm_unet_set_fused(prev)
        return bool(prev)

    def kernel_table(self) -> list[tuple[str, int]]:
        """(name, algorithmic FLOPs per image) of every launch of one sampling step, in launch order."""
        f = dict(zip(self.KERNEL_NAMES, self.KERNEL_FLOPS_PER_IMAGE))
        if not self.fused():
            return list(f.items())
        return [("rb1_fused", f["rb1_conv1"] + f["rb1_conv2"]), ("avgpool", 0), ("rb2_conv1", f["rb2_conv1"]),
                ("rb2_conv2", f["rb2_conv2"]), ("rb3_conv1", f["rb3_conv1"]), ("rb3_conv2", f["rb3_conv2"]),
                ("rb4_fused_out_step", f["rb4_conv1"] + f["rb4_conv2_out_step"])]

    def profile_kernels(self, x: torch.Tensor, t: torch.Tensor, seed: int = 0) -> list[tuple[str, int, float]]:
        """(name, FLOPs per image, milliseconds) per launch of one fused p_sample (CUDA events between the launches)."""
        ms = self.profile_p_sample(x, t, seed)
        if self.fused():   # the event slots of the launches a fused block absorbed are empty
            ms = [ms[0], ms[2], ms[3], ms[4], ms[5], ms[6], ms[7]]
        return [(n, fl, v) for (n, fl), v in zip(self.kernel_table(), ms)]

    def profile_p_sample(self, x: torch.Tensor, t: torch.Tensor, seed: int = 0) -> list[float]:
        """Per-kernel milliseconds of one fused p_sample (CUDA events between the launches)."""
        import ctypes

        b = x.shape[0]
        self._prep(b)
        ms = (ctypes.c_float * 9)()
        s = self.sched
        _lib.check(self.lib.tdm_unet_profile_p_sample(
            self.wpack.data_ptr(), x.data_ptr(), t.data_ptr(), s.betas.data_ptr(), s.alphas.data_ptr(),
            s.sqrt_one_minus_alphas_cumprod.data_ptr(), x.data_ptr(), self.ws.data_ptr(), self.ws_bytes,
            b, seed, ms, _lib.stream_ptr(self.device)), "tdm_unet_profile_p_sample")
        return list(ms)


# ---------------------------------------------------------------------------------------------
# test/debug aid: read an intermediate activation back out of the plane layout
# ---------------------------------------------------------------------------------------------
_WS_NAMES = ("t1", "cat", "p1", "t2", "s2", "h2", "t3", "t4", "s4")
_WS_GEOM = {"t1": (28, 32), "cat": (28, 96), "p1": (14, 32), "t2": (14, 64), "s2": (14, 64),
            "h2": (14, 64), "t3": (14, 64), "t4": (28, 32), "s4": (28, 32), "h3": (14, 64)}


def read_activation(engine: UNetEngine, name: str, batch: int) -> torch.Tensor:
    """Decode workspace buffer ``name`` into a (B, C, H, W) fp32 tensor (pads dropped)."""
    import ctypes

    arr = (ctypes.c_int64 * 16)()
    _lib.check(engine.lib.tdm_unet_debug_layout(batch, arr), "tdm_unet_debug_layout")
    v = list(arr)
    ps = {28: v[2], 14: v[3]}
    off = {**dict(zip(_WS_NAMES, v[4:13])), "h3": v[14]}[name]
    w, c = _WS_GEOM[name]
    halo = v[15] if w == 28 else 24       # GUARD rows in front of position 0 (unet_layout.cuh)
    wp, s = w + 1, (w + 1) * (w + 1)
    rows = ps[w] // 16
    raw = engine.ws[off: off + (c // 8) * ps[w]].view(torch.bfloat16).view(c // 8, rows, 8)
    b = torch.arange(batch, device=raw.device).view(-1, 1, 1)
    y = torch.arange(w, device=raw.device).view(1, -1, 1)
    x = torch.arange(w, device=raw.device).view(1, 1, -1)
    pos = (b * s + (y + 1) * wp + x + halo).reshape(-1)
    g = raw[:, pos, :]                                  # (C/8, B*H*W, 8)
    return g.permute(1, 0, 2).reshape(batch, w, w, c).permute(0, 3, 1, 2).float().contiguous()
