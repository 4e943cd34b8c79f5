"""Device-side state of the Shakespeare sampler: packed TinyTransformer weights, the token-state
workspace, and the rounding matrices (learned nn.Linear(dim, V) or the normalised embedding table)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .schedule import schedule_on

FF = 2048  # nn.TransformerEncoderLayer default dim_feedforward (src/shakespeare.py:108-110)


def _pad256(n: int) -> int:
    return (n + 255) // 256 * 256


def pack_linear(w: torch.Tensor, n_padded: int | None = None) -> torch.Tensor:
    """nn.Linear.weight (N, K) fp32 on CUDA -> bf16 planes [K/8][Np][8] as a uint8 tensor."""
    if not w.is_cuda:
        raise _lib.TdmError("pack_linear needs a CUDA tensor (no CPU fallback)")
    w = w.detach().float().contiguous()
    n, k = w.shape
    npad = _pad256(n) if n_padded is None else n_padded
    out = torch.empty((k // 8) * npad * 16, dtype=torch.uint8, device=w.device)
    _lib.check(_lib.load().tdm_pack_linear(w.data_ptr(), n, k, npad, out.data_ptr(), _lib.stream_ptr(w.device)),
               "tdm_pack_linear")
    return out


class TextEngine:
    """TinyTransformer (src/shakespeare.py:105-120) as packed weights + workspace for (B, L)."""

    def __init__(self, state_dict: dict, device, batch: int, seq_len: int):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TdmError("TextEngine needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        sd = {k: v.detach().to(self.device, torch.float32).contiguous() for k, v in state_dict.items()}
        self.dim = sd["time_emb.weight"].shape[0]
        self.depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.layers."))
        self.batch, self.seq_len = int(batch), int(seq_len)
        keep = []
        ptrs = []
        for i in range(self.depth):
            p = f"encoder.layers.{i}."
            if sd[p + "linear1.weight"].shape[0] != FF:
                raise _lib.TdmError("only dim_feedforward=2048 (the reference's default) is supported")
            if self.dim == 256:
                # the fused feed-forward kernel streams its weights stage by stage: one contiguous copy per ring slot
                w1p = torch.empty(FF * 256 * 2, dtype=torch.uint8, device=self.device)
                w2p = torch.empty(FF * 256 * 2, dtype=torch.uint8, device=self.device)
                _lib.check(self.lib.tdm_pack_ffn_weights(sd[p + "linear1.weight"].data_ptr(), sd[p + "linear2.weight"].data_ptr(),
                                                         w1p.data_ptr(), w2p.data_ptr(), _lib.stream_ptr(self.device)),
                           "tdm_pack_ffn_weights")
            else:
                w1p, w2p = pack_linear(sd[p + "linear1.weight"]), pack_linear(sd[p + "linear2.weight"])
            items = [pack_linear(sd[p + "self_attn.in_proj_weight"]), sd[p + "self_attn.in_proj_bias"],
                     pack_linear(sd[p + "self_attn.out_proj.weight"]), sd[p + "self_attn.out_proj.bias"],
                     w1p, sd[p + "linear1.bias"], w2p, sd[p + "linear2.bias"],
                     sd[p + "norm1.weight"], sd[p + "norm1.bias"], sd[p + "norm2.weight"], sd[p + "norm2.bias"]]
            keep += items
            ptrs += [t.data_ptr() for t in items]
        self.time_w = sd["time_emb.weight"].reshape(-1).contiguous()
        self.time_b = sd["time_emb.bias"].contiguous()
        keep += [self.time_w, self.time_b]
        ptrs += [self.time_w.data_ptr(), self.time_b.data_ptr()]
        self._keep = keep
        self.ptrs = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self.ws_bytes = int(self.lib.tdm_text_workspace_bytes(self.batch, self.seq_len, self.dim))
        if self.ws_bytes <= 0:
            raise _lib.TdmError("bad text workspace shape")
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self.sched = schedule_on(self.device)

    def _st(self):
        return _lib.stream_ptr(self.device)

    def load_state(self, x: torch.Tensor, t: torch.Tensor) -> None:
        x = x.float().contiguous()
        assert x.shape == (self.batch, self.seq_len, self.dim), (x.shape, self.batch, self.seq_len, self.dim)
        _lib.check(self.lib.tdm_text_load_state(x.data_ptr(), t.data_ptr(), self.time_w.data_ptr(),
                                                self.time_b.data_ptr(), self.ws.data_ptr(), self.ws_bytes, self.batch,
                                                self.seq_len, self.dim, self._st()), "tdm_text_load_state")

    def read(self, which: int, out: torch.Tensor | None = None) -> torch.Tensor:
        out = torch.empty(self.batch, self.seq_len, self.dim, device=self.device) if out is None else out
        _lib.check(self.lib.tdm_text_read(self.ws.data_ptr(), self.ws_bytes, which, out.data_ptr(), self.batch,
                                          self.seq_len, self.dim, self._st()), "tdm_text_read")
        return out

    def forward(self, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        """eps = TinyTransformer(x, t)."""
        t = t.to(torch.int64).contiguous()
        self.load_state(x, t)
        _lib.check(self.lib.tdm_text_forward(self.ptrs, self.depth, self.ws.data_ptr(), self.ws_bytes, t.data_ptr(),
                                             self.batch, self.seq_len, self.dim, self._st()), "tdm_text_forward")
        return self.read(1)

    def p_sample_inplace(self, t: torch.Tensor, z: torch.Tensor | None = None, *, seed: int = 0,
                         sample_offset: int = 0, step_id: int = 0) -> None:
        """One reverse step on the loaded state (t must be int64, contiguous, on the device)."""
        s = self.sched
        _lib.check(self.lib.tdm_text_p_sample(self.ptrs, self.depth, self.ws.data_ptr(), self.ws_bytes, t.data_ptr(),
                                              _lib.ptr(z), s.betas.data_ptr(), s.alphas.data_ptr(),
                                              s.sqrt_one_minus_alphas_cumprod.data_ptr(), self.batch, self.seq_len,
                                              self.dim, seed, sample_offset, step_id, self._st()), "tdm_text_p_sample")

    def p_sample(self, x, t, z=None, **kw) -> torch.Tensor:
        t = t.to(torch.int64).contiguous()
        self.load_state(x, t)
        self.p_sample_inplace(t, None if z is None else z.float().contiguous(), **kw)
        return self.read(0)

    @torch.no_grad()
    def sample_loop(self, x: torch.Tensor, *, seed: int, sample_offset: int = 0, steps: int = 1000,
                    use_graph: bool = True) -> torch.Tensor:
        """steps reverse steps t = steps-1 .. 0 starting from x (B, L, D); returns x_0."""
        dev = self.device
        t = torch.full((self.batch,), steps - 1, device=dev, dtype=torch.int64)
        self.load_state(x, t)

        def one():
            self.p_sample_inplace(t, None, seed=seed, sample_offset=sample_offset)
            _lib.check(self.lib.tdm_timestep_advance(t.data_ptr(), self.batch, -1, self._st()), "advance")

        if not use_graph or steps < 4:
            for _ in range(steps):
                one()
            return self.read(0)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            one()                       # warm-up outside capture
        torch.cuda.current_stream(dev).wait_stream(side)
        t.fill_(steps - 1)
        self.load_state(x, t)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            one()
        for _ in range(steps):
            g.replay()
        return self.read(0)


class Rounder:
    """argmax over the vocabulary of x.W^T + b (learned) or of the cosine similarity to the embedding
    table, optionally mixed with AR logits — logits never materialised."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.lib = _lib.load()
        self._packed = {}
        self._ws = None

    def _pack(self, w: torch.Tensor, normalize: bool):
        key = (w.data_ptr(), w._version, normalize, tuple(w.shape))
        if key not in self._packed:
            self._packed.clear()
            src = w.detach().to(self.device, torch.float32)
            if normalize:
                src = torch.nn.functional.normalize(src, dim=1)   # once, not per call (src/shakespeare.py:398)
            self._packed[key] = pack_linear(src.contiguous())
        return self._packed[key]

    def argmax(self, x: torch.Tensor, *, weight: torch.Tensor, bias: torch.Tensor | None = None,
               cosine: bool = False, ar_logits: torch.Tensor | None = None, alpha: float = 0.0,
               temperature: float = 1.0, return_values: bool = False):
        if not x.is_cuda:
            raise _lib.TdmError("rounding needs CUDA tensors (no CPU fallback)")
        lead = x.shape[:-1]
        dim = x.shape[-1]
        xr = x.reshape(-1, dim).float().contiguous()
        rows = xr.shape[0]
        vocab = weight.shape[0]
        wp = self._pack(weight, cosine)
        need = int(self.lib.tdm_round_workspace_bytes(rows, dim, vocab))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        idx = torch.empty(rows, dtype=torch.int64, device=self.device)
        val = torch.empty(rows, dtype=torch.float32, device=self.device) if return_values else None
        b = None if (bias is None or cosine) else bias.detach().to(self.device, torch.float32).contiguous()
        ar = None
        if ar_logits is not None:
            ar = ar_logits.float().contiguous()
            assert ar.shape == (rows, vocab)
        _lib.check(self.lib.tdm_round_argmax(xr.data_ptr(), rows, dim, wp.data_ptr(), vocab, _pad256(vocab),
                                             _lib.ptr(b), int(cosine), _lib.ptr(ar), vocab, float(alpha),
                                             float(temperature), idx.data_ptr(), _lib.ptr(val), self._ws.data_ptr(),
                                             self._ws.numel(), _lib.stream_ptr(self.device)), "tdm_round_argmax")
        idx = idx.view(lead)
        return (idx, val.view(lead)) if return_values else idx

    def logits(self, x: torch.Tensor, *, weight: torch.Tensor, bias: torch.Tensor | None = None,
               cosine: bool = False) -> torch.Tensor:
        """x.W^T + b (or the cosine similarities) as an fp32 (..., V) tensor - what ``LearnedRounding.forward``
        returns (ref src/shakespeare.py:93-102).  bf16 tensor-core products, fp32 accumulation."""
        if not x.is_cuda:
            raise _lib.TdmError("rounding needs CUDA tensors (no CPU fallback)")
        lead = x.shape[:-1]
        dim = x.shape[-1]
        xr = x.reshape(-1, dim).float().contiguous()
        rows = xr.shape[0]
        vocab = weight.shape[0]
        if rows == 0:
            return torch.empty(*lead, vocab, dtype=torch.float32, device=self.device)
        wp = self._pack(weight, cosine)
        need = int(self.lib.tdm_round_workspace_bytes(rows, dim, vocab))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        out = torch.empty(rows, vocab, dtype=torch.float32, device=self.device)
        b = None if (bias is None or cosine) else bias.detach().to(self.device, torch.float32).contiguous()
        _lib.check(self.lib.tdm_linear_logits(xr.data_ptr(), rows, dim, wp.data_ptr(), vocab, _pad256(vocab), _lib.ptr(b),
                                              int(cosine), out.data_ptr(), vocab, self._ws.data_ptr(), self._ws.numel(),
                                              _lib.stream_ptr(self.device)), "tdm_linear_logits")
        return out.view(*lead, vocab)
