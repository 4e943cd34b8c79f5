"""Device-side training state of the Shakespeare model (ref src/shakespeare.py:174-341, the inner step :221-250).

``TextTrainer`` owns ONE flat fp32 buffer holding every trainable parameter of the three modules the reference
optimises together (``TinyTransformer``, ``LearnedRounding``, and ``LearnedEmbedding`` when embeddings are learned,
ref :191-194); the modules' ``nn.Parameter``s are re-pointed to views of it, so ``state_dict()`` / checkpoints keep
the reference's keys and always see the current weights.  One optimisation step is

    tdm_text_train_step   (forward, losses, backward -> flat gradient; csrc/text_train.cu)
    [all-reduce of the flat gradient when data parallel]
    tdm_adamw_flat_lr     (torch.optim.AdamW's update, learning rate read from the device)
    tdm_text_train_pack   (bf16 operand forms of the new weights)

captured once as a CUDA graph and replayed; the learning rate (cosine / warm-up, ref :159-167) and the rounding-loss
weight (ref :169-172) are device scalars the host rewrites between replays.  There is no PyTorch autograd, no
torch.optim and no CPU path here.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib
from .schedule import schedule_on

_LAYER_KEYS = ("self_attn.in_proj_weight", "self_attn.in_proj_bias", "self_attn.out_proj.weight",
               "self_attn.out_proj.bias", "linear1.weight", "linear1.bias", "linear2.weight", "linear2.bias",
               "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias")


def _pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class TextTrainer:
    def __init__(self, model, rounding_fn, embedding_fn, device, batch: int, seq_len: int, *, lr: float = 1e-4,
                 weight_decay: float = 1e-4, use_learned_embeddings: bool = True, dropout: float | None = None,
                 seed: int = 0, use_graph: bool = True, betas=(0.9, 0.999), eps: float = 1e-8):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.TdmError("text training needs a CUDA device (no CPU fallback)")
        self.lib = _lib.load()
        self.model, self.rounding_fn, self.embedding_fn = model, rounding_fn, embedding_fn
        self.learn_emb = bool(use_learned_embeddings)
        self.batch, self.seq_len = int(batch), int(seq_len)
        self.dim = int(model.dim)
        self.depth = len(model.encoder.layers)
        self.vocab = int(rounding_fn.decoder.weight.shape[0])
        self.dropout = float(model.dropout.p if dropout is None else dropout)
        self.weight_decay, self.betas, self.eps = float(weight_decay), betas, float(eps)
        self.seed = int(seed)
        self.use_graph = use_graph
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0

        # ---- flat parameter buffer, in the order of the C ABI's offset table ----
        plist = []
        for i in range(self.depth):
            lyr = model.encoder.layers[i]
            named = dict(lyr.named_parameters())
            if named["linear1.weight"].shape[0] != 2048:
                raise _lib.TdmError("only dim_feedforward=2048 (the reference's default) is supported")
            plist += [named[k] for k in _LAYER_KEYS]
        plist += [model.time_emb.weight, model.time_emb.bias, rounding_fn.decoder.weight, rounding_fn.decoder.bias]
        if self.learn_emb:
            plist.append(embedding_fn.embeddings.weight)
        offs, o = [], 0
        for p in plist:
            offs.append(o)
            o += _pad4(p.numel())
        if not self.learn_emb:
            offs.append(-1)
        self.n = o
        self.flat = torch.zeros(o, dtype=torch.float32, device=self.device)
        with torch.no_grad():
            for p, off in zip(plist, offs):
                view = self.flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.detach().to(self.device, torch.float32))
                p.data = view
        self._params = plist
        self.offsets = (ctypes.c_int64 * len(offs))(*offs)
        self.offset_list = offs
        if self.world > 1:
            dist.broadcast(self.flat, src=0)   # replicas start identical (and stay so: same gradient sum, same update)
        self.grads = torch.zeros_like(self.flat)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.emb_table = None
        if not self.learn_emb:
            tab = embedding_fn if torch.is_tensor(embedding_fn) else embedding_fn.get_embedding_matrix()
            self.emb_table = tab.detach().to(self.device, torch.float32).contiguous()

        self.wpack_bytes = int(self.lib.tdm_text_train_wpack_bytes(self.dim, self.depth, self.vocab))
        self.ws_bytes = int(self.lib.tdm_text_train_workspace_bytes(self.batch, self.seq_len, self.dim, self.depth, self.vocab))
        if self.wpack_bytes <= 0 or self.ws_bytes <= 0:
            raise _lib.TdmError(f"unsupported training shape: batch {batch} x {seq_len}, width {self.dim} "
                                f"({_lib.load().tdm_last_error().decode(errors='replace')})")
        self.wpack = torch.empty(self.wpack_bytes, dtype=torch.uint8, device=self.device)
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=self.device)
        lay = (ctypes.c_int64 * 25)()
        _lib.check(self.lib.tdm_text_train_debug_layout(self.batch, self.seq_len, self.dim, self.depth, self.vocab, lay),
                   "tdm_text_train_debug_layout")
        self._bad_flag = self.ws[lay[24]:lay[24] + 4].view(torch.int32)
        self.sched = schedule_on(self.device)
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.device)
        self.rw_dev = torch.ones(1, dtype=torch.float32, device=self.device)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.losses = torch.zeros(3, dtype=torch.float32, device=self.device)
        self.ids = torch.zeros(self.batch, self.seq_len, dtype=torch.int64, device=self.device)
        self.sample_offset = self.rank * self.batch
        self._graph = None
        self._eval_calls = 0
        self.launches_per_step = 0
        with torch.cuda.device(self.device):
            self.pack()

    # ---- pieces ----------------------------------------------------------------------------------------------
    def _st(self):
        return _lib.stream_ptr(self.device)

    def pack(self) -> None:
        _lib.check(self.lib.tdm_text_train_pack(self.flat.data_ptr(), self.offsets, self.dim, self.depth, self.vocab,
                                                self.wpack.data_ptr(), self.wpack_bytes, self._st()), "tdm_text_train_pack")

    def _objective(self, ids, grads, t=None, noise=None, dropout=None, sample_offset=None) -> None:
        s = self.sched
        p = self.dropout if dropout is None else dropout
        off = self.sample_offset if sample_offset is None else sample_offset
        _lib.check(self.lib.tdm_text_train_step(
            self.flat.data_ptr(), _lib.ptr(grads), self.offsets, self.wpack.data_ptr(), _lib.ptr(self.emb_table),
            ids.data_ptr(), _lib.ptr(t), _lib.ptr(noise), s.sqrt_alphas_cumprod.data_ptr(),
            s.sqrt_one_minus_alphas_cumprod.data_ptr(), self.ws.data_ptr(), self.ws_bytes, self.batch, self.seq_len,
            self.dim, self.depth, self.vocab, float(p), self.rw_dev.data_ptr(), self.seed, off,
            self.step_dev.data_ptr(), self.losses.data_ptr(), self._st()), "tdm_text_train_step")

    def _update(self) -> None:
        if self.world > 1:
            dist.all_reduce(self.grads)   # the step's one exchange: the flat gradient, summed over NVLink by NCCL
        _lib.check(self.lib.tdm_adamw_flat_lr(self.flat.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(),
                                              self.exp_avg_sq.data_ptr(), self.n, self.lr_dev.data_ptr(), self.betas[0],
                                              self.betas[1], self.eps, self.weight_decay, 1.0 / self.world,
                                              self.step_dev.data_ptr(), self._st()), "tdm_adamw_flat_lr")
        self.pack()

    def _check_ids(self, token_ids) -> torch.Tensor:
        ids = token_ids.to(device=self.device, dtype=torch.int64, non_blocking=True)
        if tuple(ids.shape) != (self.batch, self.seq_len):
            raise _lib.TdmError(f"token batch {tuple(ids.shape)} does not match the trainer's ({self.batch}, {self.seq_len})")
        return ids.contiguous()

    # ---- public ----------------------------------------------------------------------------------------------
    def loss_and_grads(self, token_ids, t=None, noise=None, dropout=None):
        """Forward + backward only (no update): the flat gradient is left in ``self.grads``.  ``t`` / ``noise``
        injected (parity tests) or drawn in-kernel.  Uses the current optimiser step counter for the RNG."""
        with torch.cuda.device(self.device):
            return self._loss_and_grads(token_ids, t, noise, dropout)

    def _loss_and_grads(self, token_ids, t, noise, dropout):
        ids = self._check_ids(token_ids)
        t = None if t is None else t.to(self.device, torch.int64).contiguous()
        noise = None if noise is None else noise.to(self.device, torch.float32).contiguous()
        self._objective(ids, self.grads, t, noise, dropout)
        return self.losses

    @torch.no_grad()
    def step(self, token_ids, *, lr: float | None = None, rounding_weight: float | None = None) -> torch.Tensor:
        """One optimisation step (ref :224-248).  Returns the device tensor (diffusion, rounding, total) - reading it
        synchronises, so callers that log every step pay that; callers that do not, do not."""
        with torch.cuda.device(self.device):
            return self._step(token_ids, lr, rounding_weight)

    def _step(self, token_ids, lr, rounding_weight) -> torch.Tensor:
        self.ids.copy_(self._check_ids(token_ids), non_blocking=True)
        if lr is not None:
            self.lr_dev.fill_(float(lr))
        if rounding_weight is not None:
            self.rw_dev.fill_(float(rounding_weight))
        if not self.use_graph or self.world > 1:
            self._one()
            return self.losses
        if self._graph is None:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                n0 = _lib.launch_count()
                self._one()                                   # warm-up (a real step) outside capture
                self.launches_per_step = _lib.launch_count() - n0
            torch.cuda.current_stream(self.device).wait_stream(side)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._one()
            # the capture itself did not execute: exactly one step has run so far
            return self.losses
        self._graph.replay()
        return self.losses

    def _one(self) -> None:
        _lib.check(self.lib.tdm_timestep_advance(self.step_dev.data_ptr(), 1, 1, self._st()), "step counter")
        self._objective(self.ids, self.grads)
        self._update()

    @torch.no_grad()
    def evaluate(self, token_ids) -> torch.Tensor:
        """Losses in eval mode (no dropout, no backward): the validation pass (ref :268-287).  The batch must have the
        trainer's shape (``shakespeare.train`` skips a ragged last batch)."""
        with torch.cuda.device(self.device):
            return self._evaluate(token_ids)

    def _evaluate(self, token_ids) -> torch.Tensor:
        ids = self._check_ids(token_ids)
        # the optimiser step does not advance during validation: successive batches draw their timesteps and noise from
        # successive blocks of the global sequence index instead (the reference draws fresh ones per batch, ref :276-277)
        self._eval_calls += 1
        self._objective(ids, None, sample_offset=self.sample_offset + self._eval_calls * self.batch * self.world)
        return self.losses

    def check_token_ids(self) -> None:
        """Raise IndexError if any batch since the last check held a token id outside [0, vocab) - what nn.Embedding
        does immediately (ref :226); here the kernels flag it on the device and this read synchronises, so callers
        check once per epoch rather than once per step."""
        if int(self._bad_flag.item()):
            self._bad_flag.zero_()
            raise IndexError("index out of range in the token ids of a training / validation batch")

    def sync_modules(self) -> None:
        """Invalidate the samplers' packed copies of the weights (they key on tensor versions, which raw-pointer
        updates do not bump)."""
        from . import shakespeare as S

        if hasattr(self.model, "_engines"):
            self.model._engines.clear()
        for r in S._rounders.values():
            r._packed.clear()

    def grad_of(self, index: int, shape) -> torch.Tensor:
        off = self.offset_list[index]
        n = 1
        for s in shape:
            n *= s
        return self.grads[off:off + n].view(shape)
