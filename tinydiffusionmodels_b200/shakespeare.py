"""Shakespeare embedding-space text diffusion on B200 — host mirror of the reference's
``src/shakespeare.py``: the *sampling* paths named by the north star and (row f2) the training loop:

* ``q_sample`` / ``p_sample``                      (ref src/shakespeare.py:37-44, 343-352)
* ``TinyTransformer.forward`` (eval)               (ref :105-120)  -> tcgen05 GEMMs + attention
* ``sample`` / ``sample_diffusion_embeddings``     (ref :355-426)  -> graph-replayed reverse loop,
                                                                      fused rounding GEMM + argmax
* ``guided_generate``                              (ref :429-470)  -> fused diff-logit GEMM + AR mix + argmax
* ``LearnedEmbedding`` / ``LearnedRounding``       (ref :46-102)   weight ABI + forward (gather kernel / logits GEMM)

* ``train`` (+ ``load_text_dataset`` / ``tokenize_corpus``) (ref :122-341) -> ``text_train.TextTrainer``: the whole
                                                                      step as one replayed CUDA graph

The modules' own ``forward`` stay inference-only (no autograd graph is ever built): training goes through ``train``,
whose backward pass is hand-written CUDA.  The base LM inside ``guided_generate`` is third party (transformers) and
stays a torch module.
"""
from __future__ import annotations

import argparse
import math
import os
from pathlib import Path

import torch
import torch.nn as nn

from . import _lib, ops
from .schedule import linear_beta_schedule, make_schedule  # noqa: F401
from .text_engine import Rounder, TextEngine
from .utils import get_samples_dir, load_checkpoint, save_samples

HF_TOKEN = os.getenv("HF_TOKEN")


# ---- host-side training schedules (ref src/shakespeare.py:159-172) ---------------------------------------------
def cosine_warmup_factor(step, num_warmup_steps, num_training_steps, eta_min=0):
    """Learning-rate factor at optimiser step ``step`` (0-based): linear ramp 0 -> 1 over the warm-up steps, then half a
    cosine down to ``eta_min`` (the lambda of ref :159-167)."""
    if step < num_warmup_steps:
        return float(step) / float(max(1, num_warmup_steps))
    progress = float(step - num_warmup_steps) / float(max(1, num_training_steps - num_warmup_steps))
    return max(eta_min, 0.5 * (1.0 + math.cos(math.pi * progress)))


def get_cosine_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, eta_min=0):
    """``LambdaLR`` over ``cosine_warmup_factor`` (ref :159-167)."""
    return torch.optim.lr_scheduler.LambdaLR(
        optimizer, lambda step: cosine_warmup_factor(step, num_warmup_steps, num_training_steps, eta_min))


def dynamic_rounding_weight_schedule(epoch, total_epochs, initial_weight=1.0, final_weight=0.1):
    """Weight of the rounding (cross-entropy) loss, interpolated linearly over the epochs."""
    progress = epoch / total_epochs
    return initial_weight * (1 - progress) + final_weight * progress
T = 1_000
_tables = make_schedule(T)
betas = _tables.betas
alphas = _tables.alphas
alphas_cumprod = _tables.alphas_cumprod
sqrt_alphas_cumprod = _tables.sqrt_alphas_cumprod
sqrt_one_minus_alphas_cumprod = _tables.sqrt_one_minus_alphas_cumprod


def _fresh_seed() -> int:
    return int(torch.randint(0, 2**62, (), dtype=torch.int64))


def q_sample(x0: torch.Tensor, t: torch.Tensor, noise=None):
    if noise is None:
        return ops.q_sample(x0, t, None, seed=_fresh_seed())
    return ops.q_sample(x0, t, noise)


class LearnedEmbedding(nn.Module):
    """nn.Embedding(V, dim) under the reference's key ``embeddings.weight`` (ref :46-84)."""

    def __init__(self, vocab_size, embed_dim, pretrained_embeddings=None):
        super().__init__()
        self.vocab_size = vocab_size
        self.embed_dim = embed_dim
        self.embeddings = nn.Embedding(vocab_size, embed_dim)
        with torch.no_grad():
            if pretrained_embeddings is None:
                nn.init.normal_(self.embeddings.weight, mean=0.0, std=0.02)
            elif pretrained_embeddings.size(1) == embed_dim:
                self.embeddings.weight.copy_(pretrained_embeddings)
            else:
                proj = nn.Linear(pretrained_embeddings.size(1), embed_dim, bias=False).to(pretrained_embeddings.device)
                self.embeddings.weight.copy_(proj(pretrained_embeddings))

    def forward(self, token_ids):
        """embeddings[token_ids] (ref :71-80) - one gather kernel; inference only (no autograd graph)."""
        w = self.embeddings.weight
        if not w.is_cuda:
            raise _lib.TdmError("LearnedEmbedding runs on CUDA only (no CPU fallback): call .to('cuda')")
        ids = token_ids.to(device=w.device, dtype=torch.int64).contiguous()
        out = torch.empty(*ids.shape, self.embed_dim, dtype=torch.float32, device=w.device)
        if ids.numel() == 0:
            return out
        bad = torch.zeros(1, dtype=torch.int32, device=w.device)
        lib = _lib.load()
        _lib.check(lib.tdm_embedding_gather(w.detach().float().contiguous().data_ptr(), self.vocab_size, self.embed_dim,
                                            ids.data_ptr(), ids.numel(), out.data_ptr(), bad.data_ptr(),
                                            _lib.stream_ptr(w.device)), "tdm_embedding_gather")
        if int(bad):   # nn.Embedding raises IndexError on an out-of-range id; so do we (this read synchronises)
            raise IndexError("index out of range in LearnedEmbedding")
        return out

    def get_embedding_matrix(self):
        return self.embeddings.weight


class LearnedRounding(nn.Module):
    """nn.Linear(dim, V) under the reference's keys ``decoder.weight/bias`` (ref :87-102).
    Calling it materialises logits only for small problems; samplers use ``Rounder`` instead."""

    def __init__(self, embed_dim, vocab_size):
        super().__init__()
        self.decoder = nn.Linear(embed_dim, vocab_size)

    def forward(self, embeddings):
        """logits = decoder(embeddings), fp32 (..., V) (ref :93-102): the tcgen05 GEMM with a bias epilogue that
        writes the logits (bf16 products, fp32 accumulation).  The samplers do not call this - they fuse the same
        GEMM with the argmax (``Rounder.argmax``) so the (rows, V) logits never reach HBM.  Inference only."""
        w = self.decoder.weight
        if not w.is_cuda:
            raise _lib.TdmError("LearnedRounding runs on CUDA only (no CPU fallback): call .to('cuda')")
        return _rounder(w.device).logits(embeddings.to(w.device), weight=w, bias=self.decoder.bias)


class TinyTransformer(nn.Module):
    """Same parameters/keys as the reference's TinyTransformer; forward runs the CUDA engine."""

    def __init__(self, dim, n_heads=4, depth=3, dropout=0.1):
        super().__init__()
        if n_heads != 4:
            raise _lib.TdmError("the kernels implement the reference's 4-head configuration")
        layer = nn.TransformerEncoderLayer(d_model=dim, nhead=n_heads, batch_first=True, dropout=dropout)
        self.encoder = nn.TransformerEncoder(layer, num_layers=depth, enable_nested_tensor=False)
        self.time_emb = nn.Linear(1, dim)
        self.dropout = nn.Dropout(dropout)
        self.dim = dim
        self._engines: dict = {}

    def _param_version(self) -> int:
        return sum(p._version for p in self.parameters())

    def engine(self, batch: int, seq_len: int) -> TextEngine:
        p0 = next(self.parameters())
        if not p0.is_cuda:
            raise _lib.TdmError("TinyTransformer runs on CUDA only (no CPU fallback): call .to('cuda')")
        key = (batch, seq_len, p0.device, self._param_version(), p0.data_ptr())
        if key not in self._engines:
            self._engines.clear()
            self._engines[key] = TextEngine(self.state_dict(), p0.device, batch, seq_len)
        return self._engines[key]

    def forward(self, x: torch.Tensor, t: torch.Tensor):
        if self.training or (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
                             and x.requires_grad):
            raise _lib.TdmError("TinyTransformer.forward is inference-only (eval + no_grad): training runs through "
                                "shakespeare.train / text_train.TextTrainer, whose backward pass is CUDA, not autograd")
        return self.engine(x.shape[0], x.shape[1]).forward(x, t)


def load_text_dataset():
    """The raw Shakespeare corpus as one string (ref :122-125; needs the HF hub)."""
    from datasets import load_dataset

    ds = load_dataset("tiny_shakespeare", trust_remote_code=True)
    return "\n\n".join(ds["train"]["text"] + ds["test"]["text"] + ds["validation"]["text"])


def tokenize_corpus(text: str, tokenizer, seq_len: int, val_split=0.1):
    """Tokenise the corpus once, cut it into (n_chunks, seq_len) rows, random train/val split (ref :128-157)."""
    from torch.utils.data import random_split

    ids = tokenizer(text, add_special_tokens=False, return_attention_mask=False, return_tensors="pt").input_ids.squeeze(0)
    n_chunks = ids.size(0) // seq_len
    chunks = ids[: n_chunks * seq_len].view(n_chunks, seq_len)
    n_val = int(n_chunks * val_split)
    return random_split(chunks, [n_chunks - n_val, n_val])


def _mean_losses(sums, n):
    return {k: v / max(n, 1) for k, v in zip(("diff", "round", "total"), sums)}


def train(model, rounding_fn, embedding_fn, data_loader, val_loader, device, ckpt_path="text_ckpt.pth", epochs=1,
          lr=1e-4, weight_decay=1e-4, rounding_weight=1.0, use_learned_embeddings=True, patience=5,
          use_lr_scheduling=True, warmup_steps=100, *, seed=None, log_every=50):
    """The reference's training loop (ref :174-341) with the step on the device: per batch one replayed CUDA graph
    (forward, both losses, backward, AdamW, weight re-pack; ``text_train.TextTrainer``), the cosine / warm-up learning
    rate and the per-epoch rounding weight written to device scalars, validation in eval mode, best / final
    checkpoints in the reference's format.  Losses are accumulated on the device and read once per ``log_every`` steps
    (the reference's per-step ``.item()`` calls would serialise the stream)."""
    from .text_train import TextTrainer
    from .utils import get_vertex_checkpoint_path, save_checkpoint

    first = next(iter(data_loader))
    batch, seq_len = int(first.shape[0]), int(first.shape[1])
    tr = TextTrainer(model, rounding_fn, embedding_fn, device, batch, seq_len, lr=lr, weight_decay=weight_decay,
                     use_learned_embeddings=use_learned_embeddings, seed=_fresh_seed() if seed is None else seed)
    total_steps = len(data_loader) * epochs

    def lr_at(k):   # what LambdaLR leaves in param_groups[0]["lr"] for optimiser step k (0-based; ref :199-200, :250)
        return lr * cosine_warmup_factor(k, warmup_steps, total_steps) if use_lr_scheduling else lr

    best_val, bad_epochs, k = float("inf"), 0, 0
    for epoch in range(epochs):
        rw = dynamic_rounding_weight_schedule(epoch, epochs, rounding_weight)
        acc = torch.zeros(3, device=tr.device)
        n_tr = 0
        for token_ids in data_loader:
            if tuple(token_ids.shape) != (batch, seq_len):
                continue   # a ragged last batch would need a second captured graph; it is skipped (stated in DESIGN.md)
            acc += tr.step(token_ids, lr=lr_at(k), rounding_weight=rw)
            k += 1
            n_tr += 1
            if log_every and n_tr % log_every == 0:
                d, r, t_ = (acc / n_tr).tolist()
                print(f"Epoch {epoch + 1}/{epochs} step {n_tr}: diff={d:.4f} round={r:.4f} total={t_:.4f} rw={rw:.3f} lr={lr_at(k):.2e}")
        tr.check_token_ids()
        train_losses = _mean_losses((acc / max(n_tr, 1)).tolist(), 1)
        model.eval(); rounding_fn.eval()
        vacc = torch.zeros(3, device=tr.device)
        n_val = 0
        for token_ids in val_loader:
            if tuple(token_ids.shape) != (batch, seq_len):
                continue
            vacc += tr.evaluate(token_ids)
            n_val += 1
        tr.check_token_ids()
        val_losses = _mean_losses((vacc / max(n_val, 1)).tolist(), 1)
        print(f"Epoch {epoch + 1}/{epochs}:")
        print(f"  Train: diff={train_losses['diff']:.4f}, round={train_losses['round']:.4f}, total={train_losses['total']:.4f}")
        print(f"  Val:   diff={val_losses['diff']:.4f}, round={val_losses['round']:.4f}, total={val_losses['total']:.4f}")
        print(f"  Rounding weight: {rw:.3f}")
        if n_val == 0 or val_losses["total"] < best_val:
            best_val = val_losses["total"] if n_val else best_val
            bad_epochs = 0
            best_path = ckpt_path.replace(".pth", "_best.pth")
            ck = {"diffusion_model": model.state_dict(), "rounding_fn": rounding_fn.state_dict(), "epoch": epoch,
                  "val_loss": best_val}
            if use_learned_embeddings:
                ck["embedding_fn"] = embedding_fn.state_dict()
            save_checkpoint(ck, best_path)
            print(f"  New best validation loss! Saved to {best_path}")
        else:
            bad_epochs += 1
            if bad_epochs >= patience:
                print(f"  Early stopping triggered after {patience} epochs without improvement")
                break
    tr.sync_modules()
    final_path = get_vertex_checkpoint_path("text-model.pth") if "AIP_MODEL_DIR" in os.environ else ckpt_path
    print(f"✔ Saving final checkpoint to {final_path}...")
    ck = {"diffusion_model": model.state_dict(), "rounding_fn": rounding_fn.state_dict(), "epoch": epochs,
          "final_training": True}
    if use_learned_embeddings:
        ck["embedding_fn"] = embedding_fn.state_dict()
    save_checkpoint(ck, final_path)
    return tr


def p_sample(model, x, t):
    """One reverse step with in-kernel Philox noise."""
    return model.engine(x.shape[0], x.shape[1]).p_sample(x, t, None, seed=_fresh_seed())


_rounders: dict = {}


def _rounder(device) -> Rounder:
    key = str(torch.device(device))
    if key not in _rounders:
        _rounders[key] = Rounder(device)
    return _rounders[key]


def round_to_tokens(x, rounding_fn, embedding_fn, use_learned_rounding=True, use_learned_embeddings=True):
    """tokens = argmax over V (ref :387-401)."""
    r = _rounder(x.device)
    if use_learned_rounding:
        return r.argmax(x, weight=rounding_fn.decoder.weight, bias=rounding_fn.decoder.bias)
    emb = embedding_fn.get_embedding_matrix() if use_learned_embeddings else embedding_fn
    return r.argmax(x, weight=emb, cosine=True)


def sample_diffusion_embeddings(model, embed_dim, device, n, seq_len, seed=None, sample_offset=0):
    """Pure diffusion embeddings z (n, seq_len, embed_dim) — the T-step loop of ref :418-426."""
    seed = _fresh_seed() if seed is None else seed
    model.eval()
    x = ops.randn((n, seq_len, embed_dim), device, seed=seed, sample_offset=sample_offset, stream_id=0)
    return model.engine(n, seq_len).sample_loop(x, seed=seed, sample_offset=sample_offset, steps=T)


def sample(model, rounding_fn, embedding_fn, tokenizer, device, n_samples=4, seq_len=128,
           use_learned_rounding=True, use_learned_embeddings=True, embed_dim=None):
    model.eval()
    samples_dir = get_samples_dir("samples")
    with torch.no_grad():
        if embed_dim is None:
            embed_dim = embedding_fn.embed_dim if use_learned_embeddings else embedding_fn.shape[1]
        x = sample_diffusion_embeddings(model, embed_dim, device, n_samples, seq_len)
        tokens = round_to_tokens(x, rounding_fn, embedding_fn, use_learned_rounding, use_learned_embeddings)
        texts = tokenizer.batch_decode(tokens, skip_special_tokens=True)
        for i, text in enumerate(texts):
            print(text)
            if isinstance(samples_dir, str) and samples_dir.startswith("gs://"):
                path = f"{samples_dir}/sample_{i}.txt"
            else:
                path = Path(samples_dir) / f"sample_{i}.txt"
            save_samples(text, path)
            print(f"✔ Wrote {path}")
        return texts


class _LMStepper:
    """AR logits of the next position (ref :447-449).

    The reference re-runs the base LM over the whole prefix at every position - O(L^2) per sequence.  An LM that
    implements the Hugging Face cache protocol (``use_cache=True`` / ``past_key_values``) is fed only the newest token
    against its cached keys and values, O(L) per sequence, with the same logits up to fp rounding; any other callable
    gets exactly the reference's call.  The LM itself stays a third-party torch module.
    """

    def __init__(self, base_lm, use_kv_cache: bool = True):
        self.lm = base_lm
        self.mode = None if use_kv_cache else "prefix"
        self.cache = None

    def __call__(self, input_ids: torch.Tensor, pos: int) -> torch.Tensor:
        if self.mode is None:   # probe once, on the first position
            try:
                out = self.lm(input_ids[:, :pos + 1], use_cache=True)
                self.cache = getattr(out, "past_key_values", None)
                self.mode = "cache" if self.cache is not None else "prefix"
                return out.logits[:, -1, :]
            except TypeError:   # a plain callable(input_ids)
                self.mode = "prefix"
        if self.mode == "cache":
            out = self.lm(input_ids[:, pos:pos + 1], past_key_values=self.cache, use_cache=True)
            self.cache = out.past_key_values
            return out.logits[:, -1, :]
        return self.lm(input_ids[:, :pos + 1]).logits[:, -1, :]


def guided_generate(base_lm, rounding_fn, tokenizer, embedding_fn, diff_z, alpha=0.5, max_len=128,
                    temperature=1.0, use_learned_rounding=True, use_learned_embeddings=True, *, use_kv_cache=True):
    """Greedy AR decoding steered by the diffusion embeddings (ref :429-470).  Per position the
    diffusion logits, the (1-alpha)/alpha mix with the AR logits and the argmax are one kernel; the base LM is
    stepped incrementally through its KV cache when it has one (``_LMStepper``; ``use_kv_cache=False`` forces the
    reference's full-prefix re-forward)."""
    device = diff_z.device
    B, L, _ = diff_z.shape
    r = _rounder(device)
    start = tokenizer.bos_token_id or tokenizer.eos_token_id
    input_ids = torch.full((B, L + 1), start, device=device, dtype=torch.long)   # preallocated, no per-step cat
    if use_learned_rounding:
        kw = dict(weight=rounding_fn.decoder.weight, bias=rounding_fn.decoder.bias, cosine=False)
    else:
        emb = embedding_fn.get_embedding_matrix() if use_learned_embeddings else embedding_fn
        kw = dict(weight=emb, cosine=True)
    lm_step = _LMStepper(base_lm, use_kv_cache)
    with torch.no_grad():
        for pos in range(L):
            ar_logits = lm_step(input_ids, pos)                                   # third-party LM forward
            nxt = r.argmax(diff_z[:, pos, :], ar_logits=ar_logits, alpha=alpha, temperature=temperature, **kw)
            input_ids[:, pos + 1] = nxt
    return tokenizer.batch_decode(input_ids[:, 1:], skip_special_tokens=True)


# ---------------------------------------------------------------------------------------------
# CLI — the reference's 22 flags (src/shakespeare.py:475-496) plus an offline path
# ---------------------------------------------------------------------------------------------
class _SyntheticTokenizer:
    """Offline stand-in (no HF hub on the benchmark boxes): decodes ids as space-separated ints."""
    bos_token_id = 2
    eos_token_id = 1

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(str(int(i)) for i in row) for row in ids]


class _SyntheticLM(nn.Module):
    """Random-init bigram LM standing in for google/gemma-2b-it when --synthetic is given; like Gemma-2b its
    embedding width is 2048, so the default flags build the same TinyTransformer(2048) the reference CLI does
    (``--use_learned_embeddings --embed_dim 256`` selects the deployed width)."""

    def __init__(self, vocab, dim=2048):
        super().__init__()
        self.emb = nn.Embedding(vocab, dim)
        self.out = nn.Linear(dim, vocab)

    def get_input_embeddings(self):
        return self.emb

    def forward(self, input_ids, past_key_values=None, use_cache=False):
        import types
        if use_cache:   # a bigram LM's whole "cache" is the newest token: only its logits are needed (and computed)
            return types.SimpleNamespace(logits=self.out(self.emb(input_ids[:, -1:])), past_key_values=("bigram",))
        return types.SimpleNamespace(logits=self.out(self.emb(input_ids)))


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--train", action="store_true")
    parser.add_argument("--sample", action="store_true", help="plain diffusion sample")
    parser.add_argument("--guided_sample", action="store_true", help="AR + diffusion guidance")
    parser.add_argument("--epochs", type=int, default=1)
    parser.add_argument("--batch_size", type=int, default=32)
    parser.add_argument("--seq_len", type=int, default=64)
    parser.add_argument("--ckpt", type=str, default="gs://text-diffusion/diffusion/outputs/model/text-model.pth"
                        if "AIP_MODEL_DIR" in os.environ else "text_ckpt.pth")
    parser.add_argument("--model_id", type=str, default="google/gemma-2b-it")
    parser.add_argument("--n", type=int, default=10)
    parser.add_argument("--alpha", type=float, default=0.3)
    parser.add_argument("--rounding_weight", type=float, default=1.0)
    parser.add_argument("--use_cosine_fallback", action="store_true")
    parser.add_argument("--use_learned_embeddings", action="store_true")
    parser.add_argument("--embed_dim", type=int, default=None)
    parser.add_argument("--init_from_pretrained", action="store_true")
    parser.add_argument("--dropout", type=float, default=0.1)
    parser.add_argument("--weight_decay", type=float, default=1e-4)
    parser.add_argument("--patience", type=int, default=5)
    parser.add_argument("--use_lr_scheduling", action="store_true", default=True)
    parser.add_argument("--warmup_steps", type=int, default=100)
    parser.add_argument("--val_split", type=float, default=0.1)
    parser.add_argument("--lr", type=float, default=1e-4)
    # additive: run without the HF hub (random-init stand-ins, synthetic vocabulary)
    parser.add_argument("--synthetic", action="store_true", help="no HF downloads: synthetic tokenizer/LM")
    parser.add_argument("--vocab_size", type=int, default=8192, help="vocabulary for --synthetic")
    parser.add_argument("--seed", type=int, default=None)
    args = parser.parse_args(argv)

    if not torch.cuda.is_available():
        raise _lib.TdmError("tinydiffusionmodels_b200 needs a CUDA (sm_100a) device; there is no CPU path")
    device = "cuda"
    print(f"Device: {device}")
    if args.seed is not None:
        torch.manual_seed(args.seed)

    if args.synthetic:
        tokenizer = _SyntheticTokenizer()
        lm_model = _SyntheticLM(args.vocab_size).to(device).eval()
    else:
        from transformers import AutoModelForCausalLM, AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(args.model_id)
        lm_model = AutoModelForCausalLM.from_pretrained(args.model_id).to(device).eval()
    pretrained = lm_model.get_input_embeddings().weight.detach().to(device)
    vocab_size, pretrained_dim = pretrained.shape

    if args.use_learned_embeddings:
        embed_dim = args.embed_dim if args.embed_dim is not None else pretrained_dim
        embedding_fn = LearnedEmbedding(vocab_size, embed_dim, pretrained if args.init_from_pretrained else None).to(device)
    else:
        embed_dim = pretrained_dim
        embedding_fn = pretrained
    diff_model = TinyTransformer(embed_dim, dropout=args.dropout).to(device)
    rounding_fn = LearnedRounding(embed_dim, vocab_size).to(device)

    if args.train:
        from torch.utils.data import DataLoader
        if args.synthetic:   # no HF hub: a synthetic corpus of random ids, enough to exercise the whole loop
            gen = torch.Generator().manual_seed(0 if args.seed is None else args.seed)
            chunks = torch.randint(0, vocab_size, (args.batch_size * 12, args.seq_len), generator=gen)
            n_val = max(args.batch_size, int(chunks.shape[0] * args.val_split) // args.batch_size * args.batch_size)
            train_chunks, val_chunks = chunks[n_val:], chunks[:n_val]
        else:
            train_chunks, val_chunks = tokenize_corpus(load_text_dataset(), tokenizer, args.seq_len, args.val_split)
        train_dl = DataLoader(train_chunks, batch_size=args.batch_size, shuffle=True)
        val_dl = DataLoader(val_chunks, batch_size=args.batch_size, shuffle=False)
        print(f"Training on {len(train_chunks)} chunks, validating on {len(val_chunks)} chunks")
        train(diff_model, rounding_fn, embedding_fn, train_dl, val_dl, device, args.ckpt, epochs=args.epochs, lr=args.lr,
              weight_decay=args.weight_decay, rounding_weight=args.rounding_weight,
              use_learned_embeddings=args.use_learned_embeddings, patience=args.patience,
              use_lr_scheduling=args.use_lr_scheduling, warmup_steps=args.warmup_steps, seed=args.seed)

    def load():
        nonlocal embedding_fn
        if not (args.synthetic and not os.path.exists(args.ckpt)):
            ck = load_checkpoint(args.ckpt, device)
            if isinstance(ck, dict) and "diffusion_model" in ck:
                diff_model.load_state_dict(ck["diffusion_model"])
                rounding_fn.load_state_dict(ck["rounding_fn"])
                if args.use_learned_embeddings and "embedding_fn" in ck:
                    embedding_fn.load_state_dict(ck["embedding_fn"])
                elif args.use_learned_embeddings:
                    print("Warning: Learned embeddings requested but not found in checkpoint. Using pre-trained fallback.")
                    args.use_learned_embeddings = False
                    embedding_fn = pretrained
            else:
                diff_model.load_state_dict(ck)
                print("Warning: Using old checkpoint format. Falling back to pre-trained embeddings and cosine similarity.")
                args.use_cosine_fallback = True
                args.use_learned_embeddings = False
                embedding_fn = pretrained
        else:
            print("--synthetic without a checkpoint: sampling from random-init weights")

    if args.sample:
        load()
        sample(diff_model, rounding_fn, embedding_fn, tokenizer, device, args.n, args.seq_len,
               use_learned_rounding=not args.use_cosine_fallback,
               use_learned_embeddings=args.use_learned_embeddings, embed_dim=embed_dim)
    if args.guided_sample:
        load()
        z = sample_diffusion_embeddings(diff_model, embed_dim, device, args.n, args.seq_len)
        texts = guided_generate(lm_model, rounding_fn, tokenizer, embedding_fn, z, alpha=args.alpha,
                                max_len=args.seq_len, use_learned_rounding=not args.use_cosine_fallback,
                                use_learned_embeddings=args.use_learned_embeddings)
        samples_dir = get_samples_dir("samples")
        for i, text in enumerate(texts):
            if isinstance(samples_dir, str) and samples_dir.startswith("gs://"):
                path = f"{samples_dir}/guided_sample_{i}.txt"
            else:
                path = Path(samples_dir) / f"guided_sample_{i}.txt"
            save_samples(text, path)
            print(f"✔ Wrote {path}")
    if not (args.train or args.sample or args.guided_sample):
        print("Nothing to do. Try --train or --guided_sample.")


if __name__ == "__main__":
    main()
