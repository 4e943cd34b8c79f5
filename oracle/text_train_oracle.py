"""CPU restatement of one Shakespeare training step — TEST INFRASTRUCTURE, never imported by the product.

Follows /root/reference/src/shakespeare.py:221-250 (``train``'s inner loop) op by op in torch fp32 with autograd:

    x0 = embedding_fn(token_ids)                      :226-228
    t ~ randint, noise ~ randn, x_noisy = q_sample    :230-232
    noise_pred = model(x_noisy, t)   (TRAIN mode)     :233 -> TinyTransformer.forward :115-120
    diffusion_loss = mse(noise_pred, noise)           :236
    rounding_loss = cross_entropy(rounding_fn(x0))    :239-241
    total = diffusion + w * rounding                  :244
    backward, AdamW step                              :246-248, optimiser built at :196

The reference draws t, noise and its dropout masks from torch's global generator, whose stream cannot be matched from
another implementation, so — as for the samplers — parity is defined with those three *injected*: t and noise as
tensors, the masks as the library's own counter-based bits (csrc/text_train.cu: Philox4x32-10 keyed on (seed, optimiser
step, dropout site, element)), regenerated here in numpy through oracle/philox.py.

Where the dropout sites are: ``TinyTransformer.forward`` applies ``self.dropout`` to x + time bias (:118-119);
nn.TransformerEncoderLayer (post-norm, PyTorch ``torch/nn/modules/transformer.py``: ``x = norm1(x + _sa_block(x))``,
``x = norm2(x + _ff_block(x))``) drops the attention weights inside scaled-dot-product attention, the attention block's
output (``dropout1``), the activations inside the feed-forward block (``dropout``) and its output (``dropout2``).

Pinning: with every mask all-ones this restatement's loss and gradients equal the unmodified reference module's in
train mode with dropout=0 (tests/golden/text_train_golden.pt, recorded from the imported reference by
tests/golden/make_golden_text_train.py, and a live check when /root/reference is present); the AdamW update is pinned
against torch.optim.AdamW in tests/test_oracle.py.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from . import philox
from .ddpm_oracle import T, q_sample

DOMAIN_DROPOUT, DOMAIN_TIMESTEP = 3, 4
SITE_INPUT = 0xFFFF0000


def keep_mask(shape, p: float, seed: int, step: int, site: int) -> torch.Tensor:
    """Dropout multiplier (0 or 1/(1-p)) of every element of a contiguous tensor of this shape: element i keeps iff
    word (i % 4) of Philox(counter = (i // 4, site, step, DOMAIN_DROPOUT), key = seed) >= floor(p * 2^32)."""
    n = int(np.prod(shape))
    if p <= 0.0:
        return torch.ones(shape)
    assert n % 4 == 0
    thresh = np.uint32(min(int(float(np.float32(p)) * 4294967296.0), 4294967295))
    quad = np.arange(n // 4, dtype=np.uint32)
    words = philox.philox4x32_10(quad, np.full_like(quad, site), np.full_like(quad, step & 0xFFFFFFFF),
                                 np.full_like(quad, DOMAIN_DROPOUT), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    bits = np.stack(words, axis=-1).reshape(-1)
    keep = (bits >= thresh).astype(np.float32) * (np.float32(1.0) / (np.float32(1.0) - np.float32(p)))   # fp32, as the kernels
    return torch.from_numpy(keep).reshape(shape)


def draw_t(batch: int, seed: int, step: int, sample_offset: int = 0, n_steps: int = T) -> torch.Tensor:
    """t[b] = floor(word0 * T / 2^32) of Philox(counter = (sample lo, sample hi, step, DOMAIN_TIMESTEP))."""
    s = np.arange(batch, dtype=np.uint64) + np.uint64(sample_offset)
    lo = (s & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi = (s >> np.uint64(32)).astype(np.uint32)
    w0 = philox.philox4x32_10(lo, hi, np.full_like(lo, step & 0xFFFFFFFF), np.full_like(lo, DOMAIN_TIMESTEP),
                              seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)[0]
    return torch.from_numpy(((w0.astype(np.uint64) * np.uint64(n_steps)) >> np.uint64(32)).astype(np.int64))


def train_noise(batch: int, seq_len: int, dim: int, seed: int, step: int, sample_offset: int = 0) -> torch.Tensor:
    """The in-kernel training noise: the q_sample domain of the library's generator, stream id = optimiser step."""
    z = philox.randn(batch, seq_len * dim, seed, sample_offset, step & 0xFFFFFFFF, philox.DOMAIN_QSAMPLE)
    return torch.from_numpy(z).reshape(batch, seq_len, dim)


def transformer_forward_train(sd: dict, x: torch.Tensor, t: torch.Tensor, p: float, seed: int, step: int,
                              n_heads: int = 4, eps: float = 1e-5, collect: dict | None = None) -> torch.Tensor:
    """TinyTransformer.forward in train mode (src/shakespeare.py:115-120) with the masks of keep_mask."""
    B, L, D = x.shape
    hd = D // n_heads
    ts = (t.float() / T).unsqueeze(-1)                                                     # :116
    h = x + F.linear(ts, sd["time_emb.weight"], sd["time_emb.bias"]).unsqueeze(1)          # :117-118
    h = h * keep_mask((B, L, D), p, seed, step, SITE_INPUT)                                # :119
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.layers."))
    if collect is not None:
        collect["h0"] = h
    for i in range(depth):
        pre = f"encoder.layers.{i}."
        site = 16 * i
        qkv = F.linear(h, sd[pre + "self_attn.in_proj_weight"], sd[pre + "self_attn.in_proj_bias"])
        q, k, v = qkv.split(D, dim=-1)
        q = q.view(B, L, n_heads, hd).transpose(1, 2)
        k = k.view(B, L, n_heads, hd).transpose(1, 2)
        v = v.view(B, L, n_heads, hd).transpose(1, 2)
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        att = att * keep_mask((B, n_heads, L, L), p, seed, step, site + 1)                 # SDPA dropout_p
        o = (att @ v).transpose(1, 2).reshape(B, L, D)
        o = F.linear(o, sd[pre + "self_attn.out_proj.weight"], sd[pre + "self_attn.out_proj.bias"])
        h = F.layer_norm(h + o * keep_mask((B, L, D), p, seed, step, site + 2), (D,), sd[pre + "norm1.weight"],
                         sd[pre + "norm1.bias"], eps)                                      # dropout1, norm1
        f = F.relu(F.linear(h, sd[pre + "linear1.weight"], sd[pre + "linear1.bias"]))
        f = f * keep_mask(tuple(f.shape), p, seed, step, site + 3)                         # dropout
        f = F.linear(f, sd[pre + "linear2.weight"], sd[pre + "linear2.bias"])
        h = F.layer_norm(h + f * keep_mask((B, L, D), p, seed, step, site + 4), (D,), sd[pre + "norm2.weight"],
                         sd[pre + "norm2.bias"], eps)                                      # dropout2, norm2
        if collect is not None:
            collect[f"h{i + 1}"] = h
    return h


def text_losses_and_grads(model_sd: dict, dec_w: torch.Tensor, dec_b: torch.Tensor, emb: torch.Tensor,
                          token_ids: torch.Tensor, t: torch.Tensor, noise: torch.Tensor, tab: dict, *,
                          rounding_weight: float = 1.0, dropout: float = 0.0, seed: int = 0, step: int = 1,
                          learn_embeddings: bool = True, want_grads: bool = True):
    """Returns ((diffusion, rounding, total), grads) with grads keyed 'model.<name>', 'decoder.weight',
    'decoder.bias', 'embeddings.weight' (src/shakespeare.py:226-246)."""
    params = {"model." + k: v.detach().clone().requires_grad_(want_grads) for k, v in model_sd.items()}
    params["decoder.weight"] = dec_w.detach().clone().requires_grad_(want_grads)
    params["decoder.bias"] = dec_b.detach().clone().requires_grad_(want_grads)
    params["embeddings.weight"] = emb.detach().clone().requires_grad_(want_grads and learn_embeddings)
    sd = {k[len("model."):]: v for k, v in params.items() if k.startswith("model.")}
    x0 = params["embeddings.weight"][token_ids]                                            # :226-228
    x_noisy = q_sample(x0, t, noise, tab)                                                  # :232
    pred = transformer_forward_train(sd, x_noisy, t, dropout, seed, step)                  # :233
    diff = F.mse_loss(pred, noise)                                                         # :236
    logits = F.linear(x0, params["decoder.weight"], params["decoder.bias"])                # :239
    rnd = F.cross_entropy(logits.reshape(-1, logits.size(-1)), token_ids.reshape(-1))      # :241
    total = diff + rounding_weight * rnd                                                   # :244
    grads = {}
    if want_grads:
        keys = [k for k, v in params.items() if v.requires_grad]
        gs = torch.autograd.grad(total, [params[k] for k in keys])
        grads = dict(zip(keys, gs))
    return (diff.detach(), rnd.detach(), total.detach()), grads
