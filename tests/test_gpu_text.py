"""GPU parity of the Shakespeare sampler kernels against the CPU fp32 oracle.

Tolerances (stated): bf16 tensor-core GEMMs with fp32 accumulation, fp32 residual stream and
LayerNorm.  The post-LN denoiser output has rms 1, and one forward differs from fp32 by <= 1% rms /
6% of rms worst element; rounded tokens must be bit-exact wherever the fp32 top-2 margin exceeds the
stated score tolerance (north_star), and are reported otherwise.
"""
import types

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import philox as PX
from tests.helpers import rel_rms
from tinydiffusionmodels_b200.shakespeare import (LearnedEmbedding, LearnedRounding, TinyTransformer,
                                                  guided_generate, round_to_tokens)
from tinydiffusionmodels_b200.text_engine import Rounder

pytestmark = pytest.mark.gpu
TAB = O.make_tables()


def _model(dim, seed=0):
    torch.manual_seed(seed)
    m = TinyTransformer(dim).eval()
    # random LayerNorm affine so gamma/beta handling is actually exercised
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
    return m


@pytest.mark.parametrize("dim,batch,seq", [(256, 2, 64), (256, 5, 64), (256, 3, 128), (2048, 2, 64)])
def test_transformer_forward_matches_oracle(cuda, dim, batch, seq):
    m = _model(dim)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(dim + batch)
    x = torch.randn(batch, seq, dim, generator=g) * 3
    t = torch.randint(0, 1000, (batch,), generator=g)
    ref = O.transformer_forward(sd, x, t)
    with torch.no_grad():
        got = m.to(cuda)(x.to(cuda), t.to(cuda)).cpu()
    e_rms, e_max = rel_rms(got, ref), float((got - ref).abs().max() / ref.pow(2).mean().sqrt())
    print(f"dim {dim} B {batch} L {seq}: rel-rms {e_rms:.3e} max/rms {e_max:.3e}")
    assert e_rms < 1e-2 and e_max < 6e-2


@pytest.mark.parametrize("tval", [700, 0])
def test_text_p_sample_injected_noise(cuda, tval):
    m = _model(256, 1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 64, 256, generator=g) * 5
    z = torch.randn(4, 64, 256, generator=g)
    t = torch.full((4,), tval, dtype=torch.long)
    ref = O.text_p_sample(sd, x, t, z, TAB)
    eng = m.to(cuda).engine(4, 64)
    got = eng.p_sample(x.to(cuda), t.to(cuda), z.to(cuda)).cpu()
    # eps error (<= 6e-2 abs) is scaled by beta_t/sqrt(1-acp_t) <= 0.02
    torch.testing.assert_close(got, ref, rtol=0, atol=2e-3)


def test_text_sample_loop_philox_matches_oracle(cuda):
    """20 reverse steps (t = 19..0) with in-kernel Philox noise vs the oracle fed the numpy-Philox
    noise; also shard invariance: two half batches reproduce the full batch."""
    m = _model(256, 2)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    B, L, D, steps, seed = 4, 64, 256, 20, 1234
    x0 = torch.from_numpy(PX.randn(B, L * D, seed, 0, 0, PX.DOMAIN_INIT)).view(B, L, D)
    x = x0.clone()
    for i in reversed(range(steps)):
        t = torch.full((B,), i, dtype=torch.long)
        z = torch.from_numpy(PX.randn(B, L * D, seed, 0, i, PX.DOMAIN_REVERSE)).view(B, L, D)
        x = O.text_p_sample(sd, x, t, None if i == 0 else z, TAB)
    mg = m.to(cuda)
    from tinydiffusionmodels_b200 import ops
    xin = ops.randn((B, L, D), cuda, seed=seed, sample_offset=0, stream_id=0)
    torch.testing.assert_close(xin.cpu(), x0, rtol=0, atol=2e-5)
    got = mg.engine(B, L).sample_loop(xin.clone(), seed=seed, steps=steps).cpu()
    print("loop rel-rms", rel_rms(got, x))
    assert rel_rms(got, x) < 5e-3
    # eager (non-graph) loop is bit-identical to the graph-replayed loop
    eager = mg.engine(B, L).sample_loop(xin.clone(), seed=seed, steps=steps, use_graph=False).cpu()
    assert torch.equal(eager, got)
    # shards
    lo = mg.engine(2, L).sample_loop(xin[:2].clone(), seed=seed, sample_offset=0, steps=steps).cpu()
    hi = mg.engine(2, L).sample_loop(xin[2:].clone(), seed=seed, sample_offset=2, steps=steps).cpu()
    assert torch.equal(torch.cat([lo, hi]), got)


def _margin_check(got, logits, tol, what):
    """bit-exact wherever the fp32 top-2 margin exceeds tol; report the rest."""
    top2 = logits.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1])
    ref = logits.argmax(-1)
    safe = margin > tol
    bad = (got != ref) & safe
    frac_unsafe = float((~safe).float().mean())
    mism = int((got != ref).sum())
    print(f"{what}: {mism} mismatches of {ref.numel()}, {frac_unsafe:.2%} positions below margin {tol}")
    assert not bad.any(), f"{what}: token differs at a position whose fp32 margin exceeds {tol}"


@pytest.mark.parametrize("vocab", [8192, 5000, 257])
def test_learned_and_cosine_rounding(cuda, vocab):
    torch.manual_seed(3)
    dim = 256
    rounding = LearnedRounding(dim, vocab)
    emb = LearnedEmbedding(vocab, dim)
    x = torch.randn(5, 64, dim) * 4
    w, b = rounding.decoder.weight.detach(), rounding.decoder.bias.detach()
    e = emb.embeddings.weight.detach()
    logits = O.learned_logits(x, w, b)
    sims = O.cosine_logits(x, e)
    got = round_to_tokens(x.to(cuda), rounding.to(cuda), emb.to(cuda), True, True).cpu()
    # bf16 inputs: |x.w| error <= 2^-8 * sum|x_i w_i| ~ 0.05 here
    _margin_check(got, logits, 0.08, f"learned V={vocab}")
    got_c = round_to_tokens(x.to(cuda), rounding.to(cuda), emb.to(cuda), False, True).cpu()
    _margin_check(got_c, sims, 4e-3, f"cosine V={vocab}")
    assert got.max() < vocab and got_c.max() < vocab


def test_rounding_values_and_ties(cuda):
    # exact small-integer data: bf16 is exact, so scores and tie-breaking must match torch exactly
    torch.manual_seed(0)
    dim, vocab, rows = 128, 300, 7
    x = torch.randint(-3, 4, (rows, dim)).float()
    w = torch.randint(-2, 3, (vocab, dim)).float()
    w[17] = w[5]            # duplicate rows -> exact ties; the lower index must win
    w[250] = w[5]
    b = torch.zeros(vocab)
    r = Rounder(cuda)
    idx, val = r.argmax(x.to(cuda), weight=w.to(cuda), bias=b.to(cuda), return_values=True)
    logits = x @ w.T
    assert torch.equal(idx.cpu(), logits.argmax(-1))
    assert torch.equal(val.cpu(), logits.max(-1).values)


class _LM(torch.nn.Module):
    def __init__(self, vocab, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.table = torch.nn.Parameter(torch.randn(vocab, vocab, generator=g) * 2)

    def forward(self, input_ids):
        return types.SimpleNamespace(logits=self.table[input_ids])


class _Tok:
    bos_token_id = 2
    eos_token_id = 1

    def batch_decode(self, ids, skip_special_tokens=True):
        return [[int(i) for i in row] for row in ids]


@pytest.mark.parametrize("learned", [True, False])
def test_guided_generate(cuda, learned):
    torch.manual_seed(4)
    dim, vocab, B, L = 256, 1000, 6, 64
    rounding = LearnedRounding(dim, vocab)
    emb = LearnedEmbedding(vocab, dim)
    lm = _LM(vocab, 5)
    z = torch.randn(B, L, dim) * 4
    alpha, temp = 0.3, 0.7
    # oracle, position by position, with the margin of every decision
    w, b = rounding.decoder.weight.detach(), rounding.decoder.bias.detach()
    e = emb.embeddings.weight.detach()
    ids = torch.full((B, 1), 2, dtype=torch.long)
    margins = []
    for pos in range(L):
        ar = lm(ids).logits[:, -1, :].detach() / temp
        diff = (O.learned_logits(z[:, pos], w, b) if learned else O.cosine_logits(z[:, pos], e)) / temp
        mixed = (1 - alpha) * ar + alpha * diff
        top2 = mixed.topk(2, dim=-1).values
        margins.append(top2[:, 0] - top2[:, 1])
        ids = torch.cat([ids, mixed.argmax(-1, keepdim=True)], 1)
    ref = ids[:, 1:]
    margins = torch.stack(margins, 1)
    got = guided_generate(lm.to(cuda), rounding.to(cuda), _Tok(), emb.to(cuda), z.to(cuda), alpha=alpha,
                          temperature=temp, use_learned_rounding=learned, use_learned_embeddings=True)
    got = torch.tensor(got)
    # a sequence is comparable up to its first low-margin decision (after that the AR context differs)
    tol = 0.05 if learned else 2e-3
    for bi in range(B):
        low = (margins[bi] <= tol).nonzero()
        upto = int(low[0]) if len(low) else L
        assert torch.equal(got[bi, :upto], ref[bi, :upto]), f"sequence {bi} diverges before position {upto}"
    print("exact prefix lengths ok; full-sequence agreement", float((got == ref).float().mean()))


# ---- the two module forwards the samplers never call (ref src/shakespeare.py:71-80, 93-102) ------------------------
@pytest.mark.parametrize("vocab", [8192, 5000, 257])
def test_learned_rounding_forward_returns_logits(cuda, vocab):
    """LearnedRounding.forward = decoder(x): fp32 logits from the tcgen05 GEMM (bf16 products, fp32 accumulation).
    Tolerance: |x.w| error <= 2^-8 * sum|x_i w_i| (two bf16 roundings per product); asserted as max-abs <= 0.08
    (the margin the token tests use) and rel-rms <= 5e-3 against the fp32 oracle, the oracle itself pinned to the
    reference's own logits by tests/test_oracle.py (golden `learned_logits`)."""
    torch.manual_seed(7)
    dim = 256
    rounding = LearnedRounding(dim, vocab)
    x = torch.randn(3, 17, dim) * 4                      # 51 rows: a partial 128-row tile
    ref = O.learned_logits(x, rounding.decoder.weight.detach(), rounding.decoder.bias.detach())
    with torch.no_grad():
        got = rounding.to(cuda)(x.to(cuda))
    assert got.shape == ref.shape and got.dtype == torch.float32
    got = got.cpu()
    print(f"V={vocab}: logits rel-rms {rel_rms(got, ref):.2e} max-abs {float((got - ref).abs().max()):.3e}")
    assert rel_rms(got, ref) < 5e-3 and float((got - ref).abs().max()) < 0.08
    assert torch.equal(got.argmax(-1), round_to_tokens(x.to(cuda), rounding, None, True, True).cpu())   # same GEMM as the fused argmax


def test_learned_embedding_forward_is_an_exact_gather(cuda):
    torch.manual_seed(8)
    emb = LearnedEmbedding(1000, 256)
    ids = torch.randint(0, 1000, (4, 33))
    want = emb.embeddings.weight.detach()[ids]
    with torch.no_grad():
        got = emb.to(cuda)(ids.to(cuda))
    assert torch.equal(got.cpu(), want)                 # bit-exact: a copy
    assert emb(torch.empty(0, 5, dtype=torch.long, device=cuda)).shape == (0, 5, 256)
    with pytest.raises(IndexError):
        emb(torch.tensor([[0, 1000]], device=cuda))


def test_rounding_at_the_benchmarked_vocabulary(cuda):
    """V = 256,000 (Gemma's vocabulary, the size bench.py times), n = 5 sequences of L = 64: learned and cosine
    rounding bit-exact wherever the fp32 top-2 margin exceeds the stated tolerance."""
    torch.manual_seed(9)
    dim, vocab = 256, 256_000
    rounding = LearnedRounding(dim, vocab)
    emb = LearnedEmbedding(vocab, dim)
    x = torch.randn(5, 64, dim) * 4
    logits = O.learned_logits(x, rounding.decoder.weight.detach(), rounding.decoder.bias.detach())
    sims = O.cosine_logits(x, emb.embeddings.weight.detach())
    got = round_to_tokens(x.to(cuda), rounding.to(cuda), emb.to(cuda), True, True).cpu()
    _margin_check(got, logits, 0.08, "learned V=256000")
    got_c = round_to_tokens(x.to(cuda), rounding, emb, False, True).cpu()
    _margin_check(got_c, sims, 4e-3, "cosine V=256000")
    assert int(got.max()) < vocab and int(got_c.max()) < vocab
    # and the guided mix at this size: one position, AR logits from a seeded table
    ar = torch.randn(5, vocab) * 2
    mixed = 0.7 * ar + 0.3 * logits[:, 0]               # src/shakespeare.py:449-466 at temperature 1
    r = Rounder(cuda)
    gm = r.argmax(x[:, 0].to(cuda), weight=rounding.decoder.weight, bias=rounding.decoder.bias, ar_logits=ar.to(cuda), alpha=0.3).cpu()
    _margin_check(gm, mixed, 0.05, "guided mix V=256000")


def test_guided_generate_kv_cached_lm_step_matches_full_prefix(cuda):
    """a base LM with the Hugging Face cache protocol is stepped one token at a time through its
    KV cache instead of re-running the whole prefix (ref src/shakespeare.py:447-449).  A random-init 2-layer Gemma
    stands in for google/gemma-2b-it (no hub access): both ways of stepping it must produce the same tokens up to the
    first low-margin decision, and the cached path must call the LM with ONE token per position."""
    from transformers import GemmaConfig, GemmaForCausalLM
    torch.manual_seed(11)
    dim, vocab, B, L = 256, 1000, 4, 32
    cfg = GemmaConfig(vocab_size=vocab, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=1, head_dim=16, max_position_embeddings=128)
    lm = GemmaForCausalLM(cfg).to(cuda).eval()
    rounding, emb = LearnedRounding(dim, vocab).to(cuda), LearnedEmbedding(vocab, dim).to(cuda)
    z = (torch.randn(B, L, dim) * 4).to(cuda)
    seen = []
    orig = lm.forward

    def spy(*a, **k):
        ids = k.get("input_ids", a[0] if a else None)
        seen.append(int(ids.shape[1]))
        return orig(*a, **k)

    lm.forward = spy
    cached = torch.tensor(guided_generate(lm, rounding, _Tok(), emb, z, alpha=0.3, temperature=0.7))
    assert seen == [1] * L                                  # one token per position
    seen.clear()
    full = torch.tensor(guided_generate(lm, rounding, _Tok(), emb, z, alpha=0.3, temperature=0.7, use_kv_cache=False))
    assert seen == list(range(1, L + 1))                    # the reference's O(L^2) re-forward
    agree = float((cached == full).float().mean())
    print("kv-cached vs full-prefix token agreement", agree)
    # identical logits up to fp rounding: a sequence may only part ways at an (extremely rare) near-tie
    assert agree > 0.97
