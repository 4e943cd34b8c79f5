"""GPU parity of the Shakespeare training step (row f2; ref src/shakespeare.py:221-250) against

* the REFERENCE's own step recorded in tests/golden/text_train_golden.pt (losses, gradient / updated-parameter digests,
  dropout 0), and
* the CPU fp32 oracle (oracle/text_train_oracle.py) with the library's Philox dropout masks regenerated in numpy.

Tolerances (stated): every contraction runs on bf16 operands with fp32 accumulation, activations between kernels are
fp32.  Losses agree to 2e-3 relative; a parameter's gradient agrees to 3e-2 relative L2 with cosine >= 0.9995 (bf16
rounding of both GEMM operands, relative error ~2^-9 per product, partly averaged out by the reduction); gradients
that are mathematically zero (the key bias of softmax attention) are compared absolutely.
"""
import ctypes
import sys
from pathlib import Path

import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import text_train_oracle as TO
from tests.helpers import rel_rms
from tinydiffusionmodels_b200 import _lib
from tinydiffusionmodels_b200.shakespeare import LearnedEmbedding, LearnedRounding, TinyTransformer
from tinydiffusionmodels_b200.text_train import TextTrainer, _LAYER_KEYS

pytestmark = pytest.mark.gpu
TAB = O.make_tables()
GOLD = Path(__file__).resolve().parent / "golden" / "text_train_golden.pt"
sys.path.insert(0, str(GOLD.parent))
from make_golden_text_train import sample_index  # noqa: E402


def _modules(dim, vocab, depth=3, dropout=0.1, seed=0, emb_scale=25.0):
    torch.manual_seed(seed)
    m = TinyTransformer(dim, depth=depth, dropout=dropout)
    r = LearnedRounding(dim, vocab)
    e = LearnedEmbedding(vocab, dim)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
        e.embeddings.weight.mul_(emb_scale)
    return m, r, e


def _cpu_state(m, r, e):
    return ({k: v.detach().cpu().clone() for k, v in m.state_dict().items()},
            r.decoder.weight.detach().cpu().clone(), r.decoder.bias.detach().cpu().clone(),
            e.embeddings.weight.detach().cpu().clone())


def _named_grads(tr):
    """{oracle key: gradient tensor (CPU)} from the trainer's flat gradient."""
    out = {}
    i = 0
    for li in range(tr.depth):
        for k in _LAYER_KEYS:
            p = tr._params[i]
            out[f"model.encoder.layers.{li}.{k}"] = tr.grad_of(i, tuple(p.shape)).cpu()
            i += 1
    names = ["model.time_emb.weight", "model.time_emb.bias", "decoder.weight", "decoder.bias", "embeddings.weight"]
    for n in names[: len(tr._params) - i]:
        p = tr._params[i]
        out[n] = tr.grad_of(i, tuple(p.shape)).cpu()
        i += 1
    return out


def _compare_grads(got: dict, want: dict, rel=3e-2, cos_min=0.9995):
    worst = ("", 0.0)
    for k, w in want.items():
        g = got[k]
        assert g.shape == w.shape, k
        assert torch.isfinite(g).all(), k
        wn = float(w.double().norm())
        if wn < 1e-7 * w.numel() ** 0.5:            # mathematically zero (key bias): absolute
            assert float(g.abs().max()) < 1e-5, (k, float(g.abs().max()))
            continue
        err = float((g.double() - w.double()).norm()) / wn
        cos = float((g.double() * w.double()).sum() / (g.double().norm() * w.double().norm()))
        if err > worst[1]:
            worst = (k, err)
        assert err < rel and cos > cos_min, (k, err, cos)
    return worst


def test_one_step_equals_the_reference_recording(cuda):
    g = torch.load(GOLD)
    m = TinyTransformer(g["D"], depth=1, dropout=0.0)
    r = LearnedRounding(g["D"], g["V"])
    e = LearnedEmbedding(g["V"], g["D"])
    m.load_state_dict(g["init"]["model"]); r.load_state_dict(g["init"]["rounding"]); e.load_state_dict(g["init"]["embedding"])
    m.to(cuda); r.to(cuda); e.to(cuda)
    tr = TextTrainer(m, r, e, cuda, g["B"], g["L"], lr=g["lr"], weight_decay=g["wd"], dropout=0.0, use_graph=False)
    tr.rw_dev.fill_(g["rw"])
    losses = tr.loss_and_grads(g["ids"], g["t"], g["noise"]).cpu()
    print("losses", losses.tolist(), "reference", g["losses"].tolist())
    assert torch.allclose(losses, g["losses"], rtol=2e-3)
    got = _named_grads(tr)
    ref = {**{"model." + k: v for k, v in g["grads"]["model"].items()},
           "decoder.weight": g["grads"]["rounding"]["decoder.weight"], "decoder.bias": g["grads"]["rounding"]["decoder.bias"],
           "embeddings.weight": g["grads"]["embedding"]["embeddings.weight"]}
    for k, d in ref.items():
        flat = got[k].reshape(-1)
        s = flat[sample_index(flat.numel())]
        nrm = d["norm"]
        if nrm < 1e-7 * flat.numel() ** 0.5:
            assert float(flat.abs().max()) < 1e-5, k
            continue
        assert abs(float(flat.double().norm()) - nrm) < 2e-2 * nrm, (k, float(flat.double().norm()), nrm)
        err = float((s - d["sample"]).double().norm() / d["sample"].double().norm().clamp_min(1e-30))
        assert err < 3e-2, (k, err)


@pytest.mark.parametrize("dim,depth,batch,seq,vocab,p", [(256, 3, 2, 64, 1000, 0.1), (256, 2, 3, 128, 520, 0.1),
                                                          (512, 1, 2, 64, 768, 0.25)])
def test_losses_and_gradients_match_the_oracle_with_dropout(cuda, dim, depth, batch, seq, vocab, p):
    m, r, e = _modules(dim, vocab, depth, p, seed=dim + depth)
    sd, dw, db, ew = _cpu_state(m, r, e)
    gen = torch.Generator().manual_seed(11)
    ids = torch.randint(0, vocab, (batch, seq), generator=gen)
    ids[0, :5] = ids[0, 5]
    t = torch.randint(0, 1000, (batch,), generator=gen)
    noise = torch.randn(batch, seq, dim, generator=gen)
    m.to(cuda); r.to(cuda); e.to(cuda)
    tr = TextTrainer(m, r, e, cuda, batch, seq, dropout=p, seed=77, use_graph=False)
    tr.rw_dev.fill_(0.6)
    tr.step_dev.fill_(5)
    losses = tr.loss_and_grads(ids, t, noise).cpu()
    want_l, want_g = TO.text_losses_and_grads(sd, dw, db, ew, ids, t, noise, TAB, rounding_weight=0.6, dropout=p, seed=77, step=5)
    print("losses", losses.tolist(), "oracle", [float(x) for x in want_l])
    assert torch.allclose(losses, torch.stack(want_l), rtol=2e-3)
    worst = _compare_grads(_named_grads(tr), want_g)
    print("worst gradient", worst)


def test_in_kernel_timesteps_noise_and_input_dropout(cuda):
    """t and the noise drawn in-kernel equal the numpy Philox restatement; h0 = dropout(q_sample + time embedding)."""
    dim, vocab, batch, seq, p = 256, 300, 4, 64, 0.1
    m, r, e = _modules(dim, vocab, 1, p)
    sd, dw, db, ew = _cpu_state(m, r, e)
    ids = torch.randint(0, vocab, (batch, seq), generator=torch.Generator().manual_seed(2))
    m.to(cuda); r.to(cuda); e.to(cuda)
    tr = TextTrainer(m, r, e, cuda, batch, seq, dropout=p, seed=123456789012345, use_graph=False)
    tr.step_dev.fill_(9)
    tr.loss_and_grads(ids)
    lay = (ctypes.c_int64 * 25)()
    _lib.check(tr.lib.tdm_text_train_debug_layout(batch, seq, dim, 1, vocab, lay), "layout")
    n = batch * seq * dim
    t_dev = tr.ws[lay[1]:lay[1] + batch * 8].view(torch.int64).cpu()
    x0 = tr.ws[lay[2]:lay[2] + n * 4].view(torch.float32).view(batch, seq, dim).cpu()
    noise = tr.ws[lay[3]:lay[3] + n * 4].view(torch.float32).view(batch, seq, dim).cpu()
    h0 = tr.ws[lay[4]:lay[4] + n * 4].view(torch.float32).view(batch, seq, dim).cpu()
    t_want = TO.draw_t(batch, tr.seed, 9)
    assert torch.equal(t_dev, t_want)
    assert torch.equal(x0, ew[ids])
    z_want = TO.train_noise(batch, seq, dim, tr.seed, 9)
    assert float((noise - z_want).abs().max()) < 2e-5      # device fast log / sincos vs libm (as tests/test_gpu_elementwise.py)
    col = {}
    TO.transformer_forward_train(sd, O.q_sample(ew[ids], t_want, noise, TAB), t_want, p, tr.seed, 9, collect=col)
    assert float((h0 - col["h0"]).abs().max()) < 1e-5
    assert torch.equal(h0 == 0, col["h0"] == 0)            # the same elements are dropped


def test_eval_mode_losses_and_frozen_embeddings(cuda):
    dim, vocab, batch, seq = 256, 640, 2, 64
    m, r, e = _modules(dim, vocab, 2, 0.1, seed=5)
    sd, dw, db, ew = _cpu_state(m, r, e)
    gen = torch.Generator().manual_seed(4)
    ids = torch.randint(0, vocab, (batch, seq), generator=gen)
    t = torch.randint(0, 1000, (batch,), generator=gen)
    noise = torch.randn(batch, seq, dim, generator=gen)
    m.to(cuda); r.to(cuda)
    table = ew.to(cuda)
    tr = TextTrainer(m, r, table, cuda, batch, seq, use_learned_embeddings=False, dropout=0.1, seed=1, use_graph=False)
    tr._objective(ids.to(cuda), None, t.to(cuda), noise.to(cuda))          # eval: no dropout although p = 0.1
    want_l, _ = TO.text_losses_and_grads(sd, dw, db, ew, ids, t, noise, TAB, dropout=0.0, want_grads=False)
    assert torch.allclose(tr.losses.cpu(), torch.stack(want_l), rtol=2e-3)
    tr.step_dev.fill_(2)
    tr.loss_and_grads(ids, t, noise)
    _, want_g = TO.text_losses_and_grads(sd, dw, db, ew, ids, t, noise, TAB, dropout=0.1, seed=1, step=2, learn_embeddings=False)
    assert "embeddings.weight" not in want_g
    _compare_grads(_named_grads(tr), want_g)


def test_adamw_with_device_learning_rate_equals_the_oracle_update(cuda):
    n = 4099
    gen = torch.Generator().manual_seed(0)
    p = torch.randn(n, generator=gen)
    g = torch.randn(n, generator=gen) * 1e-2
    mm = torch.randn(n, generator=gen) * 1e-3
    v = torch.rand(n, generator=gen) * 1e-4
    lib = _lib.load()
    dp, dg, dm, dv = (x.to(cuda).clone() for x in (p, g, mm, v))
    lr = torch.tensor([3e-4], device=cuda)
    step = torch.tensor([7], dtype=torch.int64, device=cuda)
    _lib.check(lib.tdm_adamw_flat_lr(dp.data_ptr(), dg.data_ptr(), dm.data_ptr(), dv.data_ptr(), n, lr.data_ptr(), 0.9, 0.999,
                                     1e-8, 1e-2, 1.0, step.data_ptr(), _lib.stream_ptr(cuda)), "adamw")
    wp, wm, wv = O.adamw_step(p, g, mm, v, 7, lr=3e-4, wd=1e-2)
    assert torch.allclose(dp.cpu(), wp, rtol=0, atol=2e-7)
    assert torch.allclose(dm.cpu(), wm, rtol=1e-5, atol=1e-9) and torch.allclose(dv.cpu(), wv, rtol=1e-5, atol=1e-12)   # FMA contraction


def test_replayed_graph_trains_and_matches_the_eager_sequence(cuda):
    dim, vocab, batch, seq = 256, 512, 4, 64
    ids = torch.randint(0, vocab, (batch, seq), generator=torch.Generator().manual_seed(8))
    finals = []
    for use_graph in (True, False):
        m, r, e = _modules(dim, vocab, 2, 0.1, seed=3)
        m.to(cuda); r.to(cuda); e.to(cuda)
        tr = TextTrainer(m, r, e, cuda, batch, seq, lr=1e-3, seed=42, use_graph=use_graph)
        hist = []
        for k in range(12):
            hist.append(tr.step(ids, lr=1e-3 * min(1.0, (k + 1) / 4), rounding_weight=1.0).cpu().clone())
        finals.append((tr.flat.cpu().clone(), torch.stack(hist)))
        assert int(tr.step_dev) == 12
        assert hist[-1][2] < hist[0][2] and hist[-1][1] < hist[0][1], [h.tolist() for h in hist]
        assert torch.isfinite(tr.flat).all()
    # same seeds, same steps: the graph replays exactly the eager launches (the embedding scatter's float atomics are
    # the one order-dependent sum, so the comparison is to 1e-5 and not bit-wise)
    assert torch.allclose(finals[0][1], finals[1][1], rtol=1e-4, atol=1e-5)
    assert float((finals[0][0] - finals[1][0]).abs().max()) < 1e-4


def test_train_loop_writes_reference_format_checkpoints_and_samplers_see_new_weights(cuda, tmp_path):
    from torch.utils.data import DataLoader

    from tinydiffusionmodels_b200 import shakespeare as S

    dim, vocab, batch, seq = 256, 512, 4, 64
    m, r, e = _modules(dim, vocab, 1, 0.1, seed=9)
    m.to(cuda); r.to(cuda); e.to(cuda)
    m.eval()
    x = torch.randn(batch, seq, dim, generator=torch.Generator().manual_seed(1)).to(cuda)
    t = torch.full((batch,), 500, device=cuda)
    with torch.no_grad():
        before = m(x, t).clone()
    gen = torch.Generator().manual_seed(0)
    chunks = torch.randint(0, vocab, (batch * 6, seq), generator=gen)
    ck = str(tmp_path / "text_ckpt.pth")
    S.train(m, r, e, DataLoader(chunks[batch:], batch_size=batch), DataLoader(chunks[:batch], batch_size=batch), cuda,
            ckpt_path=ck, epochs=2, lr=1e-3, warmup_steps=2, seed=5, log_every=0)
    saved = torch.load(ck, map_location="cpu")
    assert set(saved) == {"diffusion_model", "rounding_fn", "epoch", "final_training", "embedding_fn"}
    assert set(saved["diffusion_model"]) == set(m.state_dict()) and saved["epoch"] == 2
    assert (tmp_path / "text_ckpt_best.pth").exists()
    m.eval()
    with torch.no_grad():
        after = m(x, t)
    assert rel_rms(after.cpu(), before.cpu()) > 1e-4          # the sampler's packed weights were refreshed
    want = O.transformer_forward({k: v.float() for k, v in saved["diffusion_model"].items()}, x.cpu(), t.cpu())
    assert rel_rms(after.cpu(), want) < 1e-2


def test_out_of_range_token_ids_are_reported(cuda):
    m, r, e = _modules(256, 300, 1, 0.0)
    m.to(cuda); r.to(cuda); e.to(cuda)
    tr = TextTrainer(m, r, e, cuda, 2, 64, use_graph=False)
    ids = torch.randint(0, 300, (2, 64))
    tr.loss_and_grads(ids)
    tr.check_token_ids()                       # clean batch: no error
    ids[1, 7] = 300
    tr.loss_and_grads(ids)
    with pytest.raises(IndexError):
        tr.check_token_ids()
    tr.check_token_ids()                       # the flag was cleared
