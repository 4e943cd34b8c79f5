"""Generate golden input/output vectors by running the REAL reference (/root/reference).

Run in the authoring container only (the reference does not travel to the GPU box):
    python tests/golden/make_golden.py
It imports the unmodified reference through a 6-line google.cloud.storage stub (the reference
imports that package at module import), feeds seeded inputs to the reference's own
functions and stores inputs + outputs.  tests/test_oracle.py pins oracle/ddpm_oracle.py to these
files; the GPU parity tests then compare the CUDA path with the oracle.
"""
import sys
import types
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent


def import_reference():
    g, gc, gs = types.ModuleType("google"), types.ModuleType("google.cloud"), types.ModuleType("google.cloud.storage")
    gs.Client = type("Client", (), {})
    gc.storage = gs
    g.cloud = gc
    sys.modules.update({"google": g, "google.cloud": gc, "google.cloud.storage": gs})
    sys.dont_write_bytecode = True
    saved = {k: v for k, v in sys.modules.items() if k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, "/root/reference")
    try:
        import src.mnist as ref_mnist
        import src.shakespeare as ref_text
    finally:
        sys.path.remove("/root/reference")
        for k in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return ref_mnist, ref_text


class _InjectNoise:
    """Make module.torch.randn_like return a prepared tensor (the reference draws its own noise)."""

    def __init__(self, z):
        self.z = z

    def __enter__(self):
        self._orig = torch.randn_like
        torch.randn_like = lambda x, *a, **k: self.z.to(x.dtype)
        return self

    def __exit__(self, *exc):
        torch.randn_like = self._orig


def make_mnist(ref):
    torch.manual_seed(0)
    model = ref.SimpleUNet().eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    B = 4
    x = torch.randn(B, 1, 28, 28, generator=g)
    noise = torch.randn(B, 1, 28, 28, generator=g)
    z = torch.randn(B, 1, 28, 28, generator=g)
    t = torch.tensor([999, 500, 37, 0])
    out = {
        "tables": {k: getattr(ref, k).clone() for k in ("betas", "alphas", "alphas_cumprod", "sqrt_alphas_cumprod",
                                                        "sqrt_one_minus_alphas_cumprod")},
        "state_dict": sd, "x": x, "noise": noise, "z": z, "t": t,
        "q_sample": ref.q_sample(x, t, noise),
    }
    with torch.no_grad():
        out["unet"] = model(x, t)
        for name, tv in (("p_sample_t700", 700), ("p_sample_t0", 0)):
            tt = torch.full((B,), tv, dtype=torch.long)
            with _InjectNoise(z):
                out[name] = ref.p_sample(model, x, tt)
        # a short trajectory with injected noises: 6 reverse steps t = 5..0
        zs = torch.randn(6, B, 1, 28, 28, generator=g)
        xt = x.clone()
        for i in reversed(range(6)):
            with _InjectNoise(zs[i]):
                xt = ref.p_sample(model, xt, torch.full((B,), i, dtype=torch.long))
        out["zs6"] = zs
        out["traj6"] = xt
        out["unit_range"] = (xt.clamp(-1, 1) + 1) / 2
    # one training step exactly as src/mnist.py:153-159, AdamW(lr=1e-3)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    x0 = torch.rand(B, 1, 28, 28, generator=g) * 2 - 1
    tt = torch.tensor([3, 250, 600, 999])
    x_noisy = ref.q_sample(x0, tt, noise)
    loss = torch.nn.functional.mse_loss(model(x_noisy, tt), noise)
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    opt.step()
    out["train"] = {
        "x0": x0, "t": tt, "loss": loss.detach(),
        "grad_norms": {k: v.norm() for k, v in grads.items()},
        "grads_small": {k: v for k, v in grads.items() if v.numel() <= 64},
        "grad_rb4_conv1_w": grads["rb4.conv1.weight"],
        "after_norms": {k: v.detach().norm() for k, v in model.named_parameters()},
        "after_small": {k: v.detach().clone() for k, v in model.named_parameters() if v.numel() <= 64},
    }
    torch.save(out, HERE / "mnist_golden.pt")
    print("mnist_golden.pt", (HERE / "mnist_golden.pt").stat().st_size)


class TinyLM(torch.nn.Module):
    """Deterministic stand-in for base_lm in guided_generate (Gemma cannot be downloaded)."""

    def __init__(self, vocab, seed):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.table = torch.nn.Parameter(torch.randn(vocab, vocab, generator=g))

    def forward(self, input_ids):
        logits = self.table[input_ids] + 0.1 * torch.cumsum(self.table[input_ids], dim=1)
        return types.SimpleNamespace(logits=logits)


class Tok:
    bos_token_id = 2
    eos_token_id = 1

    def batch_decode(self, ids, skip_special_tokens=True):
        return [" ".join(str(int(i)) for i in row) for row in ids]


def make_text(ref):
    dim, V, L, B = 32, 101, 8, 3
    torch.manual_seed(0)
    model = ref.TinyTransformer(dim).eval()
    rounding = ref.LearnedRounding(dim, V).eval()
    emb = ref.LearnedEmbedding(V, dim).eval()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, L, dim, generator=g)
    noise = torch.randn(B, L, dim, generator=g)
    z = torch.randn(B, L, dim, generator=g)
    t = torch.tensor([999, 421, 0])
    out = {
        "dim": dim, "V": V, "L": L,
        "model_sd": {k: v.clone() for k, v in model.state_dict().items()},
        "rounding_sd": {k: v.clone() for k, v in rounding.state_dict().items()},
        "emb_sd": {k: v.clone() for k, v in emb.state_dict().items()},
        "x": x, "noise": noise, "z": z, "t": t,
        "q_sample": ref.q_sample(x, t, noise),
    }
    with torch.no_grad():
        out["transformer"] = model(x, t)
        tt = torch.full((B,), 300, dtype=torch.long)
        with _InjectNoise(z):
            out["p_sample_t300"] = ref.p_sample(model, x, tt)
        out["p_sample_t0"] = ref.p_sample(model, x, torch.zeros(B, dtype=torch.long))
        logits = rounding(x)
        out["learned_logits"] = logits
        out["learned_tokens"] = logits.argmax(-1)
        E = emb.get_embedding_matrix()
        sims = torch.matmul(torch.nn.functional.normalize(x, dim=2), torch.nn.functional.normalize(E, dim=1).T)
        out["cosine_sims"] = sims
        out["cosine_tokens"] = sims.argmax(-1)
        lm = TinyLM(V, 5).eval()
        out["lm_table"] = lm.table.detach().clone()
        out["guided_learned"] = ref.guided_generate(lm, rounding, Tok(), emb, x, alpha=0.3, use_learned_rounding=True)
        out["guided_cosine"] = ref.guided_generate(lm, rounding, Tok(), emb, x, alpha=0.3, temperature=0.7,
                                                   use_learned_rounding=False, use_learned_embeddings=True)
    torch.save(out, HERE / "text_golden.pt")
    print("text_golden.pt", (HERE / "text_golden.pt").stat().st_size)


if __name__ == "__main__":
    ref_mnist, ref_text = import_reference()
    make_mnist(ref_mnist)
    make_text(ref_text)
