"""Golden vectors of ONE Shakespeare training step, recorded by running the REAL reference (/root/reference).

Run in the authoring container only:   python tests/golden/make_golden_text_train.py

What is recorded (vocab 512, width 256, depth 1, batch 2 x 64, dropout 0 — the reference's dropout masks come from torch's
global generator and cannot be injected into nn.TransformerEncoderLayer's fused attention, so the pinned case is the
mask-free one; the oracle's dropout sites are restated from torch/nn/modules/transformer.py):
  * the three modules' state_dicts (reference classes, reference init), token ids, t, noise;
  * the losses of src/shakespeare.py:233-244 computed by the reference's own modules and q_sample in TRAIN mode;
  * the gradients of total_loss w.r.t. every parameter (loss.backward());
  * every parameter after one optim.step() of torch.optim.AdamW(params, lr, weight_decay) (:196, :246-248).
tests/test_oracle.py pins oracle/text_train_oracle.py to this file; the GPU tests compare CUDA with the oracle.
"""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference  # noqa: E402


def sample_index(numel: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(numel)
    return torch.randperm(numel, generator=g)[:4096].sort().values


def digest(x: torch.Tensor) -> dict:
    flat = x.detach().reshape(-1)
    return {"norm": flat.double().norm().item(), "sample": flat[sample_index(flat.numel())].clone()}


def main():
    _, ref = import_reference()
    V, D, B, L = 512, 256, 2, 64
    lr, wd, rw = 1e-3, 1e-2, 0.7
    torch.manual_seed(0)
    model = ref.TinyTransformer(D, depth=1, dropout=0.0).train()
    rounding = ref.LearnedRounding(D, V).train()
    emb = ref.LearnedEmbedding(V, D).train()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "norm" in n:
                p.add_(0.1 * torch.randn_like(p))
        emb.embeddings.weight.mul_(25.0)   # unit-scale embeddings, so q_sample's two terms are comparable
    init = {"model": {k: v.clone() for k, v in model.state_dict().items()},
            "rounding": {k: v.clone() for k, v in rounding.state_dict().items()},
            "embedding": {k: v.clone() for k, v in emb.state_dict().items()}}
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, V, (B, L), generator=g)
    ids[0, :8] = ids[0, 8]          # repeated tokens: the embedding gradient must accumulate
    t = torch.randint(0, ref.T, (B,), generator=g)
    noise = torch.randn(B, L, D, generator=g)

    params = list(model.parameters()) + list(rounding.parameters()) + list(emb.parameters())   # :191-194
    optim = torch.optim.AdamW(params, lr=lr, weight_decay=wd)                                # :196
    x0 = emb(ids)                                                                             # :226
    x_noisy = ref.q_sample(x0, t, noise)                                                      # :232
    pred = model(x_noisy, t)                                                                  # :233
    diff = F.mse_loss(pred, noise)                                                            # :236
    logits = rounding(x0)                                                                     # :239
    rnd = F.cross_entropy(logits.reshape(-1, logits.size(-1)), ids.reshape(-1))               # :241
    total = diff + rw * rnd                                                                   # :244
    optim.zero_grad()
    total.backward()
    # the fixture stays small: of every gradient / updated parameter it keeps the L2 norm and a fixed pseudo-random
    # sample of up to 4,096 elements (indices from sample_index below), not the whole tensor
    grads = {"model": {k: digest(p.grad) for k, p in model.named_parameters()},
             "rounding": {k: digest(p.grad) for k, p in rounding.named_parameters()},
             "embedding": {k: digest(p.grad) for k, p in emb.named_parameters()}}
    optim.step()
    after = {"model": {k: digest(p.detach()) for k, p in model.named_parameters()},
             "rounding": {k: digest(p.detach()) for k, p in rounding.named_parameters()},
             "embedding": {k: digest(p.detach()) for k, p in emb.named_parameters()}}
    out = {"V": V, "D": D, "B": B, "L": L, "lr": lr, "wd": wd, "rw": rw, "init": init, "ids": ids, "t": t,
           "noise": noise, "losses": torch.stack([diff.detach(), rnd.detach(), total.detach()]), "grads": grads,
           "after": after, "pred": pred.detach()}
    torch.save(out, HERE / "text_train_golden.pt")
    print("wrote", HERE / "text_train_golden.pt", f"losses {out['losses'].tolist()}")


if __name__ == "__main__":
    main()
