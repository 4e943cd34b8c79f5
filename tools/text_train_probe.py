"""One Shakespeare training step (src/shakespeare.py:221-250) at the reference's sizes: vocabulary 256,000, width 256,
depth 3, seq_len 64, at several batch sizes.  Prints ms / step, sequences / s and the TFLOP/s of the three
vocabulary-sized contractions (2 * M * D * V FLOP each: log-sum-exp pass, d logits pass, dX0, dW = 4 passes).

    python tools/text_train_probe.py [batch ...]        # default 32 128
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.shakespeare import LearnedEmbedding, LearnedRounding, TinyTransformer
from tinydiffusionmodels_b200.text_train import TextTrainer

dev = torch.device("cuda:0")
V, D, L = 256000, 256, 64
for B in [int(a) for a in sys.argv[1:]] or [32, 128]:
    torch.manual_seed(0)
    m, r, e = TinyTransformer(D).to(dev), LearnedRounding(D, V).to(dev), LearnedEmbedding(V, D).to(dev)
    tr = TextTrainer(m, r, e, dev, B, L, lr=1e-4, seed=1)
    ids = torch.randint(0, V, (B, L), device=dev)
    for _ in range(3):
        tr.step(ids)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        tr.step(ids)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    M = B * L
    flop_vocab = 4 * 2.0 * M * D * V
    print(f"text train B={B} (M={M}) V={V} D={D}: {ms:.3f} ms/step  {B / ms * 1e3:.0f} sequences/s  "
          f"vocabulary GEMMs {flop_vocab / ms / 1e9:.0f} TFLOP/s-equivalent  losses {tr.losses.tolist()}")
    del tr, m, r, e
    torch.cuda.empty_cache()
