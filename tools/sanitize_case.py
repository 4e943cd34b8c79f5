"""Small end-to-end invocation of every kernel family (target for compute-sanitizer)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200 import ops
from tinydiffusionmodels_b200.mnist import SimpleUNet, sample_loop
from tinydiffusionmodels_b200.shakespeare import LearnedRounding, TinyTransformer
from tinydiffusionmodels_b200.text_engine import Rounder
from tinydiffusionmodels_b200.unet_train import UNetTrainer

dev = torch.device("cuda:0")
torch.manual_seed(0)
for B in (1, 3, 37):
    m = SimpleUNet().to(dev).eval()
    x = ops.randn((B, 1, 28, 28), dev, seed=1)
    t = torch.full((B,), 7, device=dev, dtype=torch.int64)
    with torch.no_grad():
        m(x, t)
    sample_loop(m, x, seed=2, steps=3, use_graph=False)
    tr = UNetTrainer(m, max_batch=B, use_graph=False)
    tr.step(torch.rand(B, 1, 28, 28, device=dev) * 2 - 1)
for dim, B, L in ((256, 3, 64), (256, 2, 128), (2048, 1, 64)):
    tm = TinyTransformer(dim).to(dev).eval()
    e = tm.engine(B, L)
    e.sample_loop(torch.randn(B, L, dim, device=dev), seed=3, steps=2, use_graph=False)
rf = LearnedRounding(256, 1000).to(dev)
r = Rounder(dev)
z = torch.randn(3, 64, 256, device=dev)
r.argmax(z, weight=rf.decoder.weight, bias=rf.decoder.bias)
r.argmax(z[:, 0], weight=rf.decoder.weight, cosine=True, ar_logits=torch.randn(3, 1000, device=dev), alpha=0.3)
torch.cuda.synchronize()
print("sanitize case ok")
