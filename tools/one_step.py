"""Run a few fused p_sample steps at a given batch (target for ncu captures)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.mnist import SimpleUNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SimpleUNet().to(dev).eval()
eng = model.engine(B)
x = torch.randn(B, 1, 28, 28, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.int64)
for _ in range(steps):
    eng.p_sample(x, t, None, out=x, seed=1)
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
