"""A few text reverse steps at a given batch (target for ncu)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.shakespeare import TinyTransformer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = TinyTransformer(dim).to(dev).eval()
eng = m.engine(B, 64)
x = torch.randn(B, 64, dim, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.int64)
eng.load_state(x, t)
for _ in range(3):
    eng.p_sample_inplace(t, None, seed=1)
torch.cuda.synchronize()
print("ok")
