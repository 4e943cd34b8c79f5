"""Repeat one fused p_sample on identical inputs many times; report any run whose output or
intermediate activations differ from the first run (race detector)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.mnist import SimpleUNet
from tinydiffusionmodels_b200.unet_engine import read_activation

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = SimpleUNet().to(dev).eval()
eng = m.engine(B)
x = torch.randn(B, 1, 28, 28, device=dev) * 50
t = torch.full((B,), 500, device=dev, dtype=torch.int64)
names = ("t1", "cat", "p1", "t2", "s2", "h2", "t3", "t4", "s4")
ref = eng.p_sample(x, t, None, seed=3).clone()
ref_act = {n: read_activation(eng, n, B).clone() for n in names}
bad = 0
for i in range(N):
    out = eng.p_sample(x, t, None, seed=3)
    if not torch.equal(out, ref):
        bad += 1
        d = (out - ref).abs()
        diffs = {n: int((read_activation(eng, n, B) != ref_act[n]).sum()) for n in names}
        if bad <= 5:
            idx = d.flatten().argmax().item()
            print(f"run {i}: {int((d > 0).sum())} elements differ, max {float(d.max()):.3e} at flat {idx} "
                  f"(b={idx // 784}, y={(idx % 784) // 28}, x={idx % 28}); layers differing: "
                  f"{ {k: v for k, v in diffs.items() if v} }")
print(f"B={B}: {bad} of {N} runs differ from the first")
