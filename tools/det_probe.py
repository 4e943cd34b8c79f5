import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from tinydiffusionmodels_b200 import ops
from tinydiffusionmodels_b200.mnist import SimpleUNet, sample_loop

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = SimpleUNet().to(dev).eval()
x_T = ops.randn((64, 1, 28, 28), dev, seed=11)
for graph in (False, True):
    for steps in (10, 50, 200, 1000):
        a = sample_loop(m, x_T.clone(), seed=11, steps=steps, use_graph=graph)
        b = sample_loop(m, x_T.clone(), seed=11, steps=steps, use_graph=graph)
        d = (a - b).abs()
        print(f"graph={graph} steps={steps}: equal={torch.equal(a, b)} ndiff={int((d > 0).sum())} max={float(d.max()):.3e} rms={float(a.pow(2).mean().sqrt()):.3e}")
# fresh engine for each run
for steps in (200, 1000):
    m._engine = None
    a = sample_loop(m, x_T.clone(), seed=11, steps=steps, use_graph=False)
    m._engine = None
    b = sample_loop(m, x_T.clone(), seed=11, steps=steps, use_graph=False)
    print(f"fresh engines steps={steps}: equal={torch.equal(a, b)}")
