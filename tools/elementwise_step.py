"""Two launches of every elementwise kernel at the benchmark shapes (target for ncu; tools/evidence_r02.sh).

    python tools/elementwise_step.py        # 16384 x 784 (MNIST) and 512 x 64 x 256 (text), injected and Philox noise
Per shape and round the launch order is: q_sample<injected>, q_sample<Philox>, reverse_step<injected>,
reverse_step<Philox>, randn, to_unit_range.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200 import ops

dev = torch.device("cuda:0")
for shape in ((16384, 1, 28, 28), (512, 64, 256)):
    B = shape[0]
    x = torch.randn(shape, device=dev)
    e = torch.randn(shape, device=dev)
    z = torch.randn(shape, device=dev)
    t = torch.randint(1, 1000, (B,), device=dev)
    out = torch.empty_like(x)
    for rep in range(2):   # the second round is the warm one; ncu captures both
        ops.q_sample(x, t, z)
        ops.q_sample(x, t, None, seed=1)
        ops.reverse_step(x, e, t, z, out=out)
        ops.reverse_step(x, e, t, None, out=out, seed=1)
        ops.randn(shape, dev, seed=3)
        ops.to_unit_range(x)
torch.cuda.synchronize()
print("ok")
