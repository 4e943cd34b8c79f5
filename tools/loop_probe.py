"""Time the graph-replayed reverse loop (the product path: sample_loop) at several batch sizes.

    python tools/loop_probe.py 25 64 1024 16384        # us per reverse step, samples/s at T=1000
Unlike perf_probe.py there are no events between the launches, so launch gaps / PDL overlap count.
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.mnist import SimpleUNet, sample_loop

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SimpleUNet().to(dev).eval()
for B in [int(a) for a in sys.argv[1:]] or [25, 64, 1024, 16384]:
    steps = 1000 if B <= 4096 else 300
    x = torch.randn(B, 1, 28, 28, device=dev)
    sample_loop(model, x.clone(), seed=1, steps=50)   # warm-up (engine, attributes, clocks)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = sample_loop(model, x, seed=1, steps=steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)   # includes one graph capture (~1 ms)
    us = ms * 1e3 / steps
    print(f"B={B:6d}  {us:8.1f} us/step  {B / (us * 1e-6) / 1000:9.1f} samples/s(T=1000)  finite={bool(torch.isfinite(out).all())}")
    model._engine = None
