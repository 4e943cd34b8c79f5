"""Quick per-kernel timing of one fused p_sample at several batch sizes (development aid)."""
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.mnist import SimpleUNet
from tinydiffusionmodels_b200.unet_engine import UNetEngine

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SimpleUNet().to(dev).eval()
batches = [int(a) for a in sys.argv[1:]] or [64, 256, 1024, 4096, 16384]
for B in batches:
    eng = model.engine(B)
    x = torch.randn(B, 1, 28, 28, device=dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.int64)
    rows = [eng.profile_p_sample(x, t, seed=1) for _ in range(12)][4:]
    ms = [statistics.mean(r[i] for r in rows) for i in range(9)]
    tot = sum(ms)
    tf = 129_002_880 * B / (tot * 1e-3) / 1e12
    print(f"B={B:6d} step {tot*1e3:8.1f} us  {B/(tot*1e-3)/1000:8.1f} samples/s(T=1000)  {tf:6.1f} TF/s | "
          + " ".join(f"{n[:9]}={v*1e3:.0f}" for n, v in zip(UNetEngine.KERNEL_NAMES, ms)))
    model._engine = None
    del eng
    torch.cuda.empty_cache()
