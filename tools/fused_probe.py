"""Fused 28x28 blocks (csrc/resblock_tc.cuh) vs the layer-by-layer kernels: agreement, then per-kernel times.

    python tools/fused_probe.py [batch ...]      # default 1 7 64 300 ; timing at 16384
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from oracle import ddpm_oracle as O
from tests.helpers import random_unet_state_dict, rel_rms
from tinydiffusionmodels_b200 import _lib
from tinydiffusionmodels_b200.unet_engine import UNetEngine

dev = torch.device("cuda:0")
lib = _lib.load()
sd = random_unet_state_dict(31)
for B in [int(a) for a in sys.argv[1:]] or [1, 7, 64, 300]:
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, 1, 28, 28, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    eng = UNetEngine(dev, B)
    eng.load_state_dict(sd)
    lib.tdm_unet_set_fused(1)
    f = eng.forward(x.to(dev), t.to(dev)).cpu()
    lib.tdm_unet_set_fused(0)
    p = eng.forward(x.to(dev), t.to(dev)).cpu()
    ref = O.unet_forward(sd, x, t)
    d = (f - p).abs()
    print(f"B={B}: fused-vs-plain rel-rms {rel_rms(f, p):.3e} max {float(d.max()):.3e}; fused-vs-oracle {rel_rms(f, ref):.3e}; "
          f"plain-vs-oracle {rel_rms(p, ref):.3e}; finite={bool(torch.isfinite(f).all())}", flush=True)
    if rel_rms(f, p) > 1e-2:
        bad = (d > 10 * float((p - ref).abs().max())).nonzero()
        print("  first bad (b, y, x):", [tuple(int(v) for v in r[[0, 2, 3]]) for r in bad[:12]], " count", len(bad))
        per_img = d.view(B, -1).max(1).values
        print("  per-image max diff:", [round(float(v), 4) for v in per_img[:16]])
        rows = d[0, 0].max(1).values
        print("  image0 per-row max diff:", [round(float(v), 3) for v in rows])

B = 16384
torch.manual_seed(0)
eng = UNetEngine(dev, B)
eng.load_state_dict(sd)
x = torch.randn(B, 1, 28, 28, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.int64)
for mode in (1, 0, 1):
    lib.tdm_unet_set_fused(mode)
    ks = [eng.profile_p_sample(x, t, seed=1) for _ in range(8)][3:]
    ms = [sum(k[i] for k in ks) / len(ks) for i in range(9)]
    print(("fused  " if mode else "plain  ") + " ".join(f"{n}={v:.3f}" for n, v in zip(eng.KERNEL_NAMES, ms)) + f"  sum={sum(ms):.3f} ms", flush=True)
