"""Which gradient tensors are non-finite / far from the oracle after one backward (debugging aid)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from oracle import ddpm_oracle as O
from tests.helpers import random_unet_state_dict
from tinydiffusionmodels_b200.mnist import SimpleUNet
from tinydiffusionmodels_b200.unet_engine import unflatten
from tinydiffusionmodels_b200.unet_train import loss_and_flat_grad

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sd = random_unet_state_dict(0)
model = SimpleUNet().to(dev)
model.load_state_dict(sd)
g = torch.Generator().manual_seed(5)
x0 = torch.rand(B, 1, 28, 28, generator=g) * 2 - 1
t = torch.randint(0, 1000, (B,), generator=g)
noise = torch.randn(B, 1, 28, 28, generator=g)
TAB = O.make_tables()
xn = O.q_sample(x0, t, noise, TAB)
loss, flat_g, _ = loss_and_flat_grad(model, xn.to(dev), t.to(dev), noise.to(dev))
ref_loss, ref = O.mnist_loss_and_grads(sd, x0, t, noise, TAB)
got = unflatten(flat_g.cpu())
print("loss", float(loss), float(ref_loss))
for k, v in got.items():
    r = ref[k]
    rel = float((v - r).norm() / (r.norm() + 1e-12))
    print(f"{k:24s} finite={bool(torch.isfinite(v).all())} rel={rel:.3e}")
