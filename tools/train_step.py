"""A few eager training steps at a given batch (target for ncu)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.mnist import SimpleUNet
from tinydiffusionmodels_b200.unet_train import UNetTrainer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = SimpleUNet().to(dev)
tr = UNetTrainer(m, max_batch=B, seed=1, use_graph=False)
x = torch.rand(B, 1, 28, 28, device=dev) * 2 - 1
for _ in range(3):
    tr.step(x)
torch.cuda.synchronize()
print("ok")
