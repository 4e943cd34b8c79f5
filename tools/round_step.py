"""ncu target: the rounding GEMM + argmax at 32,768 and 4,096 token rows and one guided-mix position at 512 sequences
(V = 256,000, width 256), three launches each.   python tools/round_step.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.shakespeare import LearnedRounding
from tinydiffusionmodels_b200.text_engine import Rounder

dev = torch.device("cuda:0")
V, dim = 256000, 256
torch.manual_seed(0)
rf = LearnedRounding(dim, V).to(dev)
r = Rounder(dev)
for rows in (32768, 4096):
    x = torch.randn(rows, dim, device=dev)
    for _ in range(3):
        r.argmax(x, weight=rf.decoder.weight, bias=rf.decoder.bias)
ar = torch.randn(512, V, device=dev)
x = torch.randn(512, dim, device=dev)
for _ in range(3):
    r.argmax(x, weight=rf.decoder.weight, bias=rf.decoder.bias, ar_logits=ar, alpha=0.3)
torch.cuda.synchronize()
print("done")
