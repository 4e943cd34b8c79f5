"""One position of guided_generate (src/shakespeare.py:451-467) at several batch sizes: the GE_ARGMAX GEMM with fp32 AR
logits mixed in.  GB/s = (fp32 AR logits B x V + bf16 W V x D, each read once) / time.

    python tools/guided_mix_probe.py [B ...]        # default 64 128 512
"""
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
for B in [int(a) for a in sys.argv[1:]] or [64, 128, 512]:
    ar = torch.randn(B, V, device=dev)
    x = torch.randn(B, dim, device=dev)
    fn = lambda: r.argmax(x, weight=rf.decoder.weight, bias=rf.decoder.bias, ar_logits=ar, alpha=0.3)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"guided mix B={B} V={V}: {ms:.3f} ms  {(V * dim * 2 + B * V * 4) / ms / 1e6:.0f} GB/s of 6537")
