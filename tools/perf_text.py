"""Throughput probe: elementwise kernels (GB/s), text reverse step, rounding (development aid)."""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200 import ops
from tinydiffusionmodels_b200.shakespeare import LearnedRounding, TinyTransformer
from tinydiffusionmodels_b200.text_engine import Rounder

dev = torch.device("cuda:0")


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


# ---- elementwise: 12 B/element (Philox) or 16 B/element (injected z) -------------------------
for B in (16384, 131072):
    x = torch.randn(B, 784, device=dev)
    e = torch.randn(B, 784, device=dev)
    z = torch.randn(B, 784, device=dev)
    t = torch.full((B,), 500, device=dev, dtype=torch.int64)
    out = torch.empty_like(x)
    n = x.numel()
    ms = timeit(lambda: ops.reverse_step(x, e, t, None, out=out, seed=1))
    print(f"reverse_step philox B={B}: {ms*1e3:.1f} us  {12*n/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: ops.reverse_step(x, e, t, z, out=out))
    print(f"reverse_step injected B={B}: {ms*1e3:.1f} us  {16*n/ms/1e6:.0f} GB/s")
    ms = timeit(lambda: ops.q_sample(x, t, z))
    print(f"q_sample B={B}: {ms*1e3:.1f} us  {12*n/ms/1e6:.0f} GB/s (incl. torch.empty_like)")
    del x, e, z, out

# ---- text reverse step ------------------------------------------------------------------------
for dim, batches in ((256, (5, 64, 512, 2048)), (2048, (5, 64))):
    torch.manual_seed(0)
    m = TinyTransformer(dim).to(dev).eval()
    for B in batches:
        eng = m.engine(B, 64)
        x = torch.randn(B, 64, dim, device=dev)
        t = torch.full((B,), 500, device=dev, dtype=torch.int64)
        eng.load_state(x, t)
        ms = timeit(lambda: eng.p_sample_inplace(t, None, seed=1), n=10, warm=3)
        flop = (8_060_928 if dim == 256 else 152_567_808) * 64 * B
        print(f"text p_sample dim={dim} B={B}: {ms*1e3:.1f} us/step  {B/(ms*1e-3)/1000:.1f} seq/s(T=1000)  {flop/ms/1e9:.1f} TF/s")

# ---- rounding ----------------------------------------------------------------------------------
V, dim = 256000, 256
torch.manual_seed(0)
rf = LearnedRounding(dim, V).to(dev)
r = Rounder(dev)
for rows in (320, 4096, 32768):
    x = torch.randn(rows, dim, device=dev)
    ms = timeit(lambda: r.argmax(x, weight=rf.decoder.weight, bias=rf.decoder.bias), n=5, warm=2)
    print(f"round_argmax V={V} rows={rows}: {ms:.3f} ms  {2*rows*dim*V/ms/1e9:.1f} TF/s  weights {V*dim*2/ms/1e6:.0f} GB/s")
ar = torch.randn(64, V, device=dev)
x = torch.randn(64, dim, device=dev)
ms = timeit(lambda: r.argmax(x, weight=rf.decoder.weight, bias=rf.decoder.bias, ar_logits=ar, alpha=0.3), n=5, warm=2)
print(f"guided mix B=64 V={V}: {ms:.3f} ms  (weights+ar bytes {(V*dim*2+64*V*4)/ms/1e6:.0f} GB/s)")
