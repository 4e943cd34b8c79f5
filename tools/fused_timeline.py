"""In-kernel timeline of the fused rb4 block (development aid).

Build with  TDM_NVCC_DEFS=-DTDM_TIMELINE=104 (rb4) or 101 (rb1) python -m tinydiffusionmodels_b200.build --force , then
    python tools/fused_timeline.py [batch]
CTA 0 records clock64() per step: MMA thread (step top, acc2 free, acc1 free, input full, conv1 issued, t full,
conv2 issued), epilogue 1 (acc1 full seen, released, t written), epilogue 2 (acc2 full seen, done), gather
(stage empty, arrived).  Cycles relative to the step's own "top".
"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200 import _lib
from tinydiffusionmodels_b200.mnist import SimpleUNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = SimpleUNet().to(dev).eval()
eng = m.engine(B)
x = torch.randn(B, 1, 28, 28, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.int64)
for _ in range(3):
    eng.p_sample(x, t, None, out=x, seed=1)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * (96 * 16))()
lib.tdm_debug_read_timeline.argtypes = [ctypes.c_void_p]
lib.tdm_debug_read_timeline.restype = ctypes.c_int
assert lib.tdm_debug_read_timeline(buf) == 0
tl = [[buf[i * 16 + e] for e in range(16)] for i in range(96)]
names = ["top", "acc2free", "acc1free", "in.full", "c1.issued", "t.full", "c2.issued", "e1.seen", "e1.rel", "e1.done",
         "e2.seen", "e2.done", "g.empty", "g.arrive", "e1.bar", "e1.stored"]
print("step " + " ".join(f"{n:>10s}" for n in names) + "   (cycles since this step's top)")
for i in range(20, 44):
    ref = tl[i][0]
    print(f"{i:4d} " + " ".join(f"{tl[i][e] - ref:10d}" for e in range(16)))
d = [tl[i][0] - tl[i - 1][0] for i in range(10, 90)]
print("mean cycles per step (top to top):", sum(d) / len(d))
