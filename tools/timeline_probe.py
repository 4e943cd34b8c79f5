"""Print the in-kernel pipeline timeline of one convolution kernel (development aid).

Build the library with  TDM_NVCC_DEFS=-DTDM_TIMELINE=<EPI>  (EPI: 0 CONV1, 1 RES, 2 RES_X = rb1.conv2, 4 FINAL = rb4.conv2),
then   python tools/timeline_probe.py [batch]
CTA 0 records clock64() per tile: producer (empty ready), MMA warp (loop top, acc free, input full, MMAs issued),
epilogue (acc full seen, acc released, tile done).  Columns are cycles relative to the previous tile's "issued".
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
names = {8: "top", 2: "accfree", 3: "full", 9: "mma2", 10: "mma6", 11: "mma12", 12: "mma18", 13: "commits", 4: "issued", 0: "p.empty", 5: "e.accfull", 7: "e.release", 6: "e.done"}
ORDER = (3, 9, 10, 11, 12, 13, 0, 5, 7, 6)
print("tile " + " ".join(f"{names[e]:>10s}" for e in ORDER) + "   (cycles since the previous tile's commits)")
for i in range(8, 28):
    ref = tl[i - 1][13]
    print(f"{i:4d} " + " ".join(f"{tl[i][e] - ref:10d}" for e in ORDER))
d = [tl[i][13] - tl[i - 1][13] for i in range(8, 90)]
print("mean cycles per tile (commit to commit):", sum(d) / len(d))
