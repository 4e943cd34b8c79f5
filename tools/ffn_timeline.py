"""Development aid: per-chunk phase timeline (globaltimer, ns) of the fused FFN kernel, block 0, its second tile.
Needs the library built with TDM_NVCC_DEFS=-DTDM_EXP_TIMELINE.   python tools/ffn_timeline.py [sequences]"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200 import _lib
from tinydiffusionmodels_b200.shakespeare import TinyTransformer

B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = TinyTransformer(256).to(dev).eval()
eng = m.engine(B, 64)
x = torch.randn(B, 64, 256, device=dev)
t = torch.full((B,), 500, device=dev, dtype=torch.int64)
eng.load_state(x, t)
mode = sys.argv[2] if len(sys.argv) > 2 else "philox"      # philox | z (injected noise) | t0 (t = 0: no noise at all) | fwd (no reverse step)
z = torch.randn(B, 64, 256, device=dev) if mode == "z" else None
if mode == "t0":
    t.zero_()
for _ in range(3):
    if mode == "fwd":
        _lib.check(eng.lib.tdm_text_forward(eng.ptrs, eng.depth, eng.ws.data_ptr(), eng.ws_bytes, t.data_ptr(), B, 64, 256, eng._st()), "fwd")
    else:
        eng.p_sample_inplace(t, z, seed=1)
torch.cuda.synchronize()
lib = ctypes.CDLL(str(_lib.LIB_PATH))
buf = (ctypes.c_ulonglong * 144)()
lib.tdm_debug_ffn_timeline(buf)
v = [[buf[w * 16 + c] for c in range(16)] for w in range(8)]
t0 = min(x for row in v for x in row if x)
names = ["mma: wait P", "mma: P seen", "mma: G2 issued", "mma: G1(c+2) issued", "E1: acc1 seen", "E1: converted", "E1: P free", "E1: P written"]
print("chunk " + " ".join(f"{n:>20}" for n in names))
for c in range(16):
    print(f"{c:5d} " + " ".join(f"{(v[w][c] - t0) if v[w][c] else -1:20d}" for w in range(8)))
e = [buf[8 * 16 + i] for i in range(16)]
for g in (0, 1):
    print(f"tile epilogue, group {g}: E1 done {e[8*g]-t0}, acc2 complete {e[8*g+1]-t0}, pass 1 done {e[8*g+2]-t0}, stats exchanged {e[8*g+3]-t0}, pass 2 done {e[8*g+4]-t0}")
