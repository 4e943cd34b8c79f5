"""Development aid: phase timeline (globaltimer, ns) of block 0 of one small tcgen05 GEMM launch.
Needs the library built with TDM_NVCC_DEFS=-DTDM_EXP_TIMELINE.   python tools/gemm_timeline.py"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200 import _lib
from tinydiffusionmodels_b200.text_engine import Rounder

dev = torch.device("cuda:0")
r = Rounder(dev)
lib = ctypes.CDLL(str(_lib.LIB_PATH))
names = ["entry", "after alloc+sync", "after pdl_wait", "epilogue: chunk 2 loaded", "mma: first stage full",
         "mma: accumulator committed", "epilogue: accumulator seen", "epilogue warp done", "after final sync", "after dealloc",
         "epilogue: chunk 1 stores issued", "epilogue: chunk 0 loaded", "epilogue: chunk 0 released/arrive", "epilogue: chunk 1 loaded",
         "epilogue: chunk 1 before math", "epilogue: chunk 1 after bias + FMA, before stores"]
for rows, N in ((128, 256), (2048, 256)):
    w = torch.randn(N, 256, device=dev)
    b = torch.randn(N, device=dev)
    x = torch.randn(rows, 256, device=dev)
    for _ in range(3):
        r.logits(x, weight=w, bias=b)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    lib.tdm_debug_gemm_timeline(buf)
    t0 = buf[0]
    print(f"rows {rows} N {N}")
    for i in sorted(range(16), key=lambda i: buf[i]):
        print(f"   {buf[i] - t0:8d} ns  {names[i]}")
