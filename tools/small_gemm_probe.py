"""Fixed cost of one small tcgen05 GEMM launch: rows x 256 . (N x 256)^T through tdm_linear_logits (a rows_to_planes launch
plus the GEMM), back to back, warm.   python tools/small_gemm_probe.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.text_engine import Rounder

dev = torch.device("cuda:0")
r = Rounder(dev)
for rows, N in ((2048, 256), (2048, 768), (2048, 2048), (128, 256), (16384, 256)):
    w = torch.randn(N, 256, device=dev)
    b = torch.randn(N, device=dev)
    x = torch.randn(rows, 256, device=dev)
    for _ in range(5):
        r.logits(x, weight=w, bias=b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            r.logits(x, weight=w, bias=b)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"rows {rows} N {N}: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us per (pack + GEMM) pair, graph-replayed")
