"""Time the graph-replayed reverse loop in the SUSTAINED regime (what bench.py measures): the board reaches its power
cap after a few seconds of back-to-back trajectories and the SM clock settles near 1.5 GHz, where the balance between
the MMA issue thread, the epilogue warps and HBM differs from a cold 300-step probe (tools/loop_probe.py).

    python tools/hot_probe.py [batch=16384] [trajectories=3]     # us per reverse step of each full T=1000 trajectory
"""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from tinydiffusionmodels_b200.mnist import SimpleUNet, sample_loop

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = SimpleUNet().to(dev).eval()
x = torch.randn(B, 1, 28, 28, device=dev)
sample_loop(model, x.clone(), seed=1, steps=1000)   # warm-up trajectory: engine, graph, clocks down to the cap
torch.cuda.synchronize()
res = []
for i in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sample_loop(model, x.clone(), seed=1, steps=1000)
    e1.record()
    if i == n - 1:   # sample the clock while the last trajectory is in flight
        clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader", "-i", "0"],
                             capture_output=True, text=True).stdout.strip()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1))
print(f"B={B}  us/step per trajectory: " + " ".join(f"{r:.1f}" for r in res) + f"   [{clk}]")
