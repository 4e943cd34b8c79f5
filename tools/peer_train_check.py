"""2+ GPU check of the fused gradient exchange (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tools/peer_train_check.py

Trains the same model on the same per-rank data twice - gradients exchanged inside the optimizer kernel over
peer-mapped buffers, and with a plain NCCL all-reduce - and checks that (a) every rank ends with bit-identical
parameters, (b) the two exchanges agree, (c) how long a step takes either way.
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

from tinydiffusionmodels_b200.mnist import SimpleUNet
from tinydiffusionmodels_b200.unet_train import UNetTrainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, STEPS = 256, 8


def run(mode: str, use_graph: bool):
    os.environ["TDM_ALLREDUCE"] = mode
    torch.manual_seed(0)
    model = SimpleUNet().to(dev)
    tr = UNetTrainer(model, max_batch=B, seed=11, use_graph=use_graph)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    xs = [(torch.rand(B, 1, 28, 28, generator=g) * 2 - 1).to(dev) for _ in range(STEPS)]
    ts = [torch.randint(0, 1000, (B,), generator=g).to(dev) for _ in range(STEPS)]
    losses = [float(tr.step(x, t)) for x, t in zip(xs, ts)]
    torch.cuda.synchronize()
    # timing: 50 more steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    e0.record()
    for i in range(50):
        tr.step(xs[i % STEPS], ts[i % STEPS])
    e1.record()
    torch.cuda.synchronize()
    flat = model.flat_params().detach().clone()
    return flat, losses, e0.elapsed_time(e1) / 50


def exact_one_step() -> bool:
    """One optimizer step with the REAL cross-GPU exchange, exact: every rank writes a seeded gradient into its
    peer slot, runs tdm_adamw_flat_peer (flags + loads over NVLink), and must end bit-equal to tdm_adamw_flat on the
    rank-ordered sum of the all-gathered gradients - and, the replicas being what matters, bit-equal across ranks.
    Also reported: agreement with NCCL's all_reduce (bit-equal for 2 ranks, where the sum order cannot differ)."""
    from tinydiffusionmodels_b200 import _lib
    from tinydiffusionmodels_b200.unet_train import PeerGrads
    lib = _lib.load()
    n = 181_473
    peer = PeerGrads(lib, dev, n)
    okv = torch.tensor([1 if peer.ok else 0], device=dev, dtype=torch.int32)
    dist.all_reduce(okv, op=dist.ReduceOp.MIN)
    if int(okv) != 1:
        if rank == 0:
            print("exact one-step: peer mapping unavailable:", peer.error)
        return False
    good = True
    g0 = torch.Generator().manual_seed(1)
    p0 = torch.randn(n, generator=g0).to(dev)
    m0 = (torch.randn(n, generator=g0) * 1e-2).to(dev)
    v0 = (torch.rand(n, generator=g0) * 1e-3).to(dev)
    p, m, v = p0.clone(), m0.clone(), v0.clone()
    pr, mr, vr = p0.clone(), m0.clone(), v0.clone()
    step = torch.ones(1, dtype=torch.int64, device=dev)
    for k in (1, 2, 3):                                     # both gradient slots, the flags advancing
        gk = torch.Generator().manual_seed(1000 * k + rank)
        mine = (torch.randn(n, generator=gk) * (10.0 ** (rank % 3 - 1))).to(dev)
        peer.grad_view(k).copy_(mine)
        step.fill_(k)
        _lib.check(lib.tdm_adamw_flat_peer(p.data_ptr(), m.data_ptr(), v.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 0.01,
                                           1.0 / world, step.data_ptr(), peer.bases, world, rank, _lib.stream_ptr(dev)), "peer adamw")
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        gsum = torch.zeros(n, device=dev)
        for r in range(world):
            gsum = gsum + gathered[r]
        _lib.check(lib.tdm_adamw_flat(pr.data_ptr(), gsum.data_ptr(), mr.data_ptr(), vr.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8,
                                      0.01, 1.0 / world, step.data_ptr(), _lib.stream_ptr(dev)), "local adamw")
        torch.cuda.synchronize()
        same = torch.equal(p, pr) and torch.equal(m, mr) and torch.equal(v, vr)
        nc = mine.clone()
        dist.all_reduce(nc)
        nccl_same = torch.equal(nc, gsum)
        allp = [torch.empty_like(p) for _ in range(world)]
        dist.all_gather(allp, p)
        replicas = all(torch.equal(allp[0], a) for a in allp)
        flag = torch.tensor([1 if (same and replicas) else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"exact one-step k={k}: fused exchange == rank-ordered sum + AdamW (bit-equal on every rank): {bool(int(flag))}; "
                  f"replicas bit-identical: {replicas}; NCCL all_reduce sum bit-equal to the rank-ordered sum: {nccl_same}")
        good = good and bool(int(flag))
    dist.barrier()
    peer.close()
    return good


ok = exact_one_step()
for use_graph in (False, True):
    pf, pl, pms = run("peer", use_graph)
    nf, nl, nms = run("nccl", use_graph)
    gathered = [torch.empty_like(pf) for _ in range(world)]
    dist.all_gather(gathered, pf)
    same_ranks = all(torch.equal(gathered[0], g) for g in gathered)
    diff = float((pf - nf).abs().max())
    rel = float((pf - nf).norm() / nf.norm())
    if rank == 0:
        print(f"graph={use_graph}: ranks identical={same_ranks}  peer-vs-nccl max|d|={diff:.3e} rel={rel:.3e}  "
              f"losses {pl[0]:.4f}->{pl[-1]:.4f} (nccl {nl[0]:.4f}->{nl[-1]:.4f})  ms/step peer {pms:.3f} nccl {nms:.3f}")
    # the backward accumulates with float atomics, so two runs differ in the last bits and AdamW amplifies that over
    # 58 steps (measured rel 3e-4 .. 5e-3 from run to run); that the exchange itself is exact is what exact_one_step()
    # and the bit-identical ranks show
    ok = ok and same_ranks and rel < 2e-2 and all(x == x for x in pl)
if rank == 0:
    print("PEER_CHECK", "OK" if ok else "FAILED")
dist.destroy_process_group()
sys.exit(0 if ok else 1)
