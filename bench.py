#!/usr/bin/env python
"""bench.py — MNIST DDPM samples/sec (T=1000) on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload sample|train]

A *step* is one complete T=1000 reverse-sampling trajectory (src/mnist.py:190-194) of a batch of
`--batch` images per GPU: x_T -> 1000 x fused p_sample -> (clamp+1)/2.  Weak scaling: every rank
samples its own `--batch` images (global sample index = rank*batch + b keys the Philox noise, so
the union over ranks is independent of N), no data-path collective.

`value`  : samples/s with x_T already resident in HBM when the clock starts (device-timed).
`e2e`    : samples/s through the public host API: x_T copied from pinned host memory, the final
           [0,1] images copied back to pinned host memory, both inside the timed region.
`roofline`: the dominant kernel (rb4.conv1, 96->32 @28x28 + its 1x1 skip GEMM) — algorithmic
           FLOPs per launch / its CUDA-event duration, against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port of the reference's p_sample on the host cores (bounded sample).
`--impl reference`: the same oracle port as the reference arm (CPU, all host threads).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

T_STEPS = 1000
METRIC = "mnist_ddpm_samples_per_sec_T1000"
UNIT = "samples/s"
FLOPS_PER_IMAGE_STEP = 129_002_880


def _peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [c.strip() for c in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's p_sample on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reverse_steps(batch: int, n_reverse_steps: int, warm: int = 2) -> tuple[float, int]:
    """Seconds per reverse step of the CPU oracle (fp32, all host threads) at `batch`."""
    import torch
    from oracle import ddpm_oracle as O
    from tests.helpers import random_unet_state_dict

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = random_unet_state_dict(0)
    tab = O.make_tables()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 1, 28, 28, generator=g)
    with torch.no_grad():
        for i in range(warm):
            t = torch.full((batch,), T_STEPS - 1 - i, dtype=torch.long)
            x = O.mnist_p_sample(sd, x, t, torch.randn_like(x), tab)
        t0 = time.perf_counter()
        for i in range(n_reverse_steps):
            t = torch.full((batch,), T_STEPS - 1 - warm - i, dtype=torch.long)
            x = O.mnist_p_sample(sd, x, t, torch.randn_like(x), tab)
        dt = time.perf_counter() - t0
    return dt / n_reverse_steps, cores


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 64  # BASELINE.json configs[0]: the reference's own CPU-runnable case
    per_step = 10  # reverse steps timed per bench "step" (bounded sample of the 1000)
    for _ in range(args.warmup):
        cpu_reverse_steps(batch, 2, warm=1)
    t0 = time.perf_counter()
    secs = []
    cores = 1
    for _ in range(args.steps):
        s, cores = cpu_reverse_steps(batch, per_step, warm=1)
        secs.append(s)
    wall = time.perf_counter() - t0
    sec_per_rstep = sum(secs) / len(secs)
    value = batch / (sec_per_rstep * T_STEPS)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"MNIST DDPM UNet 28x28x1, T={T_STEPS} reverse sampling, {args.batch} samples per GPU "
                               f"({args.gpus * args.batch} total), random-init weights",
                   "samples_per_gpu": args.batch, "T": T_STEPS,
                   "note": "reference arm: CPU fp32 oracle port of src/mnist.py:167-194 on a bounded sample of this workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} of {T_STEPS} reverse steps at batch {batch} per bench step, x1000/{per_step} extrapolated"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path for the product arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from tinydiffusionmodels_b200 import _lib, ops
    from tinydiffusionmodels_b200.mnist import SimpleUNet, sample_loop
    from tinydiffusionmodels_b200.unet_engine import UNetEngine

    torch.manual_seed(0)
    model = SimpleUNet().to(dev).eval()       # random-init weights of the reference architecture
    B = args.batch
    seed = 20260101
    offset = rank * B                          # global sample index of this rank's first image
    eng = model.engine(B)
    lib = _lib.load()

    # ---- one captured reverse step, replayed T times per trajectory -------------------------
    x = torch.empty(B, 1, 28, 28, device=dev)
    t_buf = torch.empty(B, dtype=torch.int64, device=dev)
    out01 = torch.empty_like(x)
    host_in = torch.randn(B, 1, 28, 28).pin_memory()
    host_out = torch.empty(B, 1, 28, 28).pin_memory()

    def one_step():
        eng.p_sample(x, t_buf, None, out=x, seed=seed, sample_offset=offset)
        _lib.check(lib.tdm_timestep_advance(t_buf.data_ptr(), B, -1, _lib.stream_ptr(dev)), "advance")

    t_buf.fill_(T_STEPS - 1)
    x.normal_()
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        one_step()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one_step()
    launches_per_replay = _lib.launch_count() - n0

    def trajectory_device():
        """x_T drawn on the device (Philox), T reverse steps, map to [0,1]."""
        _lib.check(lib.tdm_randn_philox(x.data_ptr(), B, 784, seed, offset, 0, _lib.stream_ptr(dev)), "randn")
        t_buf.fill_(T_STEPS - 1)
        for _ in range(T_STEPS):
            graph.replay()
        _lib.check(lib.tdm_to_unit_range(x.data_ptr(), out01.data_ptr(), x.numel(), _lib.stream_ptr(dev)), "unit")

    x_e2e = torch.empty_like(x)

    def trajectory_e2e():
        """The call a user makes (tinydiffusionmodels_b200.mnist.sample_loop, the loop of src/mnist.py:190-194) on
        host data: pinned x_T -> device, T reverse steps through the public API (it keeps the captured step per
        (model, batch, seed) and replays it), clamp + [0,1] map, images -> pinned host memory."""
        x_e2e.copy_(host_in, non_blocking=True)
        sample_loop(model, x_e2e, seed=seed, sample_offset=offset)
        host_out.copy_(ops.to_unit_range(x_e2e), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps: int) -> float:
        """Max-over-ranks device milliseconds for `steps` calls of fn."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        trajectory_device()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n_before = _lib.launch_count()
    ms = timed(trajectory_device, args.steps)
    direct_launches = _lib.launch_count() - n_before
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    trajectory_e2e()  # warm the pinned-copy path
    ms_e2e = timed(trajectory_e2e, max(1, min(args.steps, 3)))
    e2e_steps = max(1, min(args.steps, 3))
    e2e_value = world * B * e2e_steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: live CUDA events between the nine launches --------
    per_kernel = []
    t_buf.fill_(500)
    for i in range(8):
        per_kernel.append(eng.profile_kernels(x, t_buf, seed=seed))
    per_kernel = per_kernel[2:]
    names = [k[0] for k in per_kernel[0]]
    flops = [k[1] for k in per_kernel[0]]
    nk = len(names)
    kms = [statistics.mean(k[i][2] for k in per_kernel) for i in range(nk)]
    top = max(range(nk), key=lambda i: kms[i])
    peaks, peak_src = _peaks()
    peak_tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops")))
    achieved_tf = flops[top] * B / (kms[top] * 1e-3) / 1e12
    step_ms_sum = sum(kms)
    # DRAM traffic from the committed ncu captures of this build (profiles/r02_traffic_step.json: per-kernel
    # dram__bytes_read.sum + dram__bytes_write.sum of one reverse step at this batch)
    traffic = step_traffic = step_hbm_frac = None
    sfile = ROOT / "profiles" / "r02_traffic_step.json"
    if sfile.exists():
        sj = json.loads(sfile.read_text())
        if sj.get("batch") == B and sj.get("fused") == eng.fused():
            step_traffic = sj["dram_bytes_per_step"]
            step_hbm_frac = step_traffic / (step_ms_sum * 1e-3) / (float(peaks["hbm_gbs"]) * 1e9)
            traffic = sj.get("per_kernel", {}).get(names[top])
    roofline = {
        "bound": "tensor", "kernel": names[top], "achieved": achieved_tf, "peak": peak_tf,
        "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
        "peak_source": f"{peak_src} bf16_tflops_sustained (kernel timed inside a long step)",
        "kernel_ms": {n: round(v, 4) for n, v in zip(names, kms)},
        "kernel_share_of_step": round(kms[top] / step_ms_sum, 4),
        "kernel_tflops": {n: round(f * B / (v * 1e-3) / 1e12, 1) for n, f, v in zip(names, flops, kms) if f},
        "fused_blocks": eng.fused(),
        "whole_unet_tflops": FLOPS_PER_IMAGE_STEP * B / (step_ms_sum * 1e-3) / 1e12,
        "step_dram_bytes": step_traffic, "step_hbm_frac": step_hbm_frac,
    }

    peaks_all = {"hbm_gbs": float(peaks["hbm_gbs"]), "tf_sustained": peak_tf, "source": peak_src}
    # ---- BASELINE.json configs[2]: the batch sweep (and configs[0]'s batch 64), bounded to a few reverse steps ----
    sweep = None if args.no_extras else bench_sweep(model, dev, seed, offset, world)
    # ---- the reference's own eager PyTorch path on this B200 ----
    gpu_eager = None if (args.no_extras or rank != 0) else bench_gpu_eager(dev, (64, B))
    # ---- secondary metric: UNet train images/s (BASELINE.json configs[1]), data-parallel ------
    train = bench_train(dev, rank, world, args.train_batch, barrier, peaks_all)
    # BASELINE.json configs[1] as written: global batch 512 over 8 GPUs = 64 images per GPU
    train64 = None if args.no_extras else bench_train(dev, rank, world, 64, barrier, peaks_all, pipeline=False)
    # ---- secondary metric: Shakespeare sampler sequences/s (BASELINE.json configs[3], [4]) -----
    text = bench_text(dev, rank, world, args.text_batch, barrier, peaks_all) if not args.no_text else None
    text2048 = (bench_text(dev, rank, world, max(64, args.text_batch // 4), barrier, peaks_all, dim=2048)
                if not (args.no_text or args.no_extras) else None)

    # ---- row f2: the Shakespeare training step (src/shakespeare.py:221-250) at the reference CLI's batch 32 x 64 ----
    text_train = None if (args.no_text or args.no_extras) else bench_text_train(dev, rank, world, 32, barrier, peaks_all)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on this box's host cores (bounded sample; rank 0 at N=1 only) ---------
    cpu_baseline = None
    if world == 1:
        sec_per_rstep, cores = cpu_reverse_steps(64, 20, warm=3)
        cpu_baseline = {"value": 64 / (sec_per_rstep * T_STEPS), "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "20 of 1000 reverse steps at batch 64 (oracle port of src/mnist.py:167-180), extrapolated x50"}
        if not args.no_extras:
            train["cpu_baseline"] = cpu_train_baseline(cores)
            if train64 is not None:
                train64["cpu_baseline"] = train["cpu_baseline"]
            if text is not None:
                text["cpu_baseline"] = cpu_text_baseline(cores, 256)
            if text2048 is not None:
                text2048["cpu_baseline"] = cpu_text_baseline(cores, 2048)
            if text_train is not None:
                text_train["cpu_baseline"] = cpu_text_train_baseline(cores)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"MNIST DDPM UNet 28x28x1, T={T_STEPS} reverse sampling, {B} samples per GPU "
                        f"({world * B} total), random-init weights, Philox noise in-kernel",
            "samples_per_gpu": B, "T": T_STEPS,
            "l2": f"working set {eng.ws_bytes / 2**20:.0f} MiB of activations per step > 126 MB L2",
            "accumulate": "fp32", "activations": "bf16",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * 784 * 4,
                "d2h_bytes_per_step": B * 784 * 4, "steps": e2e_steps},
        "gpu_launches": int(direct_launches + launches_per_replay * T_STEPS * args.steps),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "clocks": clocks,
        "train": train,
        "train_global512_per8": train64,
        "text": text,
        "text_dim2048": text2048,
        "text_train": text_train,
        "sweep": sweep,
        "gpu_eager_baseline": gpu_eager,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _event_ms(torch, fn, n: int) -> float:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def bench_sweep(model, dev, seed, offset, world, sizes=(64, 1024, 4096, 16384, 65536, 262144)) -> dict:
    """BASELINE.json configs[0] (batch 64) and configs[2] (1K-256K samples per GPU): graph-replayed reverse steps
    through the same engine, a BOUNDED number of steps per size (the full T=1000 trajectory at 256K samples alone
    would take a minute); samples/s = samples / (1000 x measured ms per reverse step).  Sizes above the engine's
    chunk (mnist.SAMPLE_CHUNK) run as sequential chunks on one workspace, exactly as sample_loop does."""
    import torch

    from tinydiffusionmodels_b200 import _lib
    from tinydiffusionmodels_b200.mnist import SAMPLE_CHUNK

    lib = _lib.load()
    out = {}
    for n in sizes:
        chunk = min(n, SAMPLE_CHUNK)
        nchunks = (n + chunk - 1) // chunk
        eng = model.engine(chunk)
        eng._prep(chunk)
        x = torch.randn(chunk, 1, 28, 28, device=dev)
        t = torch.full((chunk,), T_STEPS - 1, dtype=torch.int64, device=dev)

        def one():
            eng.p_sample(x, t, None, out=x, seed=seed, sample_offset=offset)
            _lib.check(lib.tdm_timestep_advance(t.data_ptr(), chunk, -1, _lib.stream_ptr(dev)), "advance")

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            one()
        torch.cuda.current_stream(dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            one()
        reps = 200 if n <= 1024 else 40 if n <= 16384 else 12
        for _ in range(5):
            g.replay()
        t.fill_(T_STEPS - 1)
        ms = _event_ms(torch, g.replay, reps) * nchunks
        out[str(n)] = {"samples_per_s": world * n / (ms * 1e-3 * T_STEPS), "ms_per_reverse_step": ms,
                       "reverse_steps_timed": reps, "chunks": nchunks}
        del g, x, t
    out["note"] = ("per GPU; x world GPUs (weak scaling, no collective); bounded sample of the T=1000 trajectory: "
                   "graph-replayed reverse steps, extrapolated x1000")
    return out


def bench_gpu_eager(dev, batches) -> dict:
    """The reference's own path on this B200 (src/mnist.py:224-233 picks `cuda`): the oracle port of p_sample - the
    same ATen/cuDNN ops the reference dispatches - run eagerly on the GPU, in fp32 (TF32 off), with TF32 convolutions,
    and under bf16 autocast.  A baseline leg like cpu_baseline: never the product path."""
    import torch
    from oracle import ddpm_oracle as O
    from tests.helpers import random_unet_state_dict

    sd = {k: v.to(dev) for k, v in random_unet_state_dict(0).items()}
    tab = {k: (v.to(dev) if hasattr(v, "to") else v) for k, v in O.make_tables().items()}
    res = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for b in batches:
            x = torch.randn(b, 1, 28, 28, device=dev)
            z = torch.randn_like(x)
            t = torch.full((b,), 500, dtype=torch.long, device=dev)
            row = {}
            for mode in ("fp32", "tf32", "bf16_autocast"):
                torch.backends.cudnn.allow_tf32 = mode != "fp32"
                torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"

                def step():
                    with torch.no_grad():
                        if mode == "bf16_autocast":
                            with torch.autocast("cuda", dtype=torch.bfloat16):
                                eps = O.unet_forward(sd, x, t)
                            return O.reverse_step(x, eps.float(), t, z, tab)
                        return O.mnist_p_sample(sd, x, t, z, tab)

                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                ms = _event_ms(torch, step, 10 if b > 1024 else 50)
                row[mode] = {"ms_per_reverse_step": ms, "samples_per_s": b / (ms * 1e-3 * T_STEPS)}
            res[str(b)] = row
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    res["note"] = ("PyTorch eager (cuDNN/ATen) on the same B200, oracle port of src/mnist.py:167-180 including the "
                   "reference's per-step host sync on `t[0] == 0`; 10-50 reverse steps timed with CUDA events, extrapolated x1000")
    return res


def cpu_train_baseline(cores: int) -> dict:
    """Reference inner training step (src/mnist.py:153-159) on the host cores: oracle loss + autograd + AdamW, B=512."""
    import torch
    from oracle import ddpm_oracle as O
    from tests.helpers import random_unet_state_dict

    torch.set_num_threads(cores)
    sd = random_unet_state_dict(0)
    tab = O.make_tables()
    g = torch.Generator().manual_seed(0)
    b = 512
    x0 = torch.rand(b, 1, 28, 28, generator=g) * 2 - 1
    flat = {k: (v.clone(), torch.zeros_like(v), torch.zeros_like(v)) for k, v in sd.items()}

    def step(k):
        t = torch.randint(0, T_STEPS, (b,), generator=g)
        noise = torch.randn(b, 1, 28, 28, generator=g)
        cur = {n: p for n, (p, _, _) in flat.items()}
        _, grads = O.mnist_loss_and_grads(cur, x0, t, noise, tab)
        for n, (p, m, v) in flat.items():
            flat[n] = O.adamw_step(p, grads[n], m, v, k)

    step(1)
    t0 = time.perf_counter()
    n = 3
    for k in range(n):
        step(k + 2)
    dt = (time.perf_counter() - t0) / n
    return {"value": b / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{n} optimizer steps at batch {b} after 1 warm-up (oracle port of src/mnist.py:153-159: q_sample, UNet, MSE, backward, AdamW)"}


def cpu_text_baseline(cores: int, dim: int) -> dict:
    """Reference text reverse step (src/shakespeare.py:343-352) on the host cores, n=5, L=64 (the reference's sampling
    job), bounded to a few of the 1000 steps."""
    import torch
    from oracle import ddpm_oracle as O
    from tinydiffusionmodels_b200.shakespeare import TinyTransformer

    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = {k: v.detach().clone() for k, v in TinyTransformer(dim).state_dict().items()}
    tab = O.make_tables()
    n, L = 5, 64
    x = torch.randn(n, L, dim)
    nsteps = 20 if dim == 256 else 5
    with torch.no_grad():
        for i in range(2):
            x = O.text_p_sample(sd, x, torch.full((n,), 999 - i, dtype=torch.long), torch.randn_like(x), tab)
        t0 = time.perf_counter()
        for i in range(nsteps):
            x = O.text_p_sample(sd, x, torch.full((n,), 997 - i, dtype=torch.long), torch.randn_like(x), tab)
        dt = (time.perf_counter() - t0) / nsteps
    return {"value": n / (dt * T_STEPS), "unit": "sequences/s", "cores": cores, "kind": "port",
            "sample": f"{nsteps} of 1000 reverse steps at n={n}, L={L}, dim={dim} (oracle port of src/shakespeare.py:343-352), "
                      f"extrapolated; rounding not included"}


def cpu_text_train_baseline(cores: int, batch: int = 8, vocab: int = 256_000, dim: int = 256) -> dict:
    """One reference training step (src/shakespeare.py:221-250: forward in train mode, both losses, backward, AdamW over
    all three modules) on the host cores through the oracle port, at a bounded batch of 8 x 64 tokens (the (tokens, V)
    fp32 logits and their gradient are 0.5 GB each at this size; the reference CLI's 32 x 64 needs 4 x that)."""
    import torch
    from oracle import ddpm_oracle as O
    from oracle import text_train_oracle as TO
    from tinydiffusionmodels_b200.shakespeare import TinyTransformer

    torch.set_num_threads(cores)
    torch.manual_seed(0)
    L = 64
    sd = {k: v.detach().clone() for k, v in TinyTransformer(dim).state_dict().items()}
    dec_w, dec_b, emb = torch.randn(vocab, dim) * 0.05, torch.zeros(vocab), torch.randn(vocab, dim) * 0.5
    tab = O.make_tables()
    state = {k: (torch.zeros_like(v), torch.zeros_like(v)) for k, v in (("w", dec_w), ("b", dec_b), ("e", emb))}

    def step(k):
        nonlocal dec_w, dec_b, emb
        ids = torch.randint(0, vocab, (batch, L))
        t = torch.randint(0, 1000, (batch,))
        noise = torch.randn(batch, L, dim)
        _, g = TO.text_losses_and_grads(sd, dec_w, dec_b, emb, ids, t, noise, tab, dropout=0.1, seed=1, step=k)
        dec_w, *mv = O.adamw_step(dec_w, g["decoder.weight"], *state["w"], k, lr=1e-4, wd=1e-4); state["w"] = tuple(mv)
        dec_b, *mv = O.adamw_step(dec_b, g["decoder.bias"], *state["b"], k, lr=1e-4, wd=1e-4); state["b"] = tuple(mv)
        emb, *mv = O.adamw_step(emb, g["embeddings.weight"], *state["e"], k, lr=1e-4, wd=1e-4); state["e"] = tuple(mv)

    step(1)
    t0 = time.perf_counter()
    n = 2
    for k in range(n):
        step(2 + k)
    dt = (time.perf_counter() - t0) / n
    return {"value": batch / dt, "unit": "sequences/s", "cores": cores, "kind": "port",
            "sample": f"{n} training steps at batch {batch} x {L} tokens, V={vocab}, dim={dim}, fp32 (oracle port of "
                      f"src/shakespeare.py:221-250 incl. AdamW on the embedding / decoder matrices; the 3.2 M encoder "
                      f"parameters' update is not included); {dt:.2f} s per step"}


def bench_text_train(dev, rank, world, batch, barrier, peaks, steps: int = 20, warmup: int = 4, vocab: int = 256_000,
                     dim: int = 256) -> dict:
    """Row f2: one optimisation step of the Shakespeare model (src/shakespeare.py:221-250) at the reference CLI's batch
    (32 sequences x 64 tokens per GPU), V = 256,000, width 256, depth 3, dropout 0.1, learned embeddings: forward,
    both losses, backward, [NCCL all-reduce of the 131.6 M-float gradient when data parallel], AdamW, weight re-pack -
    one replayed CUDA graph on one GPU.  Token ids come from pinned host memory every step (the e2e figure)."""
    import torch
    import torch.distributed as dist

    from tinydiffusionmodels_b200 import _lib
    from tinydiffusionmodels_b200.shakespeare import LearnedEmbedding, LearnedRounding, TinyTransformer
    from tinydiffusionmodels_b200.text_train import TextTrainer

    torch.manual_seed(0)
    L = 64
    m, r, e = TinyTransformer(dim).to(dev), LearnedRounding(dim, vocab).to(dev), LearnedEmbedding(vocab, dim).to(dev)
    tr = TextTrainer(m, r, e, dev, batch, L, lr=1e-4, weight_decay=1e-4, seed=11)
    ids_dev = torch.randint(0, vocab, (batch, L), device=dev)
    ids_host = ids_dev.cpu().pin_memory()
    tr.step(ids_dev)
    launches = tr.launches_per_step
    for _ in range(warmup):
        tr.step(ids_dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.step(ids_dev)
    e1.record()
    torch.cuda.synchronize()
    ms_t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    barrier()
    e0.record()
    for k in range(steps):
        loss = tr.step(ids_host, lr=1e-4, rounding_weight=1.0)   # H2D of the ids + scalar updates inside the timed region
        if k % 10 == 9:
            loss.tolist()                                        # the reference reads its losses (.item()) for logging
    final = tr.losses.tolist()
    e1.record()
    torch.cuda.synchronize()
    ms_e = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_e, op=dist.ReduceOp.MAX)
    ms, mse = float(ms_t.item()), float(ms_e.item())
    # the step's HBM-bound part on its own: AdamW over every parameter (28 B per parameter: p, g, m, v read; p, m, v written)
    e0.record()
    for _ in range(5):
        _lib.check(tr.lib.tdm_adamw_flat_lr(tr.flat.data_ptr(), tr.grads.data_ptr(), tr.exp_avg.data_ptr(), tr.exp_avg_sq.data_ptr(),
                                            tr.n, tr.lr_dev.data_ptr(), 0.9, 0.999, 1e-8, 1e-4, 1.0, tr.step_dev.data_ptr(),
                                            _lib.stream_ptr(dev)), "tdm_adamw_flat_lr")
    e1.record()
    torch.cuda.synchronize()
    ms_adam = e0.elapsed_time(e1) / 5
    tokens = batch * L
    # algorithmic FLOP per token: forward + dX + dW of the encoder (3 x 8,060,928) and of the rounding
    # head (3 x 2*dim*V); the recomputation of the logits in the gradient pass is NOT counted
    flop_tok = 3 * (8_060_928 + 2.0 * dim * vocab)
    tf = flop_tok * tokens / (ms * 1e-3) / 1e12
    adam_bytes = 28.0 * tr.n
    out = {
        "metric": "shakespeare_train_sequences_per_sec", "unit": "sequences/s", "value": world * batch / (ms * 1e-3),
        "e2e": {"value": world * batch / (mse * 1e-3), "unit": "sequences/s", "h2d_bytes_per_step": batch * L * 8 + 8,
                "d2h_bytes_per_step": 12 / 10, "what": "TextTrainer.step from pinned host token ids, learning rate and "
                "rounding weight rewritten every step, losses read back every 10th step"},
        "ms_per_step": ms, "gpu_launches_per_step": int(launches), "losses_after": final,
        "config": {"workload": f"Shakespeare training step, {batch} x {L} tokens per GPU, V={vocab}, TinyTransformer({dim}), "
                               f"dropout 0.1, learned embeddings + learned rounding, AdamW over {tr.n:,} parameters; "
                               f"synthetic token ids, random-init weights", "parallelism": f"dp{world}"},
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": tf / peaks["tf_sustained"], "traffic": None,
                     "what": "whole step: 3 x (encoder 8.06 MFLOP + rounding head 2*dim*V) FLOP per token x tokens / step time; "
                             "the step also streams 28 B per parameter through AdamW "
                             f"({adam_bytes / 1e9:.2f} GB = {adam_bytes / peaks['hbm_gbs'] / 1e6:.2f} ms at the HBM copy peak)"},
        "dtype": "bf16 operands, fp32 accumulation / activations / master weights / optimiser state",
    }
    out["roofline"]["adamw"] = {"bound": "hbm", "achieved": adam_bytes / (ms_adam * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                "frac": adam_bytes / (ms_adam * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms": ms_adam,
                                "what": f"AdamW over {tr.n:,} parameters timed alone, 28 B per parameter"}
    del tr, m, r, e
    torch.cuda.empty_cache()
    return out


def bench_train(dev, rank, world, batch, barrier, peaks, steps: int = 30, warmup: int = 5, pipeline: bool = True) -> dict:
    """UNet DDPM training step (src/mnist.py:153-159) on synthetic U(-1,1) images, `batch` per GPU,
    gradients summed inside the optimizer kernel over peer-mapped buffers when world > 1 (NCCL all-reduce if the
    ranks cannot map each other).  Device-resident and host-fed variants."""
    import torch
    import torch.distributed as dist

    from tinydiffusionmodels_b200 import _lib
    from tinydiffusionmodels_b200.mnist import SimpleUNet
    from tinydiffusionmodels_b200.unet_train import UNetTrainer

    torch.manual_seed(0)
    model = SimpleUNet().to(dev)
    trainer = UNetTrainer(model, lr=1e-3, max_batch=batch, seed=7)
    g = torch.Generator(device=dev).manual_seed(rank)
    x_dev = torch.rand(batch, 1, 28, 28, device=dev, generator=g) * 2 - 1
    x_host = (torch.rand(batch, 1, 28, 28) * 2 - 1).pin_memory()
    loss_host = torch.zeros(1).pin_memory()

    def run(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_dev():
        trainer.step(x_dev)

    def step_e2e():
        loss = trainer.step(x_host.to(dev, non_blocking=True))
        loss_host.copy_(loss, non_blocking=True)

    for _ in range(warmup):
        step_dev()
    n0 = _lib.launch_count()
    ms = run(step_dev, steps)
    step_e2e()
    ms_e2e = run(step_e2e, steps)
    # launches per step are counted from one eager (non-graph) step of the same trainer
    trainer.use_graph = False
    n0 = _lib.launch_count()
    trainer.step(x_dev)
    launches = _lib.launch_count() - n0
    flops = 386_506_496 * batch   # fwd+bwd per image
    # device-resident input pipeline: one epoch of MNIST-sized uint8 data, shuffled and
    # normalised on the device in batches of `batch` (1 B read + 4 B written per pixel)
    from tinydiffusionmodels_b200.data import DeviceImages

    input_pipeline = None
    if pipeline:
        ds = DeviceImages(torch.randint(0, 256, (60000, 28, 28), dtype=torch.uint8), dev)
        for _ in ds.batches(batch, seed=1, epoch=0, max_batches=8, rank=0, world=1):
            pass
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in ds.batches(batch, seed=1, epoch=1, rank=0, world=1):
            pass
        e1.record()
        torch.cuda.synchronize()
        ms_epoch = e0.elapsed_time(e1)
        input_pipeline = {"images_per_s": 60000 / (ms_epoch * 1e-3), "gb_per_s": 60000 * 784 * 5 / (ms_epoch * 1e-3) / 1e9,
                          "what": f"one 60,000-image epoch: device randperm + gather/normalise kernel per batch of {batch}"}
    tf = flops / (ms / steps * 1e-3) / 1e12
    return {
        "metric": "unet_train_images_per_sec", "unit": "images/s",
        "input_pipeline": input_pipeline,
        "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": tf / peaks["tf_sustained"], "traffic": None,
                     "what": "whole training step (34 launches): 386,506,496 FLOP/image fwd+bwd (BASELINE.md section 3) / "
                             "step time, against the sustained bf16 peak"},
        "value": world * batch * steps / (ms * 1e-3),
        "e2e": {"value": world * batch * steps / (ms_e2e * 1e-3), "unit": "images/s",
                "h2d_bytes_per_step": batch * 784 * 4, "d2h_bytes_per_step": 4},
        "ms_per_step": ms / steps, "steps": steps, "batch_per_gpu": batch, "global_batch": world * batch,
        "dtype": "bf16 (fp32 master weights, AdamW moments and accumulation)",
        "tflops_per_gpu": flops / (ms / steps * 1e-3) / 1e12,
        "gpu_launches_per_step": int(launches),
        "allreduce_bytes_per_step": 181_473 * 4 if world > 1 else 0,
        "gradient_exchange": ("none" if world == 1 else "fused into AdamW over peer-mapped buffers"
                              if trainer.peer is not None else "NCCL all-reduce"),
        "final_loss": float(trainer.loss.item()),
    }


def bench_text(dev, rank, world, batch, barrier, peaks, rsteps: int = 100, dim: int = 256, vocab: int = 256_000) -> dict:
    """Shakespeare embedding-space sampler (src/shakespeare.py:355-426): TinyTransformer(256), L=64,
    `batch` sequences per GPU.  Times `rsteps` graph-replayed reverse steps (of the T=1000 loop) and
    the learned-rounding argmax at V=256,000; sequences/s = batch / (1000*step + rounding)."""
    import torch
    import torch.distributed as dist

    from tinydiffusionmodels_b200 import _lib
    from tinydiffusionmodels_b200.shakespeare import LearnedRounding, TinyTransformer
    from tinydiffusionmodels_b200.text_engine import Rounder

    torch.manual_seed(0)
    L, V = 64, vocab
    if dim != 256:
        rsteps = 30
    model = TinyTransformer(dim).to(dev).eval()
    eng = model.engine(batch, L)
    x = torch.randn(batch, L, dim, device=dev)
    t = torch.full((batch,), 999, device=dev, dtype=torch.int64)
    eng.load_state(x, t)

    def one():
        eng.p_sample_inplace(t, None, seed=3, sample_offset=rank * batch)
        _lib.check(eng.lib.tdm_timestep_advance(t.data_ptr(), batch, -1, _lib.stream_ptr(dev)), "advance")

    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        one()
    torch.cuda.current_stream(dev).wait_stream(side)
    n0 = _lib.launch_count()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        one()
    launches = _lib.launch_count() - n0
    t.fill_(999)
    for _ in range(5):
        g.replay()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rsteps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / rsteps
    rf = LearnedRounding(dim, V).to(dev)
    rounder = Rounder(dev)
    z = eng.read(0)
    rounder.argmax(z, weight=rf.decoder.weight, bias=rf.decoder.bias)
    e0.record()
    for _ in range(3):
        rounder.argmax(z, weight=rf.decoder.weight, bias=rf.decoder.bias)
    e1.record()
    torch.cuda.synchronize()
    ms_round = e0.elapsed_time(e1) / 3
    # one guided-mix position: a whole number of 256-row pairs of tiles (the GEMM's CTAs run as pairs over two row tiles)
    gm_batch = batch // 256 * 256 if batch >= 256 else batch
    ar = torch.randn(gm_batch, V, device=dev)
    zg0 = z[:gm_batch, 0].contiguous()
    rounder.argmax(zg0, weight=rf.decoder.weight, bias=rf.decoder.bias, ar_logits=ar, alpha=0.3)
    e0.record()
    for _ in range(3):
        rounder.argmax(zg0, weight=rf.decoder.weight, bias=rf.decoder.bias, ar_logits=ar, alpha=0.3)
    e1.record()
    torch.cuda.synchronize()
    ms_mix = e0.elapsed_time(e1) / 3
    tot = torch.tensor([1000 * ms_step + ms_round], device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    guided = None
    if dim == 256:
        # BASELINE.json configs[4] end to end: guided_generate (src/shakespeare.py:429-470) with a random-init 4-layer
        # Gemma standing in for google/gemma-2b-it (no hub access), stepped through its KV cache vs the reference's
        # full-prefix re-forward; 64 sequences x 64 positions, V = 256,000
        try:
            from transformers import GemmaConfig, GemmaForCausalLM

            from tinydiffusionmodels_b200.shakespeare import LearnedEmbedding, _SyntheticTokenizer, guided_generate

            gb = 64
            cfg = GemmaConfig(vocab_size=V, hidden_size=256, intermediate_size=1024, num_hidden_layers=4, num_attention_heads=4,
                              num_key_value_heads=1, head_dim=64, max_position_embeddings=256)
            lm = GemmaForCausalLM(cfg).to(dev).eval()
            zg = z[:gb].contiguous()
            embf = LearnedEmbedding(8, dim)   # unused with learned rounding
            res = {}
            for mode, kv in (("kv_cached", True), ("full_prefix", False)):
                guided_generate(lm, rf, _SyntheticTokenizer(), embf, zg[:, :4].contiguous(), alpha=0.3, use_kv_cache=kv)   # warm-up
                torch.cuda.synchronize()
                e0.record()
                guided_generate(lm, rf, _SyntheticTokenizer(), embf, zg, alpha=0.3, use_kv_cache=kv)
                e1.record()
                torch.cuda.synchronize()
                res[mode] = {"ms": e0.elapsed_time(e1), "sequences_per_s": world * gb / (e0.elapsed_time(e1) * 1e-3)}
            guided = {"what": f"guided_generate end to end, {gb} sequences x {L} positions, V={V}, random-init 4-layer Gemma "
                              "stand-in (hidden 256) as the base LM; includes the LM forward and the host-side decode", **res}
            del lm
        except Exception as e:   # noqa: BLE001  (transformers missing or incompatible: the leg is optional)
            guided = {"unavailable": str(e)[:200]}
    flop_tok = 8_060_928 if dim == 256 else 3 * (8 * dim * dim + 4 * L * dim + 8 * 2048 * dim)
    den_tf = flop_tok * L * batch / (ms_step * 1e-3) / 1e12
    rnd_tf = 2.0 * dim * V * L * batch / (ms_round * 1e-3) / 1e12
    mix_bytes = gm_batch * V * 4 + (dim * V * 2)              # fp32 AR logits once + the bf16 rounding matrix once
    mix_gbs = mix_bytes / (ms_mix * 1e-3) / 1e9
    return {
        "roofline": {"bound": "tensor", "achieved": den_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": den_tf / peaks["tf_sustained"], "traffic": None,
                     "what": "denoiser reverse step: algorithmic FLOP/token-step (BASELINE.md section 3) x tokens / step time",
                     "rounding": {"bound": "tensor", "achieved": rnd_tf, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                                  "frac": rnd_tf / peaks["tf_sustained"], "what": f"2*dim*V FLOP/token, {L * batch} tokens, V={V}"},
                     "guided_mix": {"bound": "hbm", "achieved": mix_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                    "frac": mix_gbs / peaks["hbm_gbs"],
                                    "what": f"one position of guided_generate at {gm_batch} sequences: fp32 AR logits ({gm_batch}xV) + bf16 W (VxD) read once"}},
        "metric": "shakespeare_sequences_per_sec_T1000", "unit": "sequences/s",
        "value": world * batch / (float(tot.item()) * 1e-3),
        "config": {"workload": f"TinyTransformer(dim={dim}, depth 3, 4 heads), L=64, {batch} sequences per GPU, "
                               f"T=1000 reverse steps + learned rounding argmax over V={V}; random-init weights"},
        "ms_per_reverse_step": ms_step, "reverse_steps_timed": rsteps, "ms_rounding": ms_round,
        "denoiser_tflops": flop_tok * L * batch / (ms_step * 1e-3) / 1e12,
        "guided_mix_ms_per_position": ms_mix, "guided_generate_e2e": guided, "gpu_launches_per_reverse_step": int(launches),
        "dtype": "bf16 (fp32 residual stream, LayerNorm, softmax, accumulation)",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=16384, help="samples per GPU per step")
    ap.add_argument("--train-batch", type=int, default=512, help="training images per GPU per step")
    ap.add_argument("--text-batch", type=int, default=592,
                    help="text sequences per GPU (secondary metric); 592 x 64 tokens = 296 row tiles = two per SM")
    ap.add_argument("--no-text", action="store_true", help="skip the text secondary metric")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the sweep / eager-PyTorch / dim-2048 / 64-per-GPU training legs and their CPU baselines")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
