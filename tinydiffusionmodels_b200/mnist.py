"""MNIST DDPM on B200 — host-side mirror of the reference's ``src/mnist.py``.

Same public names, argument order, defaults and checkpoint format as the reference
(SURVEY.md §8b); underneath, every device computation goes through libtdm_b200.so:

* ``q_sample``          -> tdm_q_sample / tdm_q_sample_philox            (ref src/mnist.py:36-42)
* ``SimpleUNet.forward``-> tdm_unet_forward (tcgen05 implicit-GEMM convs) (ref src/mnist.py:64-87)
* ``p_sample``          -> tdm_unet_p_sample (reverse step fused in the last conv epilogue)
                                                                          (ref src/mnist.py:167-180)
* ``sample`` / ``sample_images`` -> one captured CUDA graph of a reverse step, replayed T times,
  timestep and Philox noise counters advanced on the device (no host sync in the loop)
                                                                          (ref src/mnist.py:99-126,183-212)
* ``train``             -> fused forward/backward/AdamW step              (ref src/mnist.py:128-165)

There is no CPU path: tensors must live on a CUDA device and the library must be built.
"""
from __future__ import annotations

import argparse
import contextlib
import math
import os

import torch
import torch.nn as nn

from . import _lib, ops
from .schedule import TIMESTEPS, linear_beta_schedule, make_schedule, schedule_on  # noqa: F401
from .unet_engine import PARAM_COUNT, PARAM_SPEC, UNetEngine
from .utils import (get_samples_dir, get_vertex_checkpoint_path, load_checkpoint, save_checkpoint,
                    save_samples)

# ---- schedule tables as module attributes, exactly the names the reference exposes ------------
timesteps = TIMESTEPS
_tables = make_schedule(timesteps)
betas = _tables.betas
alphas = _tables.alphas
alphas_cumprod = _tables.alphas_cumprod
sqrt_alphas_cumprod = _tables.sqrt_alphas_cumprod
sqrt_one_minus_alphas_cumprod = _tables.sqrt_one_minus_alphas_cumprod


def q_sample(x_start: torch.Tensor, t: torch.Tensor, noise=None):
    """Diffuse ``x_start`` to timestep ``t`` (fused gather + axpby kernel)."""
    if noise is None:
        return ops.q_sample(x_start, t, None, seed=_fresh_seed())
    return ops.q_sample(x_start, t, noise)


def _fresh_seed() -> int:
    """A Philox key drawn from torch's CPU generator, so ``torch.manual_seed`` governs our noise
    the way it governs the reference's ``randn_like`` — without touching the device."""
    return int(torch.randint(0, 2**62, (), dtype=torch.int64))


# ---------------------------------------------------------------------------------------------
# modules — parameter containers with the reference's names/shapes; compute is in the engine
# ---------------------------------------------------------------------------------------------
class ResidualBlock(nn.Module):
    """conv3x3-ReLU-(+time bias)-conv3x3-ReLU-(+skip).  Holds the parameters under the reference's
    keys; the arithmetic runs fused inside SimpleUNet's kernels."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, out_ch, 3, padding=1)
        self.conv2 = nn.Conv2d(out_ch, out_ch, 3, padding=1)
        self.time_emb = nn.Linear(1, out_ch)
        self.skip = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else nn.Identity()

    def forward(self, x, t):
        raise _lib.TdmError(
            "ResidualBlock is only executed as part of SimpleUNet's fused kernels; "
            "there is no standalone (or PyTorch fallback) path")


class SimpleUNet(nn.Module):
    """The reference's 4-block UNet (181,473 parameters) over one flat fp32 parameter buffer."""

    def __init__(self):
        super().__init__()
        self.rb1 = ResidualBlock(1, 32)
        self.rb2 = ResidualBlock(32, 64)
        self.rb3 = ResidualBlock(64, 64)
        self.rb4 = ResidualBlock(96, 32)
        self.out = nn.Conv2d(32, 1, kernel_size=1)
        got = [(n, tuple(p.shape)) for n, p in self.named_parameters()]
        assert got == PARAM_SPEC, "parameter registration order must equal the reference state_dict"
        self._flat: torch.Tensor | None = None
        self._flat_grad: torch.Tensor | None = None
        self._engine: UNetEngine | None = None
        self._trainer = None
        self._weights_generation = 0   # bumped by every raw-pointer update of the flat parameters (UNetTrainer.step)

    # -- flat parameter storage ------------------------------------------------------------
    def _is_flat(self) -> bool:
        f = self._flat
        if f is None:
            return False
        off = 0
        for p in self.parameters():
            if p.data_ptr() != f.data_ptr() + 4 * off or p.device != f.device:
                return False
            off += p.numel()
        return True

    def flat_params(self) -> torch.Tensor:
        """The parameters as one contiguous fp32 vector (state_dict order); the nn.Parameters are
        views into it, so optimizers, ``load_state_dict`` and our kernels all see the same bytes."""
        if not self._is_flat():
            params = list(self.parameters())
            flat = torch.empty(PARAM_COUNT, dtype=torch.float32, device=params[0].device)
            off = 0
            for p in params:
                n = p.numel()
                flat[off:off + n].copy_(p.detach().reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                off += n
            self._flat = flat
            self._flat_grad = None
        return self._flat

    def flat_grads(self) -> torch.Tensor:
        """One contiguous gradient buffer with every ``p.grad`` a view into it."""
        flat = self.flat_params()
        g = self._flat_grad
        if g is None or g.device != flat.device:
            g = torch.zeros_like(flat)
            self._flat_grad = g
        off = 0
        for p in self.parameters():
            n = p.numel()
            view = g[off:off + n].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != view.data_ptr():
                p.grad = view
            off += n
        return g

    def engine(self, batch: int) -> UNetEngine:
        flat = self.flat_params()
        if not flat.is_cuda:
            raise _lib.TdmError("SimpleUNet runs on CUDA only (no CPU fallback): call .to('cuda')")
        e = self._engine
        if e is None or e.device != flat.device or e.max_batch < batch:
            e = UNetEngine(flat.device, max(batch, 1))
            self._engine = e
        # key on everything that can change the weights: writes through the parameter views (load_state_dict,
        # torch.optim) bump the views' version counters, the fused trainer / graph replays write through raw
        # pointers and bump _weights_generation instead (flat._version sees neither)
        e.ensure_packed(flat, (flat.data_ptr(), flat._version, sum(p._version for p in self.parameters()),
                               self._weights_generation))
        return e

    def forward(self, x, t):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .unet_train import unet_autograd_forward
            return unet_autograd_forward(self, x, t)
        return self.engine(x.shape[0]).forward(x.float(), t)


# ---------------------------------------------------------------------------------------------
# sampling
# ---------------------------------------------------------------------------------------------
def p_sample(model, x, t):
    """One reverse step x_t -> x_{t-1}; the noise comes from the in-kernel Philox stream."""
    x = x.float()
    return model.engine(x.shape[0]).p_sample(x, t, None, seed=_fresh_seed())


# Samples one engine workspace holds at a time: larger batches run as sequential chunks (a sample's trajectory
# depends only on its global index, so chunking cannot change a bit of the result).  The activations of a reverse
# step take < 1 MB per sample, so a chunk is < 32 GB of the 180 GB of HBM.
SAMPLE_CHUNK = 32768


@torch.no_grad()
def sample_loop(model: SimpleUNet, x: torch.Tensor, *, seed: int, sample_offset: int = 0,
                steps: int = timesteps, use_graph: bool = True, noise: torch.Tensor | None = None) -> torch.Tensor:
    """Run ``steps`` reverse steps (t = steps-1 .. 0) in place on ``x`` and return it (pre-clamp)
    (ref src/mnist.py:190-194).

    Row b's trajectory depends only on (seed, sample_offset + b): the batch may be sharded over
    any number of GPUs / calls and the union of the results is bit-identical.

    ``noise`` (optional, ``(steps, B, 1, 28, 28)`` on the device) injects the per-step posterior noise:
    ``noise[i]`` is used at timestep ``i`` (row 0 is never read: the reference adds none at t = 0) -
    the parity tests feed the oracle's noise through the same captured, device-timestep loop.

    The captured reverse step is cached on the model's engine per (batch, seed, sample_offset): a second call
    with the same key only copies ``x`` into the graph's buffer and replays.
    """
    n = x.shape[0]
    if n == 0:
        return x
    if n > SAMPLE_CHUNK:
        for lo in range(0, n, SAMPLE_CHUNK):
            hi = min(n, lo + SAMPLE_CHUNK)
            xc = x[lo:hi]
            if not xc.is_contiguous():
                raise ValueError("sample_loop: x must be contiguous")
            sample_loop(model, xc, seed=seed, sample_offset=sample_offset + lo, steps=steps, use_graph=use_graph,
                        noise=None if noise is None else noise[:, lo:hi])
        return x
    eng = model.engine(n)
    lib = eng.lib
    dev = x.device
    if noise is not None:
        if noise.shape[0] < steps or noise.shape[1] != n or noise.device != dev:
            raise ValueError("noise must be (steps, B, 1, 28, 28) on the device of x")
        noise = noise.float().contiguous().view(noise.shape[0], -1)

    def one_step(xb, t_buf, zs):
        z = None
        if zs is not None:
            z = zs.index_select(0, t_buf[:1]).view(-1)   # row t of the noise, picked on the device
        eng.p_sample(xb, t_buf, z, out=xb, seed=seed, sample_offset=sample_offset)
        _lib.check(lib.tdm_timestep_advance(t_buf.data_ptr(), n, -1, _lib.stream_ptr(dev)),
                   "tdm_timestep_advance")

    if not use_graph or steps < 4:
        t_buf = torch.full((n,), steps - 1, device=dev, dtype=torch.int64)
        xb = x if x.is_contiguous() and x.dtype == torch.float32 else x.float().contiguous()
        for _ in range(steps):
            one_step(xb, t_buf, noise)
        if xb is not x:
            x.copy_(xb)
        return x
    eng._prep(n)   # a cached graph must not replay into a workspace last laid out for another batch
    key = (n, int(seed), int(sample_offset), None if noise is None else tuple(noise.shape))
    hit = eng._loops.get(key)
    if hit is None:
        xb = torch.empty(n, 1, 28, 28, device=dev, dtype=torch.float32)
        t_buf = torch.empty(n, dtype=torch.int64, device=dev)
        zs = None if noise is None else torch.empty_like(noise)
        xb.copy_(x)
        t_buf.fill_(steps - 1)
        if zs is not None:
            zs.copy_(noise)
        # warm-up outside capture on scratch state (lazy kernel attribute setup)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            one_step(xb, t_buf, zs)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_step(xb, t_buf, zs)   # capture does not execute
        while len(eng._loops) >= 8:   # a handful of (batch, seed) keys stay warm; older ones are dropped
            eng._loops.pop(next(iter(eng._loops)))
        hit = eng._loops[key] = (graph, xb, t_buf, zs)
    graph, xb, t_buf, zs = hit
    xb.copy_(x)
    t_buf.fill_(steps - 1)
    if zs is not None:
        zs.copy_(noise)
    for _ in range(steps):
        graph.replay()
    x.copy_(xb)
    return x


@contextlib.contextmanager
def eval_mode(module: nn.Module):
    """Run a block with ``module`` in eval mode and put its previous mode back (ref src/mnist.py:89-96)."""
    was_training = module.training
    module.eval()
    try:
        yield
    finally:
        module.train(was_training)


def _encode_png(grid_u8_hwc) -> bytes:
    """PNG bytes of an HWC uint8 array, through PIL exactly as torchvision.utils.save_image does."""
    import io

    from PIL import Image

    buf = io.BytesIO()
    Image.fromarray(grid_u8_hwc).save(buf, format="png")
    return buf.getvalue()


def _save_grid(x: torch.Tensor, samples_dir, filename: str, *, from_signed: bool = False):
    """The reference's ``save_image(x, nrow=int(sqrt(n)))`` + ``save_samples`` (src/mnist.py:116-124,196-209):
    the clamp / grid / uint8 conversion runs on the device (``tdm_image_grid_u8``), only the uint8 grid is
    copied to the host, PIL encodes it in memory (no temp file) - same bytes, same file names."""
    nrow = int(math.sqrt(x.shape[0]))
    grid = ops.image_grid_u8(x, nrow=nrow, padding=2, from_signed=from_signed)
    data = _encode_png(grid.cpu().numpy())
    path = f"{samples_dir}/{filename}" if isinstance(samples_dir, str) else samples_dir / filename
    save_samples(data, path, mode="wb")
    return path


def _generate(model, device, n_samples: int, seed: int | None = None) -> torch.Tensor:
    seed = _fresh_seed() if seed is None else seed
    x = ops.randn((n_samples, 1, 28, 28), device, seed=seed, sample_offset=0, stream_id=0)
    return sample_loop(model, x, seed=seed)   # x_0 in [-1, 1] (unclamped); the output step maps it to [0, 1]


def sample_images(model: nn.Module, device: str, epoch: int, n_samples: int = 25, outdir: str = "samples"):
    samples_dir = get_samples_dir(outdir)
    with eval_mode(model), torch.no_grad():
        x = _generate(model, device, n_samples)
        path = _save_grid(x, samples_dir, f"epoch_{epoch:03d}.png", from_signed=True)
    print(f"[epoch {epoch}] saved samples to {path}")


def sample(model: nn.Module, device: str, n_samples=25, ckpt_path="ckpt.pth", outdir="samples"):
    model.load_state_dict(load_checkpoint(ckpt_path, device))
    model.eval()
    samples_dir = get_samples_dir(outdir)
    with torch.no_grad():
        x = _generate(model, device, n_samples)
        path = _save_grid(x, samples_dir, "samples.png", from_signed=True)
    print(f"Saved samples to {path}")


# ---------------------------------------------------------------------------------------------
# training
# ---------------------------------------------------------------------------------------------
def train(model: nn.Module, device: str, epochs: int = 5, batch_size: int = 128, lr: float = 1e-3,
          ckpt_path: str = "ckpt.pth", sample_every_epoch: bool = True, samples_per_epoch: int = 25,
          *, synthetic: bool = False, steps_per_epoch: int | None = None, log_every: int = 50):
    """DDPM training (ref src/mnist.py:128-165): per batch t~U{0..T-1}, noise~N(0,I), q_sample,
    UNet, MSE, AdamW(lr, torch defaults) — one fused device step per batch.

    ``synthetic=True`` trains on U(-1,1) MNIST-shaped data (no dataset download; there is no
    network on the benchmark boxes)."""
    from .unet_train import UNetTrainer

    from .data import _rank_world

    if "AIP_MODEL_DIR" in os.environ:
        ckpt_path = get_vertex_checkpoint_path("image-model.pth")
    rank, world = _rank_world()
    if not synthetic:
        # decode (and, on rank 0 only, download) the dataset BEFORE the trainer exists: the ranks next meet inside the
        # optimizer kernel's flag wait, which must not also have to cover a dataset download
        _dataset_for(device, rank, world)
    trainer = UNetTrainer(model, lr=lr, max_batch=batch_size)
    for epoch in range(epochs):
        running, seen = None, 0
        for step, x in enumerate(_batches(device, batch_size, synthetic, steps_per_epoch, epoch)):
            loss = trainer.step(x)          # device scalar; no host sync
            running = loss if running is None else running + loss
            seen += 1
            if (step + 1) % log_every == 0:
                print(f"Epoch {epoch + 1}/{epochs} step {step + 1}: loss={float(running) / seen:.4f}")
                running, seen = None, 0
        if sample_every_epoch and rank == 0:   # the replicas are bit-identical: one of them draws the samples
            sample_images(model, device, epoch + 1, samples_per_epoch)
    if rank == 0:
        save_checkpoint(model.state_dict(), ckpt_path)
    if world > 1:
        torch.distributed.barrier()   # nobody reads the checkpoint before rank 0 has written it


def _batches(device, batch_size, synthetic, steps_per_epoch, epoch):
    if synthetic:
        from .data import _rank_world

        g = torch.Generator(device=device).manual_seed(1234 + epoch + 7919 * _rank_world()[0])   # ranks see different data
        for _ in range(steps_per_epoch or 469):
            yield torch.rand(batch_size, 1, 28, 28, device=device, generator=g) * 2 - 1
        return
    from .data import mnist_on_device

    # the reference's DataLoader + ToTensor + Normalize((0.5,), (0.5,)) (src/mnist.py:139-147), with the uint8
    # training set resident on the device: a permutation per epoch, one gather + normalise kernel per batch
    from .data import _rank_world

    ds = _dataset_for(device, *_rank_world())
    # like the reference's DataLoader, the visiting order follows torch's global seed (torch.manual_seed)
    yield from ds.batches(batch_size, seed=torch.initial_seed(), epoch=epoch, max_batches=steps_per_epoch)


_datasets: dict = {}


def _dataset_for(device, rank: int = 0, world: int = 1):
    """The uint8 training set resident on ``device`` (cached).  With several ranks sharing ``./data``, rank 0
    downloads / decodes first and the others read the finished files after a barrier."""
    from .data import mnist_on_device

    dev = torch.empty(0, device=device).device   # resolved ("cuda" -> "cuda:0")
    if dev not in _datasets:
        if world > 1 and rank != 0:
            torch.distributed.barrier()
        _datasets[dev] = mnist_on_device(dev, download=(rank == 0))
        if world > 1 and rank == 0:
            torch.distributed.barrier()
    return _datasets[dev]


# ---------------------------------------------------------------------------------------------
# CLI — the reference's flags and defaults (src/mnist.py:215-241) plus additive ones
# ---------------------------------------------------------------------------------------------
def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--train", action="store_true", help="Train the model")
    parser.add_argument("--sample", action="store_true", help="Generate samples")
    parser.add_argument("--epochs", type=int, default=3)
    parser.add_argument("--batch_size", type=int, default=128)
    parser.add_argument("--ckpt", type=str,
                        default=get_vertex_checkpoint_path("image-model.pth") if "AIP_MODEL_DIR" in os.environ else "ckpt.pth")
    # additive flags (defaults reproduce the reference behaviour)
    parser.add_argument("--n_samples", type=int, default=25, help="samples to draw with --sample")
    parser.add_argument("--synthetic", action="store_true", help="train on synthetic U(-1,1) images")
    parser.add_argument("--steps_per_epoch", type=int, default=None)
    parser.add_argument("--seed", type=int, default=None)
    args = parser.parse_args(argv)

    if not torch.cuda.is_available():
        raise _lib.TdmError("tinydiffusionmodels_b200 needs a CUDA (sm_100a) device; there is no CPU path")
    device = "cuda"
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # torchrun: one process per GPU, data-parallel training (gradient exchange inside the optimizer kernel)
        from . import dist as tdist

        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        device = f"cuda:{local}"
        tdist.init_from_env(device=torch.device(device))
    if args.seed is not None:
        torch.manual_seed(args.seed)
    model = SimpleUNet().to(device)

    if args.train:
        train(model, device, epochs=args.epochs, batch_size=args.batch_size, ckpt_path=args.ckpt,
              synthetic=args.synthetic, steps_per_epoch=args.steps_per_epoch)
    if args.sample and (not torch.distributed.is_initialized() or torch.distributed.get_rank() == 0):
        sample(model, device, n_samples=args.n_samples, ckpt_path=args.ckpt)
    if not args.train and not args.sample:
        print("Nothing to do. Pass --train or --sample.")


if __name__ == "__main__":
    main()
