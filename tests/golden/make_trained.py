"""Train the REAL reference UNet (/root/reference, src/mnist.py:45-87 and its training step :153-160) for a few hundred
AdamW steps on synthetic 'digits' and store the checkpoint: tests/golden/mnist_trained.pt.

Run in the authoring container only:    python tests/golden/make_trained.py        (about a minute of CPU)

Random-init reverse trajectories blow up, so the full T = 1000 final-sample parity test
(tests/test_gpu_trajectory.py) needs a trained denoiser, and it needs the SAME one on every run: a checkpoint trained on
the GPU differs from run to run (the backward's fp32 atomics), and with it the sensitivity of the 1000-step trajectory.
There is no dataset here (no network), so the images are smooth synthetic strokes in [-1, 1].
"""
import sys
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import import_reference  # noqa: E402


def blobs(n, gen):
    """A few smooth strokes per image in [-1, 1] (also used by the test for nothing but documentation)."""
    yy, xx = torch.meshgrid(torch.arange(28.0), torch.arange(28.0), indexing="ij")
    img = torch.zeros(n, 28, 28)
    for _ in range(3):
        cx, cy = torch.rand(n, 1, 1, generator=gen) * 16 + 6, torch.rand(n, 1, 1, generator=gen) * 16 + 6
        sx, sy = torch.rand(n, 1, 1, generator=gen) * 3 + 1, torch.rand(n, 1, 1, generator=gen) * 3 + 1
        img = torch.maximum(img, torch.exp(-((xx - cx) ** 2 / (2 * sx ** 2) + (yy - cy) ** 2 / (2 * sy ** 2))))
    return (img * 2 - 1).unsqueeze(1)


def main():
    ref, _ = import_reference()
    torch.manual_seed(5)
    model = ref.SimpleUNet()
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3)      # the reference's optimizer (src/mnist.py:137), larger lr
    gen = torch.Generator().manual_seed(6)
    first = last = None
    for it in range(400):
        x0 = blobs(128, gen)
        t = torch.randint(0, ref.timesteps, (x0.shape[0],), generator=gen).long()
        noise = torch.randn(x0.shape, generator=gen)
        x_noisy = ref.q_sample(x0, t, noise)                  # src/mnist.py:155-156
        loss = torch.nn.functional.mse_loss(model(x_noisy, t), noise)
        opt.zero_grad()
        loss.backward()
        opt.step()
        first = float(loss) if first is None else first
        last = float(loss)
        if it % 50 == 0:
            print(f"step {it}: loss {last:.4f}", flush=True)
    print(f"trained 400 steps on the CPU with the reference's model: loss {first:.4f} -> {last:.4f}")
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    torch.save({"state_dict": sd, "loss_first": first, "loss_last": last}, HERE / "mnist_trained.pt")
    print("wrote", HERE / "mnist_trained.pt")


if __name__ == "__main__":
    main()
