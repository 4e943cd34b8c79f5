"""GPU parity of the elementwise diffusion kernels against the CPU oracle (bit-exact fp32)."""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import philox as PX
from tinydiffusionmodels_b200 import ops

pytestmark = pytest.mark.gpu
TAB = O.make_tables()


@pytest.mark.parametrize("shape", [(1, 1, 28, 28), (64, 1, 28, 28), (513, 1, 28, 28), (5, 64, 256), (3, 7, 4)])
def test_q_sample_bit_exact(cuda, shape):
    g = torch.Generator().manual_seed(1)
    x0 = torch.randn(shape, generator=g)
    n = torch.randn(shape, generator=g)
    t = torch.randint(0, 1000, (shape[0],), generator=g)
    ref = O.q_sample(x0, t, n, TAB)
    got = ops.q_sample(x0.to(cuda), t.to(cuda), n.to(cuda)).cpu()
    assert torch.equal(got, ref)


@pytest.mark.parametrize("shape", [(64, 1, 28, 28), (5, 64, 256), (2, 3, 8)])
@pytest.mark.parametrize("tval", [999, 500, 1, 0])
def test_reverse_step_bit_exact(cuda, shape, tval):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(shape, generator=g) * 3
    e = torch.randn(shape, generator=g)
    z = torch.randn(shape, generator=g)
    t = torch.full((shape[0],), tval, dtype=torch.long)
    ref = O.reverse_step(x, e, t, z, TAB)
    got = ops.reverse_step(x.to(cuda), e.to(cuda), t.to(cuda), z.to(cuda)).cpu()
    assert torch.equal(got, ref)


def test_reverse_step_mixed_t_follows_t0(cuda):
    # the reference branches on t[0] only (src/mnist.py:176): t = [0, 5, 9] adds no noise anywhere
    g = torch.Generator().manual_seed(3)
    x, e, z = (torch.randn(3, 16, generator=g) for _ in range(3))
    t = torch.tensor([0, 5, 9])
    assert torch.equal(ops.reverse_step(x.to(cuda), e.to(cuda), t.to(cuda), z.to(cuda)).cpu(),
                       O.reverse_step(x, e, t, z, TAB))
    t = torch.tensor([7, 0, 9])
    assert torch.equal(ops.reverse_step(x.to(cuda), e.to(cuda), t.to(cuda), z.to(cuda)).cpu(),
                       O.reverse_step(x, e, t, z, TAB))


def test_philox_randn_matches_numpy_oracle(cuda):
    got = ops.randn((37, 784), cuda, seed=0x1234_5678_9ABC, sample_offset=(1 << 33) + 5, stream_id=7).cpu().numpy()
    ref = PX.randn(37, 784, 0x1234_5678_9ABC, (1 << 33) + 5, 7, PX.DOMAIN_INIT)
    # bits are exact; the normals differ only by the device's fast log/sincos (MUFU)
    np.testing.assert_allclose(got, ref, rtol=0, atol=2e-5)
    assert abs(got.mean()) < 0.02 and abs(got.std() - 1) < 0.02


def test_philox_normals_finite_over_2_pow_26_draws(cuda):
    # 2^26 normals use 2^26 uniforms: every extreme of the 23-bit uniform grid is hit many times
    z = ops.randn((8192, 8192), cuda, seed=99)
    assert torch.isfinite(z).all()
    assert float(z.abs().max()) < 6.0          # sqrt(-2 ln 2^-24) = 5.77
    assert abs(float(z.mean())) < 1e-3 and abs(float(z.std()) - 1) < 1e-3


def test_philox_shard_invariance(cuda):
    # rows depend on (seed, global sample index) only: any split of the batch gives the same rows
    full = ops.randn((64, 784), cuda, seed=9, sample_offset=0, stream_id=3)
    lo = ops.randn((24, 784), cuda, seed=9, sample_offset=0, stream_id=3)
    hi = ops.randn((40, 784), cuda, seed=9, sample_offset=24, stream_id=3)
    assert torch.equal(full, torch.cat([lo, hi]))


def test_reverse_step_philox_matches_oracle_noise(cuda):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(8, 784, generator=g)
    e = torch.randn(8, 784, generator=g)
    t = torch.full((8,), 321, dtype=torch.long)
    z = torch.from_numpy(PX.randn(8, 784, 77, 100, 321 + 1000, PX.DOMAIN_REVERSE))
    ref = O.reverse_step(x, e, t, z, TAB)
    got = ops.reverse_step(x.to(cuda), e.to(cuda), t.to(cuda), None, seed=77, sample_offset=100, step_id=1000).cpu()
    torch.testing.assert_close(got, ref, rtol=0, atol=1e-5)


def test_q_sample_philox(cuda):
    g = torch.Generator().manual_seed(5)
    x0 = torch.rand(16, 1, 28, 28, generator=g) * 2 - 1
    t = torch.randint(0, 1000, (16,), generator=g)
    out, noise = ops.q_sample(x0.to(cuda), t.to(cuda), None, seed=5, sample_offset=32, stream_id=11, return_noise=True)
    ref_noise = torch.from_numpy(PX.randn(16, 784, 5, 32, 11, PX.DOMAIN_QSAMPLE)).view(16, 1, 28, 28)
    torch.testing.assert_close(noise.cpu(), ref_noise, rtol=0, atol=2e-5)
    # given the noise the kernel actually drew, the combination is bit-exact
    assert torch.equal(out.cpu(), O.q_sample(x0, t, noise.cpu(), TAB))


def test_to_unit_range(cuda):
    x = torch.linspace(-3, 3, 1001)
    assert torch.equal(ops.to_unit_range(x.to(cuda)).cpu(), O.to_unit_range(x))


def test_empty_batch(cuda):
    x = torch.empty(0, 1, 28, 28, device=cuda)
    t = torch.empty(0, dtype=torch.long, device=cuda)
    assert ops.q_sample(x, t, x.clone()).shape == (0, 1, 28, 28)
