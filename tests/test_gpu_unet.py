"""GPU parity of the tcgen05 UNet forward / fused p_sample against the CPU fp32 oracle.

Tolerances (stated, per north_star): the kernels multiply in bf16 with fp32 accumulation and
store inter-layer activations in bf16, the oracle is fp32 end to end.  A single forward on
random-init weights differs by <= 1% of the output rms in rms terms and <= 8% of rms in the worst
element (SURVEY.md §A.2 measured 6.8e-3 max-abs on rms 0.138 for bf16 autocast).
"""
import pytest
import torch

from oracle import ddpm_oracle as O
from tests.helpers import random_unet_state_dict, rel_rms
from tinydiffusionmodels_b200.unet_engine import UNetEngine, read_activation

pytestmark = pytest.mark.gpu
TAB = O.make_tables()
RMS_TOL = 1.0e-2
MAX_TOL = 8.0e-2


@pytest.fixture
def unfused():
    """Run the layer-by-layer kernels (every intermediate lands in the workspace) for the duration of a test."""
    from tinydiffusionmodels_b200 import _lib
    lib = _lib.load()
    prev = lib.tdm_unet_set_fused(0)
    yield
    lib.tdm_unet_set_fused(prev)


def _oracle_intermediates(sd, x, t):
    import torch.nn.functional as F
    tt = (t.float() / 1000).view(-1, 1, 1, 1)
    out = {}
    def rb(p, xin):
        h = F.relu(F.conv2d(xin, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=1))
        h = h + F.linear(tt.view(-1, 1), sd[f"{p}.time_emb.weight"], sd[f"{p}.time_emb.bias"]).view(x.shape[0], -1, 1, 1)
        out[p + ".t"] = h
        h2 = F.relu(F.conv2d(h, sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1))
        if f"{p}.skip.weight" in sd:
            s = F.conv2d(xin, sd[f"{p}.skip.weight"], sd[f"{p}.skip.bias"])
        else:
            s = xin
        out[p + ".s"] = s
        return h2 + s
    h1 = rb("rb1", x); out["h1"] = h1
    p1 = F.avg_pool2d(h1, 2); out["p1"] = p1
    h2 = rb("rb2", p1); out["h2"] = h2
    h3 = rb("rb3", h2); out["h3"] = h3
    cat = torch.cat([F.interpolate(h3, scale_factor=2, mode="nearest"), h1], 1); out["cat"] = cat
    h4 = rb("rb4", cat); out["h4"] = h4
    return out


@pytest.mark.parametrize("batch", [1, 3, 64])
def test_unet_forward_matches_oracle(cuda, batch, unfused):
    sd = random_unet_state_dict(0)
    g = torch.Generator().manual_seed(10 + batch)
    x = torch.randn(batch, 1, 28, 28, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    ref = O.unet_forward(sd, x, t)
    eng = UNetEngine(cuda, batch)
    eng.load_state_dict(sd)
    got = eng.forward(x.to(cuda), t.to(cuda))
    torch.cuda.synchronize()
    # layer-by-layer report first: a failure names the first layer that diverges
    inter = _oracle_intermediates(sd, x, t)
    # (workspace buffer, oracle tensor, channel slice of the buffer): in the sampling path the concat
    # buffer only holds h1 (channels 64..95); h3 stays at 14x14 and is upsampled inside rb4.conv1
    checks = [("t1", "rb1.t", None), ("cat", "h1", slice(64, 96)), ("p1", "p1", None), ("t2", "rb2.t", None),
              ("s2", "rb2.s", None), ("h2", "h2", None), ("t3", "rb3.t", None), ("h3", "h3", None),
              ("t4", "rb4.t", None), ("s4", "rb4.s", None)]
    report = []
    for ws_name, key, sl in checks:
        a = read_activation(eng, ws_name, batch).cpu()
        if sl is not None:
            a = a[:, sl]
        report.append((ws_name, rel_rms(a, inter[key])))
    msg = ", ".join(f"{n}={e:.2e}" for n, e in report)
    print("layer rel-rms:", msg)
    for n, e in report:
        assert e < 2e-2, f"layer {n} diverges: {msg}"
    got = got.cpu()
    rms = ref.pow(2).mean().sqrt()
    err_rms = (got - ref).pow(2).mean().sqrt() / rms
    err_max = (got - ref).abs().max() / rms
    print(f"eps rel-rms {err_rms:.3e} max/rms {err_max:.3e}")
    assert err_rms < RMS_TOL and err_max < MAX_TOL


def test_unet_forward_repeatable_and_batch_independent(cuda):
    sd = random_unet_state_dict(1)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(130, 1, 28, 28, generator=g).to(cuda)
    t = torch.randint(0, 1000, (130,), generator=g).to(cuda)
    eng = UNetEngine(cuda, 130)
    eng.load_state_dict(sd)
    a = eng.forward(x, t).clone()
    b = eng.forward(x, t).clone()
    assert torch.equal(a, b)
    # a sample's output does not depend on its neighbours in the batch or its position in it
    c = eng.forward(x[7:20].contiguous(), t[7:20].contiguous())
    assert torch.equal(c, a[7:20])


@pytest.mark.parametrize("tval", [999, 1, 0])
def test_fused_p_sample_matches_oracle(cuda, tval):
    sd = random_unet_state_dict(2)
    g = torch.Generator().manual_seed(20 + tval)
    x = torch.randn(16, 1, 28, 28, generator=g)
    z = torch.randn(16, 1, 28, 28, generator=g)
    t = torch.full((16,), tval, dtype=torch.long)
    ref = O.mnist_p_sample(sd, x, t, z, TAB)
    eng = UNetEngine(cuda, 16)
    eng.load_state_dict(sd)
    got = eng.p_sample(x.to(cuda), t.to(cuda), z.to(cuda)).cpu()
    # x_{t-1} = c1*(x - c2*eps) + sigma*z with c2 <= 0.02: the bf16 eps error is scaled by c2
    torch.testing.assert_close(got, ref, rtol=0, atol=2e-3)
    # and the fused epilogue equals forward + standalone reverse step exactly
    from tinydiffusionmodels_b200 import ops
    eps = eng.forward(x.to(cuda), t.to(cuda))
    two = ops.reverse_step(x.to(cuda), eps, t.to(cuda), z.to(cuda)).cpu()
    assert torch.equal(got, two)
    # in place
    xin = x.to(cuda).clone()
    eng.p_sample(xin, t.to(cuda), z.to(cuda), out=xin)
    assert torch.equal(xin.cpu(), got)


def _forward_without_mirror(eng, sd, x, t, cuda):
    """Re-pack through tdm_unet_pack_weights (drops the host mirror) and run the forward again."""
    from tinydiffusionmodels_b200 import _lib
    from tinydiffusionmodels_b200.unet_engine import flatten_state_dict
    flat = flatten_state_dict(sd, cuda)
    _lib.check(eng.lib.tdm_unet_pack_weights(flat.data_ptr(), eng.wpack.data_ptr(), _lib.stream_ptr(cuda)),
               "tdm_unet_pack_weights")
    out = eng.forward(x, t)
    torch.cuda.synchronize()
    return out


def test_host_mirror_path_is_bit_identical_to_smem_path(cuda, unfused):
    """Per-channel vectors by value (constant bank, tdm_unet_pack_weights_host) vs staged in shared
    memory (tdm_unet_pack_weights): same arithmetic, so the outputs must be equal bit for bit."""
    sd = random_unet_state_dict(3)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(37, 1, 28, 28, generator=g).to(cuda)
    t = torch.randint(0, 1000, (37,), generator=g).to(cuda)
    eng = UNetEngine(cuda, 37)
    eng.load_state_dict(sd)                      # registers the host mirror
    with_mirror = eng.forward(x, t).clone()
    torch.cuda.synchronize()
    without = _forward_without_mirror(eng, sd, x, t, cuda)
    assert torch.equal(with_mirror, without)


def test_repack_drops_a_stale_host_mirror(cuda):
    """Loading new weights through the plain pack entry point must not leave the old biases in the
    launch arguments: the result has to follow the NEW parameters."""
    sd_a, sd_b = random_unet_state_dict(4), random_unet_state_dict(5)
    g = torch.Generator().manual_seed(78)
    x = torch.randn(5, 1, 28, 28, generator=g)
    t = torch.randint(0, 1000, (5,), generator=g)
    eng = UNetEngine(cuda, 5)
    eng.load_state_dict(sd_a)                    # mirror of A
    eng.forward(x.to(cuda), t.to(cuda))
    got_b = _forward_without_mirror(eng, sd_b, x.to(cuda), t.to(cuda), cuda)   # device weights B, mirror dropped
    ref_b = O.unet_forward(sd_b, x, t)
    assert rel_rms(got_b.cpu(), ref_b) < RMS_TOL
    eng.load_state_dict(sd_a)                    # and back, with a fresh mirror
    got_a = eng.forward(x.to(cuda), t.to(cuda))
    torch.cuda.synchronize()
    assert rel_rms(got_a.cpu(), O.unet_forward(sd_a, x, t)) < RMS_TOL


def test_first_conv_keeps_fp32_like_precision(cuda, unfused):
    """rb1.conv1 runs on the tensor pipe as hi/lo bf16 products (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo):
    its output is then rounded to bf16 once, so it must sit at the bf16 rounding floor (2^-9 relative),
    well below the 2e-2 per-layer bar - also for inputs far outside the unit range."""
    import torch.nn.functional as F
    sd = random_unet_state_dict(6)
    g = torch.Generator().manual_seed(79)
    x = torch.randn(9, 1, 28, 28, generator=g) * 50.0
    t = torch.randint(0, 1000, (9,), generator=g)
    eng = UNetEngine(cuda, 9)
    eng.load_state_dict(sd)
    eng.forward(x.to(cuda), t.to(cuda))
    torch.cuda.synchronize()
    got = read_activation(eng, "t1", 9).cpu()
    tt = (t.float() / 1000).view(-1, 1)
    ref = F.relu(F.conv2d(x, sd["rb1.conv1.weight"], sd["rb1.conv1.bias"], padding=1))
    ref = ref + F.linear(tt, sd["rb1.time_emb.weight"], sd["rb1.time_emb.bias"]).view(9, -1, 1, 1)
    assert rel_rms(got, ref) < 3e-3


@pytest.mark.parametrize("batch", [1, 2, 7, 64, 300])
def test_fused_blocks_match_layer_by_layer_and_oracle(cuda, batch):
    """The fused 28x28 blocks (conv1 -> shared memory -> conv2 in one kernel, csrc/resblock_tc.cuh) against the
    layer-by-layer kernels and the fp32 oracle.  Same bf16 operands and the same fp32 accumulators; only the order
    in which the taps are accumulated differs (ky = 1 first), so the two CUDA schedules agree to ~1e-3 of the output
    rms, far inside the 1e-2 bar against the oracle.  Batches chosen so that a CTA band is shorter than, equal to
    and much longer than the four-slot ring, and so that the last tile is ragged."""
    from tinydiffusionmodels_b200 import _lib
    lib = _lib.load()
    sd = random_unet_state_dict(31)
    g = torch.Generator().manual_seed(900 + batch)
    x = torch.randn(batch, 1, 28, 28, generator=g)
    t = torch.randint(0, 1000, (batch,), generator=g)
    z = torch.randn(batch, 1, 28, 28, generator=g)
    eng = UNetEngine(cuda, batch)
    eng.load_state_dict(sd)
    prev = lib.tdm_unet_set_fused(1)
    try:
        fused = eng.forward(x.to(cuda), t.to(cuda)).cpu()
        fused_step = eng.p_sample(x.to(cuda), t.to(cuda), z.to(cuda)).cpu()
        again = eng.forward(x.to(cuda), t.to(cuda)).cpu()
        lib.tdm_unet_set_fused(0)
        plain = eng.forward(x.to(cuda), t.to(cuda)).cpu()
        plain_step = eng.p_sample(x.to(cuda), t.to(cuda), z.to(cuda)).cpu()
    finally:
        lib.tdm_unet_set_fused(prev)
    assert torch.equal(fused, again)                       # deterministic
    ref = O.unet_forward(sd, x, t)
    rms = ref.pow(2).mean().sqrt()
    print(f"B={batch}: fused vs layer-by-layer rel-rms {rel_rms(fused, plain):.2e}; fused vs oracle {rel_rms(fused, ref):.2e}, "
          f"max/rms {float((fused - ref).abs().max() / rms):.2e}")
    assert rel_rms(fused, plain) < 3e-3
    assert rel_rms(fused, ref) < RMS_TOL and float((fused - ref).abs().max() / rms) < MAX_TOL
    torch.testing.assert_close(fused_step, O.mnist_p_sample(sd, x, t, z, TAB), rtol=0, atol=2e-3)
    torch.testing.assert_close(fused_step, plain_step, rtol=0, atol=1e-3)


def test_gap_to_the_fp32_oracle_is_the_operand_dtype(cuda):
    """The CUDA forward against the oracle that rounds to bf16 at the kernels' own rounding points
    (oracle.unet_forward_bf16_points): it must sit several times closer to that than to the fp32 oracle - what is left
    is the summation order inside the tensor core and the hi/lo split of rb1.conv1.  This is the evidence that the
    1e-2 parity bar is the dtype the north star asks for and not a hidden bug."""
    sd = random_unet_state_dict(77)
    g = torch.Generator().manual_seed(78)
    x = torch.randn(64, 1, 28, 28, generator=g)
    t = torch.randint(0, 1000, (64,), generator=g)
    eng = UNetEngine(cuda, 64)
    eng.load_state_dict(sd)
    got = eng.forward(x.to(cuda), t.to(cuda)).cpu()
    ref, ref_q = O.unet_forward(sd, x, t), O.unet_forward_bf16_points(sd, x, t)
    e, e_q, d = rel_rms(got, ref), rel_rms(got, ref_q), rel_rms(ref_q, ref)
    print(f"CUDA vs fp32 oracle {e:.2e}; CUDA vs bf16-rounding-point oracle {e_q:.2e}; that oracle vs fp32 {d:.2e}")
    assert e < RMS_TOL
    assert e_q < 0.5 * e and e_q < 1.5e-3                 # measured 3.1e-3 (fp32) / 5.9e-4 (rounding points)
