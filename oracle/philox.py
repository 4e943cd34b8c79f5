"""numpy restatement of the library's counter-based RNG (csrc/common.cuh) — TEST INFRASTRUCTURE.

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11;
constants as in Random123).  The reference itself draws noise with torch.randn_like
(src/mnist.py:155,178,190), whose stream is not reproducible across devices or batch shards, so
parity tests inject noise; this oracle exists to pin *our* generator: bits exactly, normals to
the tolerance of the device's fast log/sincos.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
DOMAIN_QSAMPLE, DOMAIN_REVERSE, DOMAIN_INIT = 0, 1, 2


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over equal-shaped uint32 arrays. Returns four uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u01(bits):
    """23-bit uniform in (0,1), exactly as csrc/common.cuh: never 0, never 1, exactly representable."""
    return ((bits >> np.uint32(9)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)


def normal_block(seed: int, sample, quad, step: int, domain: int):
    """Four normals per (sample, quad): the restatement of philox_normal4 in common.cuh."""
    sample = np.asarray(sample, dtype=np.uint64)
    quad = np.asarray(quad, dtype=np.uint32)
    sample, quad = np.broadcast_arrays(sample, quad)
    c1 = (sample & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c3 = (np.uint32(domain) | ((sample >> np.uint64(32)).astype(np.uint32) << np.uint32(8))).astype(np.uint32)
    c2 = np.full(quad.shape, step, dtype=np.uint32)
    x, y, z, w = philox4x32_10(quad, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    two_pi = np.float32(6.283185307179586)
    r0 = np.sqrt(np.float32(-2.0) * np.log(u01(x)))
    r1 = np.sqrt(np.float32(-2.0) * np.log(u01(z)))
    a0 = two_pi * u01(y)
    a1 = two_pi * u01(w)
    return np.stack([r0 * np.cos(a0), r0 * np.sin(a0), r1 * np.cos(a1), r1 * np.sin(a1)], axis=-1).astype(np.float32)


def randn(batch: int, inner: int, seed: int, sample_offset: int, step: int, domain: int):
    """[batch, inner] normals exactly as the kernels lay them out (inner % 4 == 0)."""
    assert inner % 4 == 0
    s = (np.arange(batch, dtype=np.uint64) + np.uint64(sample_offset))[:, None]
    q = np.arange(inner // 4, dtype=np.uint32)[None, :]
    return normal_block(seed, s, q, step, domain).reshape(batch, inner)
