// text_train.cu — one optimisation step of the Shakespeare embedding-space diffusion model
// (src/shakespeare.py:221-250): x0 = embedding_fn(ids), t ~ U{0..T-1}, noise ~ N(0,1), x_noisy = q_sample,
// noise_pred = TinyTransformer(x_noisy, t) in TRAIN mode (dropout on), diffusion_loss = mse, logits =
// rounding_fn(x0), rounding_loss = cross_entropy, total = diffusion + w * rounding, backward, AdamW.
//
// Shape of the work.  With M = B*L token rows (2,048 at the reference's batch 32 x 64), width D and vocabulary V
// (256,000 for Gemma's tokenizer) the step is dominated by the three vocabulary-sized contractions of the
// rounding head (2*M*D*V FLOP each) and by AdamW over the 2*V*D embedding / decoder parameters; the encoder's
// 12 GEMMs per layer are ~1 % of the FLOP.  So:
//   * every contraction — encoder forward, dX and dW of every Linear, the rounding head — is the tcgen05 GEMM of
//     gemm_tc.cuh; operands are converted from the canonical fp32 row-major activations to bf16 planes
//     ("k-planes" [K/8][rows][8] along the reduction index) by two small packing kernels;
//   * the (M, V) logits are never written: pass 1 (GE_LSE) keeps an online (max, sum exp) per row and picks up
//     the target's logit, pass 2 (GE_DLOGITS) recomputes the logits and writes d loss / d logits as bf16
//     planes in both orientations (the second one turned through shared memory in the epilogue): the A operands
//     of dX0 = dlogits . W (K split over the SMs) and of dW = dlogits^T . X0;
//   * dropout masks are never stored: they are a pure function (Philox4x32-10) of (seed, step, site, element)
//     and are recomputed in the backward pass — the oracle regenerates the same bits in numpy;
//   * LayerNorm, attention (L <= 128 keys per sequence: 4*L*D FLOP per token, 2 % of the encoder) and the
//     element-wise glue are SIMT kernels over fp32 rows; reductions that feed parameters are two-stage and
//     run-to-run deterministic except the embedding scatter (float atomics over repeated token ids).
#include "common.cuh"
#include "gemm_tc.cuh"
#include "tc05.cuh"

namespace tdm {

constexpr uint32_t kDomainDropout = 3;
constexpr uint32_t kDomainTimestep = 4;
constexpr int kTrFF = 2048;   // nn.TransformerEncoderLayer default dim_feedforward
constexpr int kTrHeads = 4;
constexpr uint32_t kSiteInput = 0xFFFF0000u;
// dropout sites of encoder layer li: 16*li + {1: attention weights, 2: after out_proj, 3: inside the FFN, 4: after linear2}

struct Rng {
    uint64_t seed;
    const int64_t* step_dev;   // optimiser step counter (device): part of every Philox counter, so graph replays differ
    uint32_t thresh;           // keep an element iff its 32 random bits >= thresh (= p * 2^32); 0 keeps everything
    float keep_scale;          // 1 / (1 - p)
};

__device__ __forceinline__ uint32_t rng_step(const Rng& r) { return (uint32_t)__ldg(r.step_dev); }

__device__ __forceinline__ Philox4 drop_bits(const Rng& r, uint32_t step, uint32_t site, uint32_t quad) {
    return philox4x32_10(quad, site, step, kDomainDropout, (uint32_t)r.seed, (uint32_t)(r.seed >> 32));
}
__device__ __forceinline__ bool drop_keep1(const Rng& r, uint32_t step, uint32_t site, uint64_t idx) {
    if (r.thresh == 0) return true;
    const Philox4 b = drop_bits(r, step, site, (uint32_t)(idx >> 2));
    const uint32_t c = (uint32_t)idx & 3u;
    const uint32_t bits = c == 0 ? b.x : c == 1 ? b.y : c == 2 ? b.z : b.w;
    return bits >= r.thresh;
}

// ---------------------------------------------------------------------------------------------
// t ~ U{0..T-1} per sequence (src/shakespeare.py:230): floor(bits * T / 2^32) of one Philox word
// ---------------------------------------------------------------------------------------------
__global__ void draw_t_kernel(int64_t* __restrict__ t, int64_t B, int T, Rng r, uint64_t sample_offset) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const uint64_t s = sample_offset + (uint64_t)b;
    const Philox4 p = philox4x32_10((uint32_t)s, (uint32_t)(s >> 32), rng_step(r), kDomainTimestep, (uint32_t)r.seed,
                                    (uint32_t)(r.seed >> 32));
    t[b] = (int64_t)(((uint64_t)p.x * (uint64_t)T) >> 32);
}

// ---------------------------------------------------------------------------------------------
// x0 = table[ids]; noise; x_noisy = sqrt_acp[t] x0 + sqrt_om[t] noise; h0 = dropout(x_noisy + time_emb(t/T))
// (src/shakespeare.py:226-232, 115-119).  One block per token row.
// ---------------------------------------------------------------------------------------------
struct EmbedArgs {
    const float* table;
    const int64_t* ids;
    const int64_t* t;
    const float* sqrt_acp;
    const float* sqrt_om;
    const float* noise_in;   // injected noise [M][D] or null -> Philox (keyed like tdm_q_sample_philox, inner = L*D)
    const float* tw;
    const float* tb;
    float* x0;
    float* noise;
    float* h0;
    int64_t V;
    int L, D;
    uint64_t sample_offset;
    Rng rng;
    int* bad;
};

__global__ void __launch_bounds__(64) embed_noise_kernel(const EmbedArgs a) {
    const int64_t m = blockIdx.x;
    const int64_t b = m / a.L;
    const int l = (int)(m - b * a.L);
    const int64_t id = __ldg(a.ids + m);
    const bool ok = id >= 0 && id < a.V;
    if (!ok && threadIdx.x == 0) atomicExch(a.bad, 1);
    const int64_t tb_ = __ldg(a.t + b);
    const float ca = __ldg(a.sqrt_acp + tb_), cb = __ldg(a.sqrt_om + tb_);
    const float ts = (float)tb_ / 1000.0f;
    const uint32_t step = rng_step(a.rng);
    const int D4 = a.D / 4;
    for (int q = threadIdx.x; q < D4; q += blockDim.x) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) x = __ldg(reinterpret_cast<const float4*>(a.table + id * a.D) + q);
        float4 n;
        if (a.noise_in) n = __ldg(reinterpret_cast<const float4*>(a.noise_in + m * a.D) + q);
        else n = philox_normal4(a.rng.seed, a.sample_offset + (uint64_t)b, (uint32_t)(l * D4 + q), step, kDomainQSample);
        float4 xn;
        xn.x = __fadd_rn(__fmul_rn(ca, x.x), __fmul_rn(cb, n.x));
        xn.y = __fadd_rn(__fmul_rn(ca, x.y), __fmul_rn(cb, n.y));
        xn.z = __fadd_rn(__fmul_rn(ca, x.z), __fmul_rn(cb, n.z));
        xn.w = __fadd_rn(__fmul_rn(ca, x.w), __fmul_rn(cb, n.w));
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(a.tw) + q);
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.tb) + q);
        float4 h = make_float4(xn.x + fmaf(w4.x, ts, b4.x), xn.y + fmaf(w4.y, ts, b4.y), xn.z + fmaf(w4.z, ts, b4.z),
                               xn.w + fmaf(w4.w, ts, b4.w));
        if (a.rng.thresh) {
            const Philox4 k = drop_bits(a.rng, step, kSiteInput, (uint32_t)(m * D4 + q));
            const float s = a.rng.keep_scale;
            h.x = k.x >= a.rng.thresh ? h.x * s : 0.f;
            h.y = k.y >= a.rng.thresh ? h.y * s : 0.f;
            h.z = k.z >= a.rng.thresh ? h.z * s : 0.f;
            h.w = k.w >= a.rng.thresh ? h.w * s : 0.f;
        }
        reinterpret_cast<float4*>(a.x0 + m * a.D)[q] = x;
        reinterpret_cast<float4*>(a.noise + m * a.D)[q] = n;
        reinterpret_cast<float4*>(a.h0 + m * a.D)[q] = h;
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 rows -> bf16 planes
//   k-planes  [C/8][Rp][8]:  element (r, c) at plane c/8, row r  — operand whose reduction index is the column
//   m-planes  [Rp/8][Cp][8]: element (r, c) at plane r/8, row c  — operand whose reduction index is the row
// (rows / columns past the valid range are written as zeros: they are tile padding or reduction padding)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_kplanes_kernel(const float* __restrict__ x, int64_t R, int C, int64_t ld, int64_t Rp,
                                                           uint8_t* __restrict__ out) {
    // one block = 32 rows x 64 columns (eight planes) through shared memory: the reads are 256 contiguous bytes per row
    // (eight threads x 32 B), the writes 512 contiguous bytes per plane (32 rows x 16 B).  Thread = (row, plane) on the
    // way in, (plane, row) on the way out; planes are padded to 33 rows in shared memory (conflict-free both ways).
    // (Thread = one output row directly, the reads were 32 B per thread a whole row apart: 2.8 TB/s on the 131 M-element
    // decoder matrix.)
    __shared__ uint4 tile[8 * 33];
    const int tiles_c = C / 64;
    const int64_t tiles_r = Rp / 32;
    const int t = threadIdx.x;
    for (int64_t b = blockIdx.x; b < tiles_r * tiles_c; b += gridDim.x) {
        const int tc = (int)(b % tiles_c);
        const int64_t r0 = (b / tiles_c) * 32;
        {
            const int row = t >> 3, cg = t & 7;
            const int64_t r = r0 + row;
            uint4 o = make_uint4(0, 0, 0, 0);
            if (r < R) {
                const float4 lo = *reinterpret_cast<const float4*>(x + r * ld + tc * 64 + cg * 8);
                const float4 hi = *reinterpret_cast<const float4*>(x + r * ld + tc * 64 + cg * 8 + 4);
                o = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
            }
            tile[cg * 33 + row] = o;
        }
        __syncthreads();
        {
            const int pl = t >> 5, row = t & 31;
            *reinterpret_cast<uint4*>(out + ((int64_t)(tc * 8 + pl) * Rp + r0 + row) * 16) = tile[pl * 33 + row];
        }
        __syncthreads();
    }
}

__global__ void pack_mplanes_kernel(const float* __restrict__ x, int64_t R, int C, int64_t ld, int64_t Rp, int Cp,
                                    uint8_t* __restrict__ out) {
    const int64_t total = (Rp / 8) * Cp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cp);
        const int64_t rp = i / Cp;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int64_t r = rp * 8 + j;
            v[j] = (r < R && c < C) ? __ldg(x + r * ld + c) : 0.f;
        }
        *reinterpret_cast<uint4*>(out + i * 16) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
}

// column sums of bf16 planes [P][Q][8] over the Q rows -> out[8p + c] (the decoder's bias gradient); one warp per plane
__global__ void __launch_bounds__(256) plane_colsum_kernel(const uint8_t* __restrict__ in, int64_t P, int64_t Q,
                                                           int64_t n_valid, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < P; p += nwarps) {
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int64_t q = lane; q < Q; q += 32) {
            const uint4 v = *reinterpret_cast<const uint4*>(in + (p * Q + q) * 16);
            const float2 a = unpack_bf16x2(v.x), b = unpack_bf16x2(v.y), c = unpack_bf16x2(v.z), d = unpack_bf16x2(v.w);
            s[0] += a.x; s[1] += a.y; s[2] += b.x; s[3] += b.y; s[4] += c.x; s[5] += c.y; s[6] += d.x; s[7] += d.y;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        }
        if (lane < 8 && p * 8 + lane < n_valid) {
            float v = s[0];
#pragma unroll
            for (int k = 1; k < 8; ++k) v = lane == k ? s[k] : v;
            out[p * 8 + lane] = v;
        }
    }
}

// out[i] = sum over s of part[s][i]   (K-split partial sums)
__global__ void sum_splits_kernel(const float* __restrict__ part, int nsplit, int64_t n4, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 acc = reinterpret_cast<const float4*>(part)[i];
        for (int s = 1; s < nsplit; ++s) {
            const float4 v = reinterpret_cast<const float4*>(part)[(int64_t)s * n4 + i];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
}

// out[c] = sum over rows of x[r][c]  (bias gradients, LayerNorm parameter gradients from per-block partials).
// One block per 32 columns, 32 row groups, fixed summation order.
__global__ void __launch_bounds__(1024) colsum_kernel(const float* __restrict__ x, int64_t R, int C, int64_t ld,
                                                      float* __restrict__ out) {
    __shared__ float s[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (c < C)
        for (int64_t r = threadIdx.y; r < R; r += 32) acc += x[r * ld + c];
    s[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float tot = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) tot += s[j][threadIdx.x];
        out[c] = tot;
    }
}

// ---------------------------------------------------------------------------------------------
// z = res + dropout(branch); y = LayerNorm(z)   (nn.TransformerEncoderLayer post-norm: x = norm(x + dropout(sa(x))))
// One warp per row.  stats[row] = (mean, rstd).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) res_drop_ln_kernel(const float* __restrict__ res, const float* __restrict__ br,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, int64_t M, int D, uint32_t site, Rng rng,
                                                          float* __restrict__ z, float2* __restrict__ stats,
                                                          float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const uint32_t step = rng_step(rng);
    const int D4 = D / 4;
    float sum = 0.f, sq = 0.f;
    for (int q = lane; q < D4; q += 32) {
        const float4 r4 = reinterpret_cast<const float4*>(res + row * D)[q];
        float4 b4 = reinterpret_cast<const float4*>(br + row * D)[q];
        if (rng.thresh) {
            const Philox4 k = drop_bits(rng, step, site, (uint32_t)(row * D4 + q));
            const float s = rng.keep_scale;
            b4.x = k.x >= rng.thresh ? b4.x * s : 0.f;
            b4.y = k.y >= rng.thresh ? b4.y * s : 0.f;
            b4.z = k.z >= rng.thresh ? b4.z * s : 0.f;
            b4.w = k.w >= rng.thresh ? b4.w * s : 0.f;
        }
        const float4 v = make_float4(r4.x + b4.x, r4.y + b4.y, r4.z + b4.z, r4.w + b4.w);
        reinterpret_cast<float4*>(z + row * D)[q] = v;
        sum += (v.x + v.y) + (v.z + v.w);
        sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, sq))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    const float mean = sum / (float)D;
    const float rstd = rsqrtf(fmaxf(sq / (float)D - mean * mean, 0.f) + eps);
    if (lane == 0) stats[row] = make_float2(mean, rstd);
    for (int q = lane; q < D4; q += 32) {
        const float4 v = reinterpret_cast<const float4*>(z + row * D)[q];   // this lane's own writes
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q);
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + q);
        reinterpret_cast<float4*>(y + row * D)[q] =
            make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y, (v.z - mean) * rstd * g.z + b.z,
                        (v.w - mean) * rstd * g.w + b.w);
    }
}

// LayerNorm backward + the split of dz into the residual path (dz itself) and the dropped branch (dz * mask / (1-p)).
// Per-block partial sums of dgamma / dbeta go to part[block][2][D]; colsum_kernel finishes them.  QPL = D / 128.
template <int QPL>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                     const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                     int64_t M, uint32_t site, Rng rng, float* __restrict__ dz,
                                                     float* __restrict__ dbr, float* __restrict__ part) {
    constexpr int D = QPL * 128, D4 = D / 4;
    __shared__ float4 s_g[D4], s_b[D4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t step = rng_step(rng);
    float4 ag[QPL], ab[QPL];
#pragma unroll
    for (int i = 0; i < QPL; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < M; row += (int64_t)gridDim.x * 8) {
        const float2 st = stats[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < QPL; ++i) {
            const int q = lane + 32 * i;
            const float4 d = reinterpret_cast<const float4*>(dy + row * D)[q];
            const float4 v = reinterpret_cast<const float4*>(z + row * D)[q];
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + q);
            const float4 xh = make_float4((v.x - st.x) * st.y, (v.y - st.x) * st.y, (v.z - st.x) * st.y, (v.w - st.x) * st.y);
            const float4 g4 = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
            ag[i].x = fmaf(d.x, xh.x, ag[i].x); ag[i].y = fmaf(d.y, xh.y, ag[i].y);
            ag[i].z = fmaf(d.z, xh.z, ag[i].z); ag[i].w = fmaf(d.w, xh.w, ag[i].w);
            ab[i].x += d.x; ab[i].y += d.y; ab[i].z += d.z; ab[i].w += d.w;
            s1 += (g4.x + g4.y) + (g4.z + g4.w);
            s2 = fmaf(g4.x, xh.x, fmaf(g4.y, xh.y, fmaf(g4.z, xh.z, fmaf(g4.w, xh.w, s2))));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float m1 = s1 / (float)D, m2 = s2 / (float)D;
        // second pass over the row (it is in L1): registers hold only the parameter-gradient accumulators
#pragma unroll
        for (int i = 0; i < QPL; ++i) {
            const int q = lane + 32 * i;
            const float4 d = reinterpret_cast<const float4*>(dy + row * D)[q];
            const float4 v = reinterpret_cast<const float4*>(z + row * D)[q];
            const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + q);
            float4 o;
            o.x = st.y * (d.x * gm.x - m1 - (v.x - st.x) * st.y * m2);
            o.y = st.y * (d.y * gm.y - m1 - (v.y - st.x) * st.y * m2);
            o.z = st.y * (d.z * gm.z - m1 - (v.z - st.x) * st.y * m2);
            o.w = st.y * (d.w * gm.w - m1 - (v.w - st.x) * st.y * m2);
            reinterpret_cast<float4*>(dz + row * D)[q] = o;
            if (rng.thresh) {
                const Philox4 k = drop_bits(rng, step, site, (uint32_t)(row * D4 + q));
                const float s = rng.keep_scale;
                o.x = k.x >= rng.thresh ? o.x * s : 0.f;
                o.y = k.y >= rng.thresh ? o.y * s : 0.f;
                o.z = k.z >= rng.thresh ? o.z * s : 0.f;
                o.w = k.w >= rng.thresh ? o.w * s : 0.f;
            }
            reinterpret_cast<float4*>(dbr + row * D)[q] = o;
        }
    }
    // block sums in warp order (fixed order: deterministic)
    for (int w = 0; w < 8; ++w) {
        if (warp == w) {
#pragma unroll
            for (int i = 0; i < QPL; ++i) {
                const int q = lane + 32 * i;
                if (w == 0) {
                    s_g[q] = ag[i];
                    s_b[q] = ab[i];
                } else {
                    float4 tg = s_g[q], tb = s_b[q];
                    tg.x += ag[i].x; tg.y += ag[i].y; tg.z += ag[i].z; tg.w += ag[i].w;
                    tb.x += ab[i].x; tb.y += ab[i].y; tb.z += ab[i].z; tb.w += ab[i].w;
                    s_g[q] = tg;
                    s_b[q] = tb;
                }
            }
        }
        __syncthreads();
    }
    for (int q = threadIdx.x; q < D4; q += 256) {
        reinterpret_cast<float4*>(part + (int64_t)blockIdx.x * 2 * D)[q] = s_g[q];
        reinterpret_cast<float4*>(part + (int64_t)blockIdx.x * 2 * D + D)[q] = s_b[q];
    }
}

// f = dropout(f) in place (f already passed through ReLU in the GEMM epilogue)
__global__ void drop_inplace_kernel(float* __restrict__ f, int64_t n4, uint32_t site, Rng rng) {
    if (rng.thresh == 0) return;
    const uint32_t step = rng_step(rng);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4*>(f)[i];
        const Philox4 k = drop_bits(rng, step, site, (uint32_t)i);
        const float s = rng.keep_scale;
        v.x = k.x >= rng.thresh ? v.x * s : 0.f;
        v.y = k.y >= rng.thresh ? v.y * s : 0.f;
        v.z = k.z >= rng.thresh ? v.z * s : 0.f;
        v.w = k.w >= rng.thresh ? v.w * s : 0.f;
        reinterpret_cast<float4*>(f)[i] = v;
    }
}

// d pre-activation = (f > 0) ? df / (1-p) : 0   (f = dropout(relu(pre)) is positive exactly where both let the value through)
__global__ void relu_drop_bwd_kernel(float* __restrict__ df, const float* __restrict__ f, int64_t n4, float keep_scale) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 d = reinterpret_cast<float4*>(df)[i];
        const float4 v = reinterpret_cast<const float4*>(f)[i];
        d.x = v.x > 0.f ? d.x * keep_scale : 0.f;
        d.y = v.y > 0.f ? d.y * keep_scale : 0.f;
        d.z = v.z > 0.f ? d.z * keep_scale : 0.f;
        d.w = v.w > 0.f ? d.w * keep_scale : 0.f;
        reinterpret_cast<float4*>(df)[i] = d;
    }
}

// diffusion loss (F.mse_loss, src/shakespeare.py:236): per-block partial sums of (pred - noise)^2 and, when training,
// d loss / d pred = 2 (pred - noise) / n
__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ noise, int64_t n4,
                                                  float inv_n, float* __restrict__ dpred, float* __restrict__ part) {
    __shared__ float s[8];
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = reinterpret_cast<const float4*>(pred)[i];
        const float4 q = reinterpret_cast<const float4*>(noise)[i];
        const float4 d = make_float4(p.x - q.x, p.y - q.y, p.z - q.z, p.w - q.w);
        acc = fmaf(d.x, d.x, fmaf(d.y, d.y, fmaf(d.z, d.z, fmaf(d.w, d.w, acc))));
        if (dpred) {
            const float c = 2.0f * inv_n;
            reinterpret_cast<float4*>(dpred)[i] = make_float4(c * d.x, c * d.y, c * d.z, c * d.w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s[w];
        part[blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// attention in train mode, one block per (sequence, head): softmax(q k^T / sqrt(hd)) -> P (kept for backward),
// dropout(P) v -> att.  fp32 SIMT: 4*L*D FLOP per token, ~2 % of an encoder layer.  The head dimension is walked in
// chunks of 64 so that any width fits: smem = S [L][L+1] + two [L][65] operand tiles.
// ---------------------------------------------------------------------------------------------
constexpr int kAttnThreads = 512;   // one block per (sequence, head): with 128 blocks on 148 SMs at the reference batch, warps per block are the only parallelism
template <int L>
__global__ void __launch_bounds__(kAttnThreads) attn_train_fwd_kernel(const float* __restrict__ qkv, int D, Rng rng, uint32_t site,
                                                             float* __restrict__ P, float* __restrict__ att) {
    extern __shared__ float sm[];
    float* sS = sm;                    // [L][L+1]
    float* sA = sS + L * (L + 1);      // [L][65]
    float* sB = sA + L * 65;
    const int HD = D / kTrHeads;
    const int64_t b = blockIdx.x / kTrHeads;
    const int h = blockIdx.x % kTrHeads;
    const float* base = qkv + b * L * (int64_t)(3 * D) + h * HD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < L * (L + 1); e += kAttnThreads) sS[e] = 0.f;
    for (int d0 = 0; d0 < HD; d0 += 64) {
        __syncthreads();
        for (int e = tid; e < L * 64; e += kAttnThreads) {
            const int i = e >> 6, d = e & 63;
            sA[i * 65 + d] = base[(int64_t)i * 3 * D + d0 + d];
            sB[i * 65 + d] = base[(int64_t)i * 3 * D + D + d0 + d];
        }
        __syncthreads();
        for (int e = tid; e < L * L; e += kAttnThreads) {
            const int i = e / L, j = e % L;
            float acc = 0.f;
#pragma unroll 16
            for (int d = 0; d < 64; ++d) acc = fmaf(sA[i * 65 + d], sB[j * 65 + d], acc);
            sS[i * (L + 1) + j] += acc;
        }
    }
    __syncthreads();
    const float scale = rsqrtf((float)HD);
    const uint32_t step = rng_step(rng);
    for (int i = warp; i < L; i += kAttnThreads / 32) {
        float v[L / 32];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < L / 32; ++k) {
            v[k] = sS[i * (L + 1) + lane + 32 * k] * scale;
            mx = fmaxf(mx, v[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < L / 32; ++k) {
            v[k] = __expf(v[k] - mx);
            sum += v[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float inv = 1.0f / sum;
#pragma unroll
        for (int k = 0; k < L / 32; ++k) {
            const int j = lane + 32 * k;
            const float p = v[k] * inv;
            const int64_t idx = ((int64_t)blockIdx.x * L + i) * L + j;
            P[idx] = p;
            sS[i * (L + 1) + j] = drop_keep1(rng, step, site, (uint64_t)idx) ? p * rng.keep_scale : 0.f;
        }
    }
    for (int d0 = 0; d0 < HD; d0 += 64) {
        __syncthreads();
        for (int e = tid; e < L * 64; e += kAttnThreads) {
            const int j = e >> 6, d = e & 63;
            sB[j * 65 + d] = base[(int64_t)j * 3 * D + 2 * D + d0 + d];
        }
        __syncthreads();
        for (int e = tid; e < L * 64; e += kAttnThreads) {
            const int i = e >> 6, d = e & 63;
            float acc = 0.f;
#pragma unroll 16
            for (int j = 0; j < L; ++j) acc = fmaf(sS[i * (L + 1) + j], sB[j * 65 + d], acc);
            att[(b * L + i) * (int64_t)D + h * HD + d0 + d] = acc;
        }
    }
}

// backward of the above: datt [M][D] -> dqkv [M][3D]
template <int L>
__global__ void __launch_bounds__(kAttnThreads) attn_train_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ P,
                                                             const float* __restrict__ datt, int D, Rng rng, uint32_t site,
                                                             float* __restrict__ dqkv) {
    extern __shared__ float sm[];
    float* sP = sm;                    // dropout(P)
    float* sS = sP + L * (L + 1);      // d dropout(P), then dS
    float* sA = sS + L * (L + 1);
    float* sB = sA + L * 65;
    const int HD = D / kTrHeads;
    const int64_t b = blockIdx.x / kTrHeads;
    const int h = blockIdx.x % kTrHeads;
    const int64_t ld = 3 * (int64_t)D;
    const float* base = qkv + b * L * ld + h * HD;
    float* dbase = dqkv + b * L * ld + h * HD;
    const float* dobase = datt + b * L * (int64_t)D + h * HD;
    const float* Pb = P + (int64_t)blockIdx.x * L * L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t step = rng_step(rng);
    for (int e = tid; e < L * L; e += kAttnThreads) {
        const int i = e / L, j = e % L;
        const float p = Pb[e];
        sP[i * (L + 1) + j] = drop_keep1(rng, step, site, (uint64_t)((int64_t)blockIdx.x * L * L + e)) ? p * rng.keep_scale : 0.f;
        sS[i * (L + 1) + j] = 0.f;
    }
    // phase A: d dropout(P) = dO v^T (accumulated over head-dimension chunks), dv = dropout(P)^T dO
    for (int d0 = 0; d0 < HD; d0 += 64) {
        __syncthreads();
        for (int e = tid; e < L * 64; e += kAttnThreads) {
            const int i = e >> 6, d = e & 63;
            sA[i * 65 + d] = dobase[(int64_t)i * D + d0 + d];
            sB[i * 65 + d] = base[(int64_t)i * ld + 2 * D + d0 + d];
        }
        __syncthreads();
        for (int e = tid; e < L * L; e += kAttnThreads) {
            const int i = e / L, j = e % L;
            float acc = 0.f;
#pragma unroll 16
            for (int d = 0; d < 64; ++d) acc = fmaf(sA[i * 65 + d], sB[j * 65 + d], acc);
            sS[i * (L + 1) + j] += acc;
        }
        for (int e = tid; e < L * 64; e += kAttnThreads) {
            const int j = e >> 6, d = e & 63;
            float acc = 0.f;
#pragma unroll 16
            for (int i = 0; i < L; ++i) acc = fmaf(sP[i * (L + 1) + j], sA[i * 65 + d], acc);
            dbase[(int64_t)j * ld + 2 * D + d0 + d] = acc;
        }
    }
    __syncthreads();
    // softmax backward per row, with the 1/sqrt(hd) of the scores folded in
    const float scale = rsqrtf((float)HD);
    for (int i = warp; i < L; i += kAttnThreads / 32) {
        float p[L / 32], dp[L / 32];
        float rs = 0.f;
#pragma unroll
        for (int k = 0; k < L / 32; ++k) {
            const int j = lane + 32 * k;
            p[k] = Pb[i * L + j];
            dp[k] = sP[i * (L + 1) + j] != 0.f ? sS[i * (L + 1) + j] * rng.keep_scale : 0.f;
            rs = fmaf(dp[k], p[k], rs);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
#pragma unroll
        for (int k = 0; k < L / 32; ++k) sS[i * (L + 1) + lane + 32 * k] = p[k] * (dp[k] - rs) * scale;
    }
    // phase B: dq = dS k, dk = dS^T q
    for (int d0 = 0; d0 < HD; d0 += 64) {
        __syncthreads();
        for (int e = tid; e < L * 64; e += kAttnThreads) {
            const int i = e >> 6, d = e & 63;
            sA[i * 65 + d] = base[(int64_t)i * ld + D + d0 + d];   // k
            sB[i * 65 + d] = base[(int64_t)i * ld + d0 + d];       // q
        }
        __syncthreads();
        for (int e = tid; e < L * 64; e += kAttnThreads) {
            const int i = e >> 6, d = e & 63;
            float aq = 0.f, ak = 0.f;
#pragma unroll 8
            for (int j = 0; j < L; ++j) {
                aq = fmaf(sS[i * (L + 1) + j], sA[j * 65 + d], aq);
                ak = fmaf(sS[j * (L + 1) + i], sB[j * 65 + d], ak);
            }
            dbase[(int64_t)i * ld + d0 + d] = aq;
            dbase[(int64_t)i * ld + D + d0 + d] = ak;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// input side of the backward pass: dh0 -> d x_noisy (input dropout) -> d x0 (diffusion branch) = sqrt_acp[t] d x_noisy,
// and the per-sequence sums that give the time embedding's gradient.  One block per sequence.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) input_bwd_kernel(const float* __restrict__ dh0, const int64_t* __restrict__ t,
                                                        const float* __restrict__ sqrt_acp, int L, int D, Rng rng,
                                                        float* __restrict__ dx0, float* __restrict__ dtb) {
    const int64_t b = blockIdx.x;
    const float ca = __ldg(sqrt_acp + __ldg(t + b));
    const uint32_t step = rng_step(rng);
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float acc = 0.f;
        for (int l = 0; l < L; ++l) {
            const int64_t idx = (b * L + l) * (int64_t)D + d;
            float g = dh0[idx];
            g = drop_keep1(rng, step, kSiteInput, (uint64_t)idx) ? g * rng.keep_scale : 0.f;
            acc += g;
            dx0[idx] = ca * g;
        }
        dtb[b * D + d] = acc;
    }
}

// time_emb = nn.Linear(1, D) on t/T: dW[d] = sum_b dtb[b][d] * t_b / T, db[d] = sum_b dtb[b][d]
__global__ void time_grad_kernel(const float* __restrict__ dtb, const int64_t* __restrict__ t, int64_t B, int D,
                                 float* __restrict__ dw, float* __restrict__ db) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    float aw = 0.f, ab = 0.f;
    for (int64_t b = 0; b < B; ++b) {
        const float g = dtb[b * D + d];
        aw = fmaf(g, (float)__ldg(t + b) / 1000.0f, aw);
        ab += g;
    }
    dw[d] = aw;
    db[d] = ab;
}

// embedding gradient: grad[ids[m]] += d x0 (diffusion branch) + d x0 (rounding branch); grad is zeroed beforehand
__global__ void __launch_bounds__(64) embed_scatter_kernel(const int64_t* __restrict__ ids, const float* __restrict__ da,
                                                           const float* __restrict__ db, int64_t V, int D,
                                                           float* __restrict__ grad) {
    const int64_t m = blockIdx.x;
    const int64_t id = __ldg(ids + m);
    if (id < 0 || id >= V) return;
    for (int q = threadIdx.x; q < D / 4; q += blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(da + m * D)[q];
        const float4 b = reinterpret_cast<const float4*>(db + m * D)[q];
        atomicAdd(reinterpret_cast<float4*>(grad + id * D) + q, make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w));
    }
}

// log-sum-exp per row from the per-(split, group) partial (max, sum) pairs, and the row's loss lse - logit[target]
__global__ void lse_merge_kernel(const float* __restrict__ pm, const float* __restrict__ ps, int nparts, int64_t M, int64_t Mp,
                                 const float* __restrict__ tgt_logit, float* __restrict__ lse, float* __restrict__ row_loss) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    float mx = -INFINITY;
    for (int p = 0; p < nparts; ++p) mx = fmaxf(mx, pm[(int64_t)p * Mp + row]);
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) {
        const float m = pm[(int64_t)p * Mp + row];
        if (m > -INFINITY) s += ps[(int64_t)p * Mp + row] * expf(m - mx);
    }
    const float l = mx + logf(s);
    lse[row] = l;
    row_loss[row] = l - tgt_logit[row];
}

// losses[0] = diffusion, [1] = rounding, [2] = diffusion + w * rounding   (src/shakespeare.py:236-244); one block
__global__ void __launch_bounds__(256) finish_losses_kernel(const float* __restrict__ mse_part, int n_mse, float inv_n,
                                                            const float* __restrict__ row_loss, int64_t M,
                                                            const float* __restrict__ rw_dev, float* __restrict__ losses) {
    __shared__ double s[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_mse; i += 256) acc += (double)mse_part[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    const double diff = s[0] * (double)inv_n;
    __syncthreads();
    acc = 0.0;
    for (int64_t i = threadIdx.x; i < M; i += 256) acc += (double)row_loss[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double rnd = s[0] / (double)M;
        losses[0] = (float)diff;
        losses[1] = (float)rnd;
        losses[2] = (float)(diff + (double)__ldg(rw_dev) * rnd);
    }
}

// AdamW (torch.optim.AdamW defaults apart from lr / weight decay; src/shakespeare.py:196) over a flat fp32 buffer, four
// elements per thread, learning rate read from the device (the cosine / warm-up schedule changes it every step while the
// step itself is a replayed CUDA graph).  Same arithmetic as adamw_kernel in unet_bwd.cu.
__global__ void __launch_bounds__(256) adamw_vec_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, int64_t n, const float* __restrict__ lr_dev,
                                                        double b1d, double b2d, float eps, float wd, float grad_scale,
                                                        const int64_t* __restrict__ step) {
    // torch passes beta and 1 - beta as double scalars rounded to fp32 once (1.0f - 0.999f is 4.7e-5 off 0.001f)
    const float b2 = (float)b2d, omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);
    __shared__ float s_c[2];
    const float lr = __ldg(lr_dev);
    if (threadIdx.x == 0) {
        const double k = (double)*step;
        s_c[0] = (float)((double)lr / (1.0 - pow(b1d, k)));
        s_c[1] = (float)sqrt(1.0 - pow(b2d, k));
    }
    __syncthreads();
    const float step_size = s_c[0], sbc2 = s_c[1];
    const float decay = 1.0f - lr * wd;
    auto upd = [&](float& pi, float gi, float& mi, float& vi) {
        gi *= grad_scale;
        float pn = pi * decay;
        mi = mi + (gi - mi) * omb1;
        vi = b2 * vi + omb2 * gi * gi;
        const float denom = sqrtf(vi) / sbc2 + eps;
        pn -= step_size * (mi / denom);
        pi = pn;
    };
    const int64_t n4 = n / 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 P = reinterpret_cast<float4*>(p)[i];
        const float4 G = __ldcs(reinterpret_cast<const float4*>(g) + i);
        float4 Mv = reinterpret_cast<float4*>(m)[i];
        float4 Vv = reinterpret_cast<float4*>(v)[i];
        upd(P.x, G.x, Mv.x, Vv.x);
        upd(P.y, G.y, Mv.y, Vv.y);
        upd(P.z, G.z, Mv.z, Vv.z);
        upd(P.w, G.w, Mv.w, Vv.w);
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = Mv;
        reinterpret_cast<float4*>(v)[i] = Vv;
    }
    if (blockIdx.x == 0 && (int64_t)threadIdx.x < n - n4 * 4) {
        const int64_t i = n4 * 4 + threadIdx.x;
        upd(p[i], g[i], m[i], v[i]);
    }
}

// ---------------------------------------------------------------------------------------------
// host side: parameter table, packed weights, workspace
// ---------------------------------------------------------------------------------------------
// offsets (in floats) of every parameter inside the flat buffer, in this order: 12 per encoder layer (the order of
// tdm_text_forward's pointer table), then time_emb.weight, time_emb.bias, decoder.weight, decoder.bias, embeddings.weight
enum : int { TP_WQKV = 0, TP_BQKV, TP_WO, TP_BO, TP_W1, TP_B1, TP_W2, TP_B2, TP_G1, TP_BE1, TP_G2, TP_BE2, TP_LAYER };
enum : int { TP_TW = 0, TP_TB, TP_DECW, TP_DECB, TP_EMB, TP_TAIL };

static inline int64_t up256(int64_t v) { return (v + 255) / 256 * 256; }

struct TrainPack {   // byte offsets of the bf16 operand forms of every weight matrix
    int64_t k[4], t[4];   // per layer, relative to the layer's base: qkv, out, ffn1, ffn2 as k-planes / transposed
    int64_t layer_bytes;
    int64_t dec_k, dec_t;
    int64_t total;
};
static TrainPack make_train_pack(int D, int depth, int64_t V) {
    TrainPack p{};
    const int64_t n[4] = {3 * (int64_t)D, D, kTrFF, D};
    const int64_t kk[4] = {D, D, D, kTrFF};
    int64_t o = 0;
    for (int i = 0; i < 4; ++i) {
        p.k[i] = o; o += up256(n[i] * kk[i] * 2);
        p.t[i] = o; o += up256(n[i] * kk[i] * 2);
    }
    p.layer_bytes = o;
    o *= depth;
    const int64_t Vp = up256(V);
    p.dec_k = o; o += up256(Vp * D * 2);
    p.dec_t = o; o += up256(Vp * D * 2);
    p.total = o;
    return p;
}

struct TrainWs {
    int64_t M, Mp, Vp;
    int nsplit, n_mse, n_lnblk, ksplit_dx;
    int64_t split_floats;
    // fp32 rows
    int64_t t, x0, noise, hin /* depth+1 of them */, qkv, P, att, a, z1, st1, h1, f, g, z2, st2;   // last 11: per layer, + layer_stride
    int64_t layer_stride;
    int64_t dA, dB, dC, dbig, dqkv, dx0a, dx0b, dtb, lnpart, msepart, pm, ps, tgtl, lse, rowl, split;
    // bf16 planes
    int64_t pk, pm2, pk2, dl, dlT;
    int64_t bad;
    int64_t total;
};

static int choose_nsplit(int m_tiles, int n_tiles) {
    const int sms = num_sms();
    int nsplit = sms / m_tiles;
    if (nsplit < 1) {
        double best = 0.0;
        nsplit = 1;
        for (int s = 1; s <= 16 && s <= n_tiles; ++s) {
            const int64_t items = (int64_t)m_tiles * s;
            const double eff = (double)items / (double)(((items + sms - 1) / sms) * sms);
            if (eff > best + 1e-9) { best = eff; nsplit = s; }
        }
    }
    if (nsplit > n_tiles) nsplit = n_tiles;
    if (nsplit > 148) nsplit = 148;
    return nsplit;
}

static TrainWs make_train_ws(int64_t B, int L, int D, int depth, int64_t V) {
    TrainWs w{};
    w.M = B * L;
    w.Mp = up256(w.M);
    w.Vp = up256(V);
    const int m_tiles = (int)(w.Mp / kBM), n_tiles = (int)(w.Vp / kBN);
    w.nsplit = choose_nsplit(m_tiles, n_tiles);
    w.n_mse = 296;
    w.n_lnblk = (int)((w.M + 7) / 8 < 592 ? (w.M + 7) / 8 : 592);
    {   // dX0 = dlogits . W: few output tiles, reduction over the vocabulary -> split K over the idle SMs
        const int tiles = m_tiles * (D / kBN);
        int ks = num_sms() / tiles;
        const int kblocks = (int)(w.Vp / kBK);
        if (ks < 1) ks = 1;
        if (ks > 32) ks = 32;
        if (ks > kblocks) ks = kblocks;
        w.ksplit_dx = ks;
    }
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        const int64_t at = o;
        o += up256(bytes);
        return at;
    };
    const int64_t MD = w.M * D * 4;
    w.t = take(B * 8);
    w.x0 = take(MD);
    w.noise = take(MD);
    w.hin = take((int64_t)(depth + 1) * up256(MD));
    const int64_t l0 = o;
    w.qkv = take(3 * MD);
    w.P = take(B * kTrHeads * (int64_t)L * L * 4);
    w.att = take(MD);
    w.a = take(MD);
    w.z1 = take(MD);
    w.st1 = take(w.M * 8);
    w.h1 = take(MD);
    w.f = take(w.M * kTrFF * 4);
    w.g = take(MD);
    w.z2 = take(MD);
    w.st2 = take(w.M * 8);
    w.layer_stride = o - l0;
    o = l0 + w.layer_stride * depth;
    w.dA = take(MD);
    w.dB = take(MD);
    w.dC = take(MD);
    w.dbig = take(w.M * kTrFF * 4);
    w.dqkv = take(3 * MD);
    w.dx0a = take(MD);
    w.dx0b = take(MD);
    w.dtb = take(B * D * 4);
    w.lnpart = take((int64_t)w.n_lnblk * 2 * D * 4);
    w.msepart = take(w.n_mse * 4);
    w.pm = take((int64_t)2 * w.nsplit * w.Mp * 4);
    w.ps = take((int64_t)2 * w.nsplit * w.Mp * 4);
    w.tgtl = take(w.M * 4);
    w.lse = take(w.M * 4);
    w.rowl = take(w.M * 4);
    w.split_floats = (int64_t)w.ksplit_dx * w.M * D;
    if (w.split_floats < (int64_t)16 * kTrFF * D) w.split_floats = (int64_t)16 * kTrFF * D;
    w.split = take(w.split_floats * 4);
    const int64_t widest = 3 * (int64_t)D > kTrFF ? 3 * (int64_t)D : kTrFF;
    w.pk = take(widest * w.Mp * 2);    // k-planes of an activation / gradient [C/8][Mp][8]
    w.pm2 = take(w.Mp * widest * 2);   // m-planes [Mp/8][C][8]
    w.pk2 = take(w.Mp * widest * 2);   // second m-planes operand
    w.dl = take(w.Vp * w.Mp * 2);
    w.dlT = take(w.Vp * w.Mp * 2);
    w.bad = take(256);
    w.total = o;
    return w;
}

static int check_train_shape(int64_t B, int L, int D, int depth, int64_t V, const char* who) {
    TDM_CHECK_ARG(B > 0 && depth > 0 && depth <= 64 && V > 0, "%s: bad batch / depth / vocabulary", who);
    TDM_CHECK_ARG(L == 64 || L == 128, "%s: seq_len must be 64 or 128 (got %d)", who, L);
    TDM_CHECK_ARG(D == 256 || D == 512 || D == 1024 || D == 2048, "%s: model width must be 256, 512, 1024 or 2048 (got %d)", who, D);
    TDM_CHECK_ARG(B * L <= (1 << 20) && up256(V) * up256(B * L) * 2 <= ((int64_t)24 << 30),
                  "%s: batch too large (the bf16 d logits planes would exceed 24 GB)", who);
    return TDM_OK;
}

static unsigned ew_grid(int64_t n, int threads) {
    const int64_t want = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * 16;
    return (unsigned)(want < cap ? (want > 0 ? want : 1) : cap);
}

static int pack_k(const float* x, int64_t R, int C, int64_t ld, int64_t Rp, uint8_t* out, cudaStream_t st) {
    TDM_CHECK_ARG(C % 64 == 0 && Rp % 32 == 0, "pack_kplanes: columns must be a multiple of 64 and padded rows of 32");
    const int64_t blocks = (Rp / 32) * (C / 64);
    const int64_t cap = (int64_t)num_sms() * 16;
    pack_kplanes_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(x, R, C, ld, Rp, out);
    TDM_CHECK_LAUNCH("pack_kplanes");
    return TDM_OK;
}
static int pack_m(const float* x, int64_t R, int C, int64_t ld, int64_t Rp, int Cp, uint8_t* out, cudaStream_t st) {
    pack_mplanes_kernel<<<ew_grid((Rp / 8) * Cp, 256), 256, 0, st>>>(x, R, C, ld, Rp, Cp, out);
    TDM_CHECK_LAUNCH("pack_mplanes");
    return TDM_OK;
}
static int colsum(const float* x, int64_t R, int C, int64_t ld, float* out, cudaStream_t st) {
    colsum_kernel<<<(C + 31) / 32, dim3(32, 32), 0, st>>>(x, R, C, ld, out);
    TDM_CHECK_LAUNCH("colsum");
    return TDM_OK;
}

// out[M][N_valid] (fp32 rows, leading dimension ld) = A . W^T (+ bias) (+ add) (ReLU); A: bf16 planes along K with a_rows
// rows per plane, W: bf16 planes along K with w_rows rows per plane
struct SplitBuf {
    float* p;
    int64_t cap;   // floats
};

static int gemm_rows(const uint8_t* a_planes, int64_t a_rows, const uint8_t* w_planes, int64_t w_rows, const float* bias,
                     int64_t M, int64_t n_valid, int64_t K, float* out, int64_t ld, const float* add, int relu, int ksplit,
                     SplitBuf split, cudaStream_t st, const char* name) {
    float* split_buf = split.p;
    GemmArgs g{};
    g.a = a_planes; g.a_ps = a_rows * 16; g.w = w_planes; g.w_ps = w_rows * 16; g.bias = bias;
    g.M = (int)M; g.Mp = (int)a_rows; g.N = (int)w_rows; g.n_valid = (int)n_valid; g.K = (int)K; g.relu = relu;
    g.nsplit = (int)(w_rows / kBN);
    g.logits_ld = ld; g.logits_add = add;
    if (ksplit < 0) {
        // automatic: a long reduction on a handful of output tiles (the dW GEMMs: K = token rows; linear2 and its dX:
        // K = 2048) leaves most SMs idle - split K until the items fill the machine, at least 4 K blocks per item,
        // at most as many ways as the partial-sum buffer holds
        const int items = (int)(a_rows / kBM) * g.nsplit;
        const int kblocks = (int)(K / kBK);
        int ks = num_sms() / items;
        if (ks > kblocks / 4) ks = kblocks / 4;
        if (ks > 32) ks = 32;
        if (split.p == nullptr || M * ld <= 0) ks = 1;
        else if (ks > split.cap / (M * ld)) ks = (int)(split.cap / (M * ld));
        ksplit = (relu || ld != n_valid || ks < 2) ? 1 : ks;
    }
    if (ksplit > 1) {
        TDM_CHECK_ARG(!relu && split_buf && ld == n_valid, "%s: bad K-split arguments", name);
        g.ksplit = ksplit; g.logits = split_buf; g.split_stride = M * ld;
        int rc;
        if ((rc = launch_gemm<GE_LOGITS>(g, st, name))) return rc;
        sum_splits_kernel<<<ew_grid(M * ld / 4, 256), 256, 0, st>>>(split_buf, ksplit, M * ld / 4, out);
        TDM_CHECK_LAUNCH("sum_splits");
        return TDM_OK;
    }
    g.logits = out;
    return launch_gemm<GE_LOGITS>(g, st, name);
}

template <int L>
static int launch_attn_train(bool bwd, const float* qkv, const float* P, const float* datt, int D, const Rng& rng,
                             uint32_t site, float* Pout, float* out, int64_t B, cudaStream_t st) {
    if (!bwd) {
        constexpr int smem = (L * (L + 1) + 2 * L * 65) * 4;
        auto kern = attn_train_fwd_kernel<L>;
        TDM_SET_MAX_DYN_SMEM(kern, smem);
        kern<<<(unsigned)(B * kTrHeads), kAttnThreads, smem, st>>>(qkv, D, rng, site, Pout, out);
        TDM_CHECK_LAUNCH("attn_train_fwd");
    } else {
        constexpr int smem = (2 * L * (L + 1) + 2 * L * 65) * 4;
        auto kern = attn_train_bwd_kernel<L>;
        TDM_SET_MAX_DYN_SMEM(kern, smem);
        kern<<<(unsigned)(B * kTrHeads), kAttnThreads, smem, st>>>(qkv, P, datt, D, rng, site, out);
        TDM_CHECK_LAUNCH("attn_train_bwd");
    }
    return TDM_OK;
}

static int launch_ln_bwd(const float* dy, const float* z, const float2* stats, const float* gamma, int64_t M, int D,
                         uint32_t site, const Rng& rng, float* dz, float* dbr, float* part, int nblk, cudaStream_t st) {
    switch (D) {
        case 256: ln_bwd_kernel<2><<<nblk, 256, 0, st>>>(dy, z, stats, gamma, M, site, rng, dz, dbr, part); break;
        case 512: ln_bwd_kernel<4><<<nblk, 256, 0, st>>>(dy, z, stats, gamma, M, site, rng, dz, dbr, part); break;
        case 1024: ln_bwd_kernel<8><<<nblk, 256, 0, st>>>(dy, z, stats, gamma, M, site, rng, dz, dbr, part); break;
        default: ln_bwd_kernel<16><<<nblk, 256, 0, st>>>(dy, z, stats, gamma, M, site, rng, dz, dbr, part); break;
    }
    TDM_CHECK_LAUNCH("ln_bwd");
    return TDM_OK;
}

}  // namespace tdm

using namespace tdm;

extern "C" int64_t tdm_text_train_workspace_bytes(int64_t batch, int seq_len, int dim, int depth, int64_t vocab) {
    if (check_train_shape(batch, seq_len, dim, depth, vocab, "tdm_text_train_workspace_bytes")) return 0;
    return make_train_ws(batch, seq_len, dim, depth, vocab).total;
}

extern "C" int64_t tdm_text_train_wpack_bytes(int dim, int depth, int64_t vocab) {
    if (dim <= 0 || dim % 256 || depth <= 0 || vocab <= 0) return 0;
    return make_train_pack(dim, depth, vocab).total;
}

extern "C" int tdm_text_train_debug_layout(int64_t batch, int seq_len, int dim, int depth, int64_t vocab, int64_t* out) {
    int rc;
    if ((rc = check_train_shape(batch, seq_len, dim, depth, vocab, "tdm_text_train_debug_layout"))) return rc;
    TDM_CHECK_ARG(out, "tdm_text_train_debug_layout: null pointer");
    const TrainWs w = make_train_ws(batch, seq_len, dim, depth, vocab);
    const int64_t v[] = {w.total, w.t, w.x0, w.noise, w.hin, up256(w.M * dim * 4), w.qkv, w.P, w.att, w.a, w.z1, w.h1, w.f,
                         w.g, w.z2, w.layer_stride, w.dx0a, w.dx0b, w.lse, w.rowl, w.dl, w.dlT, w.Mp, w.Vp, w.bad};
    for (size_t i = 0; i < sizeof(v) / sizeof(v[0]); ++i) out[i] = v[i];
    return TDM_OK;
}

// bf16 operand forms of every weight matrix from the flat fp32 parameters (after each optimiser step)
extern "C" int tdm_text_train_pack(const float* flat, const int64_t* offsets, int dim, int depth, int64_t vocab,
                                   void* wpack, int64_t wpack_bytes, void* stream) {
    TDM_CHECK_ARG(flat && offsets && wpack && dim > 0 && dim % 256 == 0 && depth > 0 && vocab > 0,
                  "tdm_text_train_pack: bad arguments");
    const TrainPack p = make_train_pack(dim, depth, vocab);
    TDM_CHECK_ARG(wpack_bytes >= p.total, "tdm_text_train_pack: packed-weight buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* wp = reinterpret_cast<uint8_t*>(wpack);
    const int D = dim;
    const int64_t n[4] = {3 * (int64_t)D, D, kTrFF, D};
    const int kk[4] = {D, D, D, kTrFF};
    const int which[4] = {TP_WQKV, TP_WO, TP_W1, TP_W2};
    int rc;
    for (int li = 0; li < depth; ++li) {
        uint8_t* base = wp + (int64_t)li * p.layer_bytes;
        for (int i = 0; i < 4; ++i) {
            const float* w = flat + offsets[li * TP_LAYER + which[i]];
            if ((rc = pack_k(w, n[i], kk[i], kk[i], n[i], base + p.k[i], st))) return rc;
            if ((rc = pack_m(w, n[i], kk[i], kk[i], n[i], kk[i], base + p.t[i], st))) return rc;
        }
    }
    const float* dw = flat + offsets[depth * TP_LAYER + TP_DECW];
    const int64_t Vp = up256(vocab);
    if ((rc = pack_k(dw, vocab, D, D, Vp, wp + p.dec_k, st))) return rc;
    if ((rc = pack_m(dw, vocab, D, D, Vp, D, wp + p.dec_t, st))) return rc;
    return TDM_OK;
}

extern "C" int tdm_adamw_flat_lr(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 const float* lr_dev, double beta1, double beta2, float eps, float weight_decay,
                                 float grad_scale, const int64_t* step_dev, void* stream) {
    TDM_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && lr_dev && step_dev && n >= 0, "tdm_adamw_flat_lr: bad arguments");
    TDM_CHECK_ARG(((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(grads) | reinterpret_cast<uintptr_t>(exp_avg) |
                    reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15) == 0, "tdm_adamw_flat_lr: 16-byte alignment required");
    if (n == 0) return TDM_OK;
    adamw_vec_kernel<<<ew_grid(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, lr_dev, beta1,
                                                                                beta2, eps, weight_decay, grad_scale, step_dev);
    TDM_CHECK_LAUNCH("tdm_adamw_flat_lr");
    return TDM_OK;
}

// One forward (+ backward when grads != NULL) of the training objective.  See include/tdm_b200.h.
extern "C" int tdm_text_train_step(const float* flat, float* grads, const int64_t* offsets, const void* wpack,
                                   const float* emb_table, const int64_t* token_ids, const int64_t* t_in,
                                   const float* noise_in, const float* sqrt_acp, const float* sqrt_om_acp,
                                   void* workspace, int64_t workspace_bytes, int64_t batch, int seq_len, int dim, int depth,
                                   int64_t vocab, float dropout_p, const float* rounding_weight_dev, uint64_t seed,
                                   uint64_t sample_offset, const int64_t* step_dev, float* losses, void* stream) {
    int rc;
    if ((rc = check_train_shape(batch, seq_len, dim, depth, vocab, "tdm_text_train_step"))) return rc;
    TDM_CHECK_ARG(flat && offsets && wpack && token_ids && sqrt_acp && sqrt_om_acp && workspace && rounding_weight_dev &&
                      step_dev && losses, "tdm_text_train_step: null pointer");
    TDM_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "tdm_text_train_step: dropout must be in [0, 1)");
    const bool train = grads != nullptr;
    const int D = dim, L = seq_len;
    const int64_t B = batch, V = vocab;
    const TrainWs W = make_train_ws(B, L, D, depth, V);
    TDM_CHECK_ARG(workspace_bytes >= W.total, "tdm_text_train_step: workspace too small");
    const TrainPack PK = make_train_pack(D, depth, V);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    const uint8_t* wp = reinterpret_cast<const uint8_t*>(wpack);
    auto F = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
    const int64_t M = W.M, Mp = W.Mp, Vp = W.Vp;
    const int64_t hin_stride = up256(M * D * 4);
    const int64_t* tail = offsets + depth * TP_LAYER;
    const bool learn_emb = tail[TP_EMB] >= 0;
    TDM_CHECK_ARG(learn_emb || emb_table, "tdm_text_train_step: no embedding table");
    const float* table = learn_emb ? flat + tail[TP_EMB] : emb_table;

    Rng rng{};
    rng.seed = seed;
    rng.step_dev = step_dev;
    const bool drop = train && dropout_p > 0.f;
    rng.thresh = drop ? (uint32_t)fmin((double)dropout_p * 4294967296.0, 4294967295.0) : 0u;
    rng.keep_scale = drop ? 1.0f / (1.0f - dropout_p) : 1.0f;

    int64_t* t = reinterpret_cast<int64_t*>(ws + W.t);
    if (t_in) {
        TDM_CHECK_CUDA(cudaMemcpyAsync(t, t_in, B * 8, cudaMemcpyDeviceToDevice, st));
    } else {
        draw_t_kernel<<<(unsigned)((B + 127) / 128), 128, 0, st>>>(t, B, 1000, rng, sample_offset);
        TDM_CHECK_LAUNCH("draw_t");
    }
    int* bad = reinterpret_cast<int*>(ws + W.bad);
    {
        EmbedArgs e{};
        e.table = table; e.ids = token_ids; e.t = t; e.sqrt_acp = sqrt_acp; e.sqrt_om = sqrt_om_acp; e.noise_in = noise_in;
        e.tw = flat + tail[TP_TW]; e.tb = flat + tail[TP_TB]; e.x0 = F(W.x0); e.noise = F(W.noise); e.h0 = F(W.hin);
        e.V = V; e.L = L; e.D = D; e.sample_offset = sample_offset; e.rng = rng; e.bad = bad;
        embed_noise_kernel<<<(unsigned)M, 64, 0, st>>>(e);
        TDM_CHECK_LAUNCH("embed_noise");
    }
    const SplitBuf SP{F(W.split), W.split_floats};
    uint8_t* pk = ws + W.pk;
    uint8_t* pm = ws + W.pm2;
    uint8_t* pk2 = ws + W.pk2;

    // ---------------- encoder forward ----------------
    for (int li = 0; li < depth; ++li) {
        const int64_t lo = (int64_t)li * W.layer_stride;
        const int64_t* po = offsets + li * TP_LAYER;
        const uint8_t* lw = wp + (int64_t)li * PK.layer_bytes;
        float* hin = reinterpret_cast<float*>(ws + W.hin + (int64_t)li * hin_stride);
        float* hout = reinterpret_cast<float*>(ws + W.hin + (int64_t)(li + 1) * hin_stride);
        const uint32_t site = 16u * (uint32_t)li;
        if ((rc = pack_k(hin, M, D, D, Mp, pk, st))) return rc;
        if ((rc = gemm_rows(pk, Mp, lw + PK.k[0], 3 * D, flat + po[TP_BQKV], M, 3 * D, D, F(W.qkv + lo), 3 * D, nullptr, 0, -1,
                            SP, st, "train_qkv"))) return rc;
        if (L == 64) rc = launch_attn_train<64>(false, F(W.qkv + lo), nullptr, nullptr, D, rng, site + 1, F(W.P + lo), F(W.att + lo), B, st);
        else rc = launch_attn_train<128>(false, F(W.qkv + lo), nullptr, nullptr, D, rng, site + 1, F(W.P + lo), F(W.att + lo), B, st);
        if (rc) return rc;
        if ((rc = pack_k(F(W.att + lo), M, D, D, Mp, pk, st))) return rc;
        if ((rc = gemm_rows(pk, Mp, lw + PK.k[1], D, flat + po[TP_BO], M, D, D, F(W.a + lo), D, nullptr, 0, -1, SP, st,
                            "train_out_proj"))) return rc;
        res_drop_ln_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(hin, F(W.a + lo), flat + po[TP_G1], flat + po[TP_BE1], 1e-5f, M,
                                                                  D, site + 2, rng, F(W.z1 + lo),
                                                                  reinterpret_cast<float2*>(ws + W.st1 + lo), F(W.h1 + lo));
        TDM_CHECK_LAUNCH("res_drop_ln1");
        if ((rc = pack_k(F(W.h1 + lo), M, D, D, Mp, pk, st))) return rc;
        if ((rc = gemm_rows(pk, Mp, lw + PK.k[2], kTrFF, flat + po[TP_B1], M, kTrFF, D, F(W.f + lo), kTrFF, nullptr, 1, 0,
                            SP, st, "train_ffn1"))) return rc;
        if (drop) {
            drop_inplace_kernel<<<ew_grid(M * kTrFF / 4, 256), 256, 0, st>>>(F(W.f + lo), M * kTrFF / 4, site + 3, rng);
            TDM_CHECK_LAUNCH("ffn_dropout");
        }
        if ((rc = pack_k(F(W.f + lo), M, kTrFF, kTrFF, Mp, pk, st))) return rc;
        if ((rc = gemm_rows(pk, Mp, lw + PK.k[3], D, flat + po[TP_B2], M, D, kTrFF, F(W.g + lo), D, nullptr, 0, -1, SP, st,
                            "train_ffn2"))) return rc;
        res_drop_ln_kernel<<<(unsigned)((M + 7) / 8), 256, 0, st>>>(F(W.h1 + lo), F(W.g + lo), flat + po[TP_G2], flat + po[TP_BE2],
                                                                  1e-5f, M, D, site + 4, rng, F(W.z2 + lo),
                                                                  reinterpret_cast<float2*>(ws + W.st2 + lo), hout);
        TDM_CHECK_LAUNCH("res_drop_ln2");
    }
    const float* pred = reinterpret_cast<const float*>(ws + W.hin + (int64_t)depth * hin_stride);
    const float inv_n = 1.0f / (float)(M * D);
    mse_kernel<<<W.n_mse, 256, 0, st>>>(pred, F(W.noise), M * D / 4, inv_n, train ? F(W.dA) : nullptr, F(W.msepart));
    TDM_CHECK_LAUNCH("mse");

    // ---------------- rounding head: log-sum-exp pass ----------------
    if ((rc = pack_k(F(W.x0), M, D, D, Mp, pk, st))) return rc;
    GemmArgs g{};
    g.a = pk; g.a_ps = Mp * 16; g.w = wp + PK.dec_k; g.w_ps = Vp * 16; g.bias = flat + tail[TP_DECB];
    g.M = (int)M; g.Mp = (int)Mp; g.N = (int)Vp; g.n_valid = (int)V; g.K = D; g.nsplit = W.nsplit;
    g.target = token_ids; g.part_val = F(W.pm); g.part_sum = F(W.ps); g.tgt_logit = F(W.tgtl);
    if ((rc = launch_gemm<GE_LSE>(g, st, "round_lse"))) return rc;
    lse_merge_kernel<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(F(W.pm), F(W.ps), 2 * W.nsplit, M, Mp, F(W.tgtl), F(W.lse), F(W.rowl));
    TDM_CHECK_LAUNCH("lse_merge");
    finish_losses_kernel<<<1, 256, 0, st>>>(F(W.msepart), W.n_mse, inv_n, F(W.rowl), M, rounding_weight_dev, losses);
    TDM_CHECK_LAUNCH("finish_losses");
    if (!train) return TDM_OK;

    // ---------------- rounding head: backward ----------------
    g.lse = F(W.lse); g.dev_scale = rounding_weight_dev; g.scale = 1.0f / (float)M;
    g.out_bf16 = ws + W.dl; g.ob_ps = Mp * 16; g.out_bf16_t = ws + W.dlT; g.obt_rows = Vp;
    if ((rc = launch_gemm<GE_DLOGITS>(g, st, "round_dlogits"))) return rc;
    plane_colsum_kernel<<<ew_grid(Vp / 8 * 32, 256), 256, 0, st>>>(ws + W.dl, Vp / 8, Mp, V, grads + tail[TP_DECB]);
    TDM_CHECK_LAUNCH("decoder_bias_grad");
    // d x0 (rounding) = dlogits . W   [M][D], reduction over the vocabulary
    if ((rc = gemm_rows(ws + W.dl, Mp, wp + PK.dec_t, D, nullptr, M, D, Vp, F(W.dx0b), D, nullptr, 0, W.ksplit_dx, SP, st,
                        "round_dx0"))) return rc;
    // d W = dlogits^T . x0   [V][D], reduction over the token rows (the transposed planes came out of the same epilogue)
    if ((rc = pack_m(F(W.x0), M, D, D, Mp, D, pm, st))) return rc;
    if ((rc = gemm_rows(ws + W.dlT, Vp, pm, D, nullptr, V, D, Mp, grads + tail[TP_DECW], D, nullptr, 0, -1, SP, st,
                        "round_dw"))) return rc;

    // ---------------- encoder backward ----------------
    float* dcur = F(W.dA);   // gradient w.r.t. the current layer's output
    float* dtmp = F(W.dB);
    float* dres = F(W.dC);
    for (int li = depth - 1; li >= 0; --li) {
        const int64_t lo = (int64_t)li * W.layer_stride;
        const int64_t* po = offsets + li * TP_LAYER;
        const uint8_t* lw = wp + (int64_t)li * PK.layer_bytes;
        const float* hin = reinterpret_cast<const float*>(ws + W.hin + (int64_t)li * hin_stride);
        const uint32_t site = 16u * (uint32_t)li;
        // LN2: dcur -> dres (= d h1 through the residual), dtmp (= d g through the dropout)
        if ((rc = launch_ln_bwd(dcur, F(W.z2 + lo), reinterpret_cast<const float2*>(ws + W.st2 + lo), flat + po[TP_G2], M, D, site + 4,
                                rng, dres, dtmp, F(W.lnpart), W.n_lnblk, st))) return rc;
        if ((rc = colsum(F(W.lnpart), W.n_lnblk, D, 2 * D, grads + po[TP_G2], st))) return rc;
        if ((rc = colsum(F(W.lnpart) + D, W.n_lnblk, D, 2 * D, grads + po[TP_BE2], st))) return rc;
        // linear2: d f = dg . W2, dW2 = dg^T f, db2
        if ((rc = colsum(dtmp, M, D, D, grads + po[TP_B2], st))) return rc;
        if ((rc = pack_k(dtmp, M, D, D, Mp, pk, st))) return rc;
        if ((rc = gemm_rows(pk, Mp, lw + PK.t[3], kTrFF, nullptr, M, kTrFF, D, F(W.dbig), kTrFF, nullptr, 0, -1, SP, st,
                            "train_dffn2_x"))) return rc;
        if ((rc = pack_m(dtmp, M, D, D, Mp, D, pm, st))) return rc;
        if ((rc = pack_m(F(W.f + lo), M, kTrFF, kTrFF, Mp, kTrFF, pk2, st))) return rc;
        if ((rc = gemm_rows(pm, D, pk2, kTrFF, nullptr, D, kTrFF, Mp, grads + po[TP_W2], kTrFF, nullptr, 0, -1, SP, st,
                            "train_dffn2_w"))) return rc;
        relu_drop_bwd_kernel<<<ew_grid(M * kTrFF / 4, 256), 256, 0, st>>>(F(W.dbig), F(W.f + lo), M * kTrFF / 4, rng.keep_scale);
        TDM_CHECK_LAUNCH("relu_drop_bwd");
        // linear1: d h1 += d pre . W1, dW1 = d pre^T h1, db1
        if ((rc = colsum(F(W.dbig), M, kTrFF, kTrFF, grads + po[TP_B1], st))) return rc;
        if ((rc = pack_k(F(W.dbig), M, kTrFF, kTrFF, Mp, pk, st))) return rc;
        if ((rc = gemm_rows(pk, Mp, lw + PK.t[2], D, nullptr, M, D, kTrFF, dcur, D, dres, 0, -1, SP, st, "train_dffn1_x"))) return rc;
        if ((rc = pack_m(F(W.dbig), M, kTrFF, kTrFF, Mp, kTrFF, pm, st))) return rc;
        if ((rc = pack_m(F(W.h1 + lo), M, D, D, Mp, D, pk2, st))) return rc;
        if ((rc = gemm_rows(pm, kTrFF, pk2, D, nullptr, kTrFF, D, Mp, grads + po[TP_W1], D, nullptr, 0, -1, SP, st,
                            "train_dffn1_w"))) return rc;
        // LN1: dcur (= d h1) -> dres (= d h_in through the residual), dtmp (= d a through the dropout)
        if ((rc = launch_ln_bwd(dcur, F(W.z1 + lo), reinterpret_cast<const float2*>(ws + W.st1 + lo), flat + po[TP_G1], M, D, site + 2,
                                rng, dres, dtmp, F(W.lnpart), W.n_lnblk, st))) return rc;
        if ((rc = colsum(F(W.lnpart), W.n_lnblk, D, 2 * D, grads + po[TP_G1], st))) return rc;
        if ((rc = colsum(F(W.lnpart) + D, W.n_lnblk, D, 2 * D, grads + po[TP_BE1], st))) return rc;
        // out_proj: d att = da . Wo, dWo = da^T att, dbo
        if ((rc = colsum(dtmp, M, D, D, grads + po[TP_BO], st))) return rc;
        if ((rc = pack_k(dtmp, M, D, D, Mp, pk, st))) return rc;
        if ((rc = gemm_rows(pk, Mp, lw + PK.t[1], D, nullptr, M, D, D, dcur, D, nullptr, 0, -1, SP, st, "train_dout_x"))) return rc;
        if ((rc = pack_m(dtmp, M, D, D, Mp, D, pm, st))) return rc;
        if ((rc = pack_m(F(W.att + lo), M, D, D, Mp, D, pk2, st))) return rc;
        if ((rc = gemm_rows(pm, D, pk2, D, nullptr, D, D, Mp, grads + po[TP_WO], D, nullptr, 0, -1, SP, st, "train_dout_w"))) return rc;
        // attention: dcur (= d att) -> dqkv
        if (L == 64) rc = launch_attn_train<64>(true, F(W.qkv + lo), F(W.P + lo), dcur, D, rng, site + 1, nullptr, F(W.dqkv), B, st);
        else rc = launch_attn_train<128>(true, F(W.qkv + lo), F(W.P + lo), dcur, D, rng, site + 1, nullptr, F(W.dqkv), B, st);
        if (rc) return rc;
        // in_proj: d h_in = dres + dqkv . Wqkv, dWqkv = dqkv^T h_in, dbqkv
        if ((rc = colsum(F(W.dqkv), M, 3 * D, 3 * D, grads + po[TP_BQKV], st))) return rc;
        if ((rc = pack_k(F(W.dqkv), M, 3 * D, 3 * D, Mp, pk, st))) return rc;
        if ((rc = gemm_rows(pk, Mp, lw + PK.t[0], D, nullptr, M, D, 3 * D, dcur, D, dres, 0, -1, SP, st, "train_dqkv_x"))) return rc;
        if ((rc = pack_m(F(W.dqkv), M, 3 * D, 3 * D, Mp, 3 * D, pm, st))) return rc;
        if ((rc = pack_m(hin, M, D, D, Mp, D, pk2, st))) return rc;
        if ((rc = gemm_rows(pm, 3 * D, pk2, D, nullptr, 3 * D, D, Mp, grads + po[TP_WQKV], D, nullptr, 0, -1, SP, st,
                            "train_dqkv_w"))) return rc;
    }
    // ---------------- input side ----------------
    input_bwd_kernel<<<(unsigned)B, 256, 0, st>>>(dcur, t, sqrt_acp, L, D, rng, F(W.dx0a), F(W.dtb));
    TDM_CHECK_LAUNCH("input_bwd");
    time_grad_kernel<<<(D + 127) / 128, 128, 0, st>>>(F(W.dtb), t, B, D, grads + tail[TP_TW], grads + tail[TP_TB]);
    TDM_CHECK_LAUNCH("time_grad");
    if (learn_emb) {
        TDM_CHECK_CUDA(cudaMemsetAsync(grads + tail[TP_EMB], 0, (size_t)V * D * 4, st));
        embed_scatter_kernel<<<(unsigned)M, 64, 0, st>>>(token_ids, F(W.dx0a), F(W.dx0b), V, D, grads + tail[TP_EMB]);
        TDM_CHECK_LAUNCH("embed_scatter");
    }
    return TDM_OK;
}
