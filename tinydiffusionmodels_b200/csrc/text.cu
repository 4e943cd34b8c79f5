// text.cu — the Shakespeare embedding-space sampler (src/shakespeare.py:105-120, 343-352,
// 382-401, 418-470): TinyTransformer forward, reverse step, rounding (learned / cosine / guided).
//
// One reverse step of a batch of B sequences (M = B*L token rows, model width D) is, per encoder
// layer (post-LN, ReLU FFN, 4 heads — nn.TransformerEncoderLayer defaults):
//     QKV GEMM -> attention (tcgen05 QK^T and PV, softmax in registers) -> out-proj GEMM (+residual)
//     -> LayerNorm -> FFN1 GEMM (+ReLU) -> FFN2 GEMM (+residual) -> LayerNorm
// and the last LayerNorm also applies the reverse-step update (in-kernel Philox noise) and adds
// the next timestep's time embedding, so the denoiser output never round-trips as a separate
// tensor.  The token state lives in "planes" between steps: fp32 [D/4][Mp][4] for the residual
// stream, bf16 [D/8][Mp][8] for GEMM operands.
#include "common.cuh"
#include "diffusion_math.cuh"
#include "ffn_tc.cuh"
#include "gemm_tc.cuh"
#include "tc05.cuh"

namespace tdm {

// ---------------------------------------------------------------------------------------------
// weight packing: row-major fp32 [N][K] -> bf16 planes [K/8][Np][8] (rows >= N zero-filled)
// ---------------------------------------------------------------------------------------------
__global__ void pack_linear_kernel(const float* __restrict__ w, int N, int K, int Np,
                                   __nv_bfloat16* __restrict__ out) {
    const int64_t total = (int64_t)(K / 8) * Np;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % Np);
        const int kp = (int)(i / Np);
        uint4 o = make_uint4(0, 0, 0, 0);
        if (n < N) {
            const float4 lo = *reinterpret_cast<const float4*>(w + (int64_t)n * K + kp * 8);
            const float4 hi = *reinterpret_cast<const float4*>(w + (int64_t)n * K + kp * 8 + 4);
            o = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y),
                           pack_bf16x2(hi.z, hi.w));
        }
        *reinterpret_cast<uint4*>(out + i * 8) = o;
    }
}

// linear1.weight (2048, 256) and linear2.weight (256, 2048) in the order the fused FFN kernel's ring consumes them
// (ffn_tc.cuh): per hidden chunk c of 128 units and K block kb of 64, eight 16-byte-row planes -
//   W1 stage (c, kb): [plane j][row r < 128][8]  = W1[128c + r][64kb + 8j .. +8]        16 KB
//   W2 stage (c, kb): [plane j][row r < 256][8]  = W2[r][128c + 64kb + 8j .. +8]        32 KB
__global__ void pack_ffn_weights_kernel(const float* __restrict__ w1, const float* __restrict__ w2,
                                        __nv_bfloat16* __restrict__ o1, __nv_bfloat16* __restrict__ o2) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one 16-byte row of either output
    constexpr int kRows1 = 16 * 4 * 8 * 128, kRows2 = 16 * 2 * 8 * 256;
    const float* src;
    __nv_bfloat16* dst;
    if (i < kRows1) {
        const int r = i % 128, j = (i / 128) % 8, kb = (i / 1024) % 4, c = i / 4096;
        src = w1 + (int64_t)(128 * c + r) * 256 + 64 * kb + 8 * j;
        dst = o1 + (int64_t)i * 8;
    } else if (i < kRows1 + kRows2) {
        const int t = i - kRows1;
        const int r = t % 256, j = (t / 256) % 8, kb = (t / 2048) % 2, c = t / 4096;
        src = w2 + (int64_t)r * 2048 + 128 * c + 64 * kb + 8 * j;
        dst = o2 + (int64_t)t * 8;
    } else {
        return;
    }
    const float4 lo = *reinterpret_cast<const float4*>(src);
    const float4 hi = *reinterpret_cast<const float4*>(src + 4);
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(lo.x, lo.y), pack_bf16x2(lo.z, lo.w), pack_bf16x2(hi.x, hi.y), pack_bf16x2(hi.z, hi.w));
}

// ---------------------------------------------------------------------------------------------
// rows <-> planes
// ---------------------------------------------------------------------------------------------
// x [M][D] fp32 row-major -> state fp32 planes; h0 = x + time_emb(t) -> fp32 + bf16 planes.
// Also usable as a plain converter (tw == null: no time bias; any output pointer may be null) and
// optionally emits 1/max(||x_row||, 1e-12) (F.normalize's denominator) for cosine rounding.
// block = 256 threads handling 32 rows; channels in chunks of 128 through smem.
__global__ void __launch_bounds__(256)
rows_to_planes_kernel(const float* __restrict__ x, int M, int Mp, int D, int L, const int64_t* __restrict__ t,
                      const float* __restrict__ tw, const float* __restrict__ tb, uint8_t* __restrict__ state,
                      uint8_t* __restrict__ h32, uint8_t* __restrict__ h16, float* __restrict__ rnorm) {
    __shared__ float s[32][129];
    __shared__ float s_ss[8][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * 32;
    const int64_t ps = (int64_t)Mp * 16;
    float ssq = 0.f;
    for (int c0 = 0; c0 < D; c0 += 128) {
        __syncthreads();
        for (int rr = warp; rr < 32; rr += 8) {
            const int row = r0 + rr;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = k * 32 + lane;
                s[rr][c] = (row < M) ? x[(int64_t)row * D + c0 + c] : 0.f;
            }
        }
        __syncthreads();
        const int row = r0 + lane;
        float ts = 0.f;
        if (tw && row < M) ts = (float)__ldg(t + row / L) / 1000.0f;
        // warp w handles fp32 planes 4w..4w+3 of this chunk (= bf16 planes 2w, 2w+1)
        float v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = s[lane][warp * 16 + k];
#pragma unroll
        for (int k = 0; k < 16; ++k) ssq = fmaf(v[k], v[k], ssq);
        if (row < Mp) {
            const int cb = c0 + warp * 16;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (state)
                    *reinterpret_cast<float4*>(state + (int64_t)(cb / 4 + p) * ps + (int64_t)row * 16) =
                        make_float4(v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]);
            }
            float h[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) h[k] = tw ? v[k] + fmaf(__ldg(tw + cb + k), ts, __ldg(tb + cb + k)) : v[k];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                if (h32)
                    *reinterpret_cast<float4*>(h32 + (int64_t)(cb / 4 + p) * ps + (int64_t)row * 16) =
                        make_float4(h[4 * p], h[4 * p + 1], h[4 * p + 2], h[4 * p + 3]);
            }
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                if (h16)
                    *reinterpret_cast<uint4*>(h16 + (int64_t)(cb / 8 + p) * ps + (int64_t)row * 16) =
                        make_uint4(pack_bf16x2(h[8 * p], h[8 * p + 1]), pack_bf16x2(h[8 * p + 2], h[8 * p + 3]),
                                   pack_bf16x2(h[8 * p + 4], h[8 * p + 5]), pack_bf16x2(h[8 * p + 6], h[8 * p + 7]));
            }
        }
    }
    if (rnorm) {
        s_ss[warp][lane] = ssq;
        __syncthreads();
        if (warp == 0) {
            float tot = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_ss[w][lane];
            const int row = r0 + lane;
            if (row < M) rnorm[row] = 1.0f / fmaxf(sqrtf(tot), 1e-12f);
        }
    }
}

// fp32 planes -> [M][D] row-major
__global__ void __launch_bounds__(256)
planes_to_rows_kernel(const uint8_t* __restrict__ src, int M, int Mp, int D, float* __restrict__ out) {
    __shared__ float s[32][129];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * 32;
    const int64_t ps = (int64_t)Mp * 16;
    for (int c0 = 0; c0 < D; c0 += 128) {
        __syncthreads();
        const int row = r0 + lane;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < Mp) v = *reinterpret_cast<const float4*>(src + (int64_t)((c0 + warp * 16) / 4 + p) * ps + (int64_t)row * 16);
            s[lane][warp * 16 + 4 * p + 0] = v.x;
            s[lane][warp * 16 + 4 * p + 1] = v.y;
            s[lane][warp * 16 + 4 * p + 2] = v.z;
            s[lane][warp * 16 + 4 * p + 3] = v.w;
        }
        __syncthreads();
        for (int rr = warp; rr < 32; rr += 8) {
            const int orow = r0 + rr;
            if (orow < M) {
#pragma unroll
                for (int k = 0; k < 4; ++k) out[(int64_t)orow * D + c0 + k * 32 + lane] = s[rr][k * 32 + lane];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// attention: one CTA per (sequence, head).  S = Q K^T (tcgen05, K-major planes), softmax with one
// thread per query row, P -> smem planes, O = P V (V read MN-major from its plane tile).
// ---------------------------------------------------------------------------------------------
template <int L, int HD>
struct AttnCfg {
    static constexpr int QKV_BYTES = L * HD * 2;
    static constexpr int P_BYTES = L * L * 2;
    static constexpr int SMEM = 3 * QKV_BYTES + P_BYTES + 256;
    static constexpr int NO = HD > 256 ? 256 : HD;       // columns of one PV MMA
    // O reuses the columns of S: the softmax has read all of S into registers (and the block has synchronised)
    // before the first PV MMA is issued.  Halving the TMEM footprint doubles the CTAs resident per SM, which is
    // what this latency-bound kernel (one short dependent chain per CTA) is limited by.
    static constexpr int TCOLS_RAW = L > NO ? L : NO;
    static constexpr int TCOLS = TCOLS_RAW <= 32 ? 32 : TCOLS_RAW <= 64 ? 64 : TCOLS_RAW <= 128 ? 128 : TCOLS_RAW <= 256 ? 256 : 512;
    static_assert(SMEM <= 227 * 1024, "attention tile does not fit in shared memory");
};

template <int L, int HD>
__global__ void __launch_bounds__(128) attn_tc_kernel(const uint8_t* __restrict__ qkv, int64_t ps, int D,
                                                      int heads, uint8_t* __restrict__ out, int64_t out_ps) {
    using C = AttnCfg<L, HD>;
    static_assert(L == 64 || L == 128, "sequence length 64 or 128");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + C::QKV_BYTES;
    uint8_t* sV = sK + C::QKV_BYTES;
    uint8_t* sP = sV + C::QKV_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + C::P_BYTES);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / heads, h = blockIdx.x - b * heads;

    if (threadIdx.x == 0) {
        mbar_init(bars + 0, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<C::TCOLS>(s_tmem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_wait();                 // PDL (common.cuh): nothing above touches global memory
    pdl_launch_dependents();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        if (lane == 0) mbar_arrive_expect_tx(bars + 0, 3 * C::QKV_BYTES);
        __syncwarp();
        constexpr int NPL = HD / 8;
        for (int i = lane; i < 3 * NPL; i += 32) {
            const int which = i / NPL, j = i - which * NPL;
            const int64_t plane = (int64_t)(which * D + h * HD) / 8 + j;
            bulk_g2s(smem + which * C::QKV_BYTES + j * (L * 16), qkv + plane * ps + (int64_t)b * L * 16, L * 16, bars + 0);
        }
    }
    // ---- S = Q K^T ----
    if (threadIdx.x == 0) {
        mbar_wait(bars + 0, 0);
        tc_fence_after_sync();
        constexpr uint32_t idesc = make_idesc_bf16(L, L);
        const uint32_t q_addr = smem_u32(sQ), k_addr = smem_u32(sK);
#pragma unroll 1
        for (int ks = 0; ks < HD / 16; ++ks) {
            const uint64_t ad = make_smem_desc(q_addr + (2 * ks) * (L * 16), L * 16, 128);
            const uint64_t bd = make_smem_desc(k_addr + (2 * ks) * (L * 16), L * 16, 128);
            umma_bf16(tmem_base, ad, bd, idesc, ks != 0);
        }
        umma_commit(bars + 1);
    }
    // ---- softmax: accumulator row r lives in TMEM lane r (L=128) or (r%16)+32*(r/16) (L=64) ----
    mbar_wait(bars + 1, 0);
    tc_fence_after_sync();
    const bool active = (L == 128) || lane < 16;
    const int row = (L == 128) ? warp * 32 + lane : warp * 16 + (lane & 15);
    {
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float sc[L];
#pragma unroll
        for (int c0 = 0; c0 < L; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) sc[c0 + k] = __uint_as_float(r[k]);
        }
        const float scale = rsqrtf((float)HD);
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < L; ++k) mx = fmaxf(mx, sc[k]);
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < L; ++k) {
            sc[k] = __expf((sc[k] - mx) * scale);
            sum += sc[k];
        }
        const float inv = 1.0f / sum;
        if (active) {
#pragma unroll
            for (int j = 0; j < L / 8; ++j) {
                const uint4 o = make_uint4(pack_bf16x2(sc[8 * j] * inv, sc[8 * j + 1] * inv),
                                           pack_bf16x2(sc[8 * j + 2] * inv, sc[8 * j + 3] * inv),
                                           pack_bf16x2(sc[8 * j + 4] * inv, sc[8 * j + 5] * inv),
                                           pack_bf16x2(sc[8 * j + 6] * inv, sc[8 * j + 7] * inv));
                *reinterpret_cast<uint4*>(sP + j * (L * 16) + row * 16) = o;
            }
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    // ---- O = P V, HD columns in chunks of <= 256 ----
    constexpr int NCHUNK = HD / C::NO;
#pragma unroll 1
    for (int ch = 0; ch < NCHUNK; ++ch) {
        if (threadIdx.x == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(L, C::NO, 0, 1);   // A K-major, B MN-major
            const uint32_t p_addr = smem_u32(sP), v_addr = smem_u32(sV);
#pragma unroll 1
            for (int ks = 0; ks < L / 16; ++ks) {
                const uint64_t ad = make_smem_desc(p_addr + (2 * ks) * (L * 16), L * 16, 128);
                const uint64_t bd = make_smem_desc(v_addr + (ch * C::NO / 8) * (L * 16) + ks * 256, 128, L * 16);
                umma_bf16(tmem_base, ad, bd, idesc, ks != 0);
            }
            umma_commit(bars + 2);
        }
        mbar_wait(bars + 2, ch & 1);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < C::NO; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            tmem_ld_wait();
            if (active) {
#pragma unroll
                for (int pj = 0; pj < 4; ++pj) {
                    const uint4 o = make_uint4(
                        pack_bf16x2(__uint_as_float(r[pj * 8 + 0]), __uint_as_float(r[pj * 8 + 1])),
                        pack_bf16x2(__uint_as_float(r[pj * 8 + 2]), __uint_as_float(r[pj * 8 + 3])),
                        pack_bf16x2(__uint_as_float(r[pj * 8 + 4]), __uint_as_float(r[pj * 8 + 5])),
                        pack_bf16x2(__uint_as_float(r[pj * 8 + 6]), __uint_as_float(r[pj * 8 + 7])));
                    const int64_t plane = (int64_t)(h * HD + ch * C::NO + c0) / 8 + pj;
                    *reinterpret_cast<uint4*>(out + plane * out_ps + ((int64_t)b * L + row) * 16) = o;
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();   // all TMEM reads of this chunk done before the next chunk's MMAs overwrite it
        tc_fence_after_sync();
    }
    if (warp == 1) tmem_dealloc<C::TCOLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over fp32 planes (+ optionally the fused reverse step and next time embedding)
// block = 256 threads = 32 rows x 8 column slices.
// ---------------------------------------------------------------------------------------------
struct LnArgs {
    const uint8_t* in;       // fp32 planes, pre-LN
    const float* gamma;
    const float* beta;
    float eps;
    uint8_t* out32;          // post-LN fp32 planes (nullable)
    uint8_t* out16;          // post-LN bf16 planes (nullable)
    int M, Mp, D, L;
    // fused reverse step (final LN of the denoiser): y = LN output = eps_hat
    int fuse_step;
    uint8_t* state;          // fp32 planes of x_t, updated in place to x_{t-1}
    const int64_t* t;        // [B]
    const float* z;          // injected noise [M][D] row-major, or null -> Philox
    const float* betas;
    const float* alphas;
    const float* sqrt_om;
    const float* tw;         // time_emb of the NEXT step: h0 = x_{t-1} + tw*(t-1)/1000 + tb
    const float* tb;
    uint64_t seed, sample_offset;
    uint32_t step_id;
};

__global__ void __launch_bounds__(256) layernorm_kernel(const LnArgs a) {
    pdl_wait();   // PDL (common.cuh): first statement, nothing before it touches global memory
    pdl_launch_dependents();
    __shared__ float s_sum[8][32], s_sq[8][32];
    const int slice = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 32 + lane;
    const int64_t ps = (int64_t)a.Mp * 16;
    const int npairs = a.D / 8;   // pairs of fp32 planes == bf16 planes
    const bool rv = row < a.M;
    float sum = 0.f, sq = 0.f;
    if (rv) {
        for (int pp = slice; pp < npairs; pp += 8) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(a.in + (int64_t)(2 * pp + h) * ps + (int64_t)row * 16);
                sum += v.x + v.y + v.z + v.w;
                sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, sq))));
            }
        }
    }
    s_sum[slice][lane] = sum;
    s_sq[slice][lane] = sq;
    __syncthreads();
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        ts += s_sum[w][lane];
        tq += s_sq[w][lane];
    }
    if (!rv) return;
    const float mean = ts / (float)a.D;
    const float var = fmaxf(tq / (float)a.D - mean * mean, 0.f);
    const float rstd = rsqrtf(var + a.eps);

    StepCoef sc{};
    bool add_noise = false;
    float tsn = 0.f;
    int64_t tb_ = 0;
    const int bidx = row / a.L, l = row - bidx * a.L;
    if (a.fuse_step) {
        add_noise = __ldg(a.t) != 0;   // the reference branches on t[0] (src/shakespeare.py:349)
        tb_ = __ldg(a.t + bidx);
        sc = step_coef(tb_, a.betas, a.alphas, a.sqrt_om);
        tsn = (float)(tb_ - 1) / 1000.0f;
    }
    for (int pp = slice; pp < npairs; pp += 8) {
        float y[8];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 v = *reinterpret_cast<const float4*>(a.in + (int64_t)(2 * pp + h) * ps + (int64_t)row * 16);
            const int c = pp * 8 + h * 4;
            y[h * 4 + 0] = (v.x - mean) * rstd * __ldg(a.gamma + c + 0) + __ldg(a.beta + c + 0);
            y[h * 4 + 1] = (v.y - mean) * rstd * __ldg(a.gamma + c + 1) + __ldg(a.beta + c + 1);
            y[h * 4 + 2] = (v.z - mean) * rstd * __ldg(a.gamma + c + 2) + __ldg(a.beta + c + 2);
            y[h * 4 + 3] = (v.w - mean) * rstd * __ldg(a.gamma + c + 3) + __ldg(a.beta + c + 3);
        }
        if (a.fuse_step) {
            // y is eps_hat: x_{t-1} = (1/sqrt(alpha_t)) (x_t - beta_t/sqrt(1-acp_t) eps_hat) + sqrt(beta_t) z
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = pp * 8 + h * 4;
                float4* sp = reinterpret_cast<float4*>(a.state + (int64_t)(2 * pp + h) * ps + (int64_t)row * 16);
                const float4 xv = *sp;
                float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
                if (add_noise) {
                    if (a.z) {
                        const float* zr = a.z + (int64_t)row * a.D + c;
                        zz = make_float4(__ldg(zr), __ldg(zr + 1), __ldg(zr + 2), __ldg(zr + 3));
                    } else {
                        zz = philox_normal4(a.seed, a.sample_offset + (uint64_t)bidx, (uint32_t)((l * a.D + c) >> 2),
                                            a.step_id + (uint32_t)tb_, kDomainReverse);
                    }
                }
                float4 xn;
                xn.x = rstep1(sc, xv.x, y[h * 4 + 0], zz.x, add_noise);
                xn.y = rstep1(sc, xv.y, y[h * 4 + 1], zz.y, add_noise);
                xn.z = rstep1(sc, xv.z, y[h * 4 + 2], zz.z, add_noise);
                xn.w = rstep1(sc, xv.w, y[h * 4 + 3], zz.w, add_noise);
                *sp = xn;
                // input of the next step's first GEMM: x_{t-1} + time_emb(t-1) (src/shakespeare.py:116-118)
                y[h * 4 + 0] = xn.x + fmaf(__ldg(a.tw + c + 0), tsn, __ldg(a.tb + c + 0));
                y[h * 4 + 1] = xn.y + fmaf(__ldg(a.tw + c + 1), tsn, __ldg(a.tb + c + 1));
                y[h * 4 + 2] = xn.z + fmaf(__ldg(a.tw + c + 2), tsn, __ldg(a.tb + c + 2));
                y[h * 4 + 3] = xn.w + fmaf(__ldg(a.tw + c + 3), tsn, __ldg(a.tb + c + 3));
            }
        }
        if (a.out32) {
            *reinterpret_cast<float4*>(a.out32 + (int64_t)(2 * pp) * ps + (int64_t)row * 16) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(a.out32 + (int64_t)(2 * pp + 1) * ps + (int64_t)row * 16) = make_float4(y[4], y[5], y[6], y[7]);
        }
        if (a.out16) {
            *reinterpret_cast<uint4*>(a.out16 + (int64_t)pp * ps + (int64_t)row * 16) =
                make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        }
    }
}

// merge the per-(split, group) partial argmaxes: max value, lowest index on ties (torch.argmax)
__global__ void argmax_merge_kernel(const float* __restrict__ pv, const int64_t* __restrict__ pi, int nparts,
                                    int M, int Mp, int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= M) return;
    float best = -INFINITY;
    int64_t bi = INT64_MAX;
    for (int p = 0; p < nparts; ++p) {
        const float v = pv[(int64_t)p * Mp + row];
        const int64_t i = pi[(int64_t)p * Mp + row];
        if (v > best || (v == best && i < bi)) {
            best = v;
            bi = i;
        }
    }
    out_idx[row] = bi;
    if (out_val) out_val[row] = best;
}

// ---------------------------------------------------------------------------------------------
// workspace of the denoiser
// ---------------------------------------------------------------------------------------------
struct TextWs {
    int M, Mp;
    int64_t ps;                  // plane stride (bytes) — same for fp32 and bf16 planes (16 B rows)
    int64_t state, h32, h16;     // x_t (fp32), residual stream (fp32), GEMM operand (bf16)
    int64_t qkv, att, pre, ffn;  // bf16 [3D], bf16 [D], fp32 [D] pre-LN, bf16 [2048]
    int64_t total;
};
constexpr int kFF = 2048;        // nn.TransformerEncoderLayer default dim_feedforward

static TextWs make_text_ws(int64_t B, int L, int D) {
    TextWs w{};
    w.M = (int)(B * L);
    w.Mp = (w.M + 255) / 256 * 256;   // row tiles come in pairs (2-CTA clusters share a weight stream)
    w.ps = (int64_t)w.Mp * 16;
    int64_t o = 0;
    auto take = [&](int64_t planes) {
        int64_t at = o;
        o += (planes * w.ps + 255) / 256 * 256;
        return at;
    };
    w.state = take(D / 4);
    w.h32 = take(D / 4);
    w.h16 = take(D / 8);
    w.qkv = take(3 * D / 8);
    w.att = take(D / 8);
    w.pre = take(D / 4);
    w.ffn = take(kFF / 8);
    w.total = o;
    return w;
}

static int check_text_shape(int64_t B, int L, int D, const char* who) {
    TDM_CHECK_ARG(B > 0 && B * L <= (1 << 24), "%s: bad batch", who);
    TDM_CHECK_ARG(L == 64 || L == 128, "%s: seq_len must be 64 or 128 (got %d)", who, L);
    TDM_CHECK_ARG(D == 256 || D == 2048, "%s: model width must be 256 or 2048 (got %d)", who, D);
    TDM_CHECK_ARG(!(L == 128 && D == 2048), "%s: seq_len 128 with width 2048 is not supported", who);
    return TDM_OK;
}

template <int L, int HD>
static int launch_attn(const uint8_t* qkv, int64_t ps, int D, int heads, int64_t B, uint8_t* out, cudaStream_t st) {
    using C = AttnCfg<L, HD>;
    auto kern = attn_tc_kernel<L, HD>;
    TDM_SET_MAX_DYN_SMEM(kern, C::SMEM);
    launch_pdl(kern, dim3((unsigned)(B * heads)), dim3(128), C::SMEM, st, qkv, ps, D, heads, out, ps);
    TDM_CHECK_LAUNCH("attention");
    return TDM_OK;
}

// per-layer pointer table (host array of device pointers), 12 entries per layer then time_emb w, b
enum : int { LW_QKV = 0, LB_QKV, LW_O, LB_O, LW_1, LB_1, LW_2, LB_2, LN1_G, LN1_B, LN2_G, LN2_B, L_COUNT };

struct TextStep {
    int fuse_step = 0;
    const float* z = nullptr;
    const float* betas = nullptr;
    const float* alphas = nullptr;
    const float* sqrt_om = nullptr;
    uint64_t seed = 0, sample_offset = 0;
    uint32_t step_id = 0;
};

static int text_forward_impl(const void* const* ptrs, int depth, uint8_t* ws, int64_t ws_bytes, const int64_t* t,
                             int64_t B, int L, int D, const TextStep& sa, cudaStream_t st) {
    int rc;
    if ((rc = check_text_shape(B, L, D, "text_forward"))) return rc;
    TDM_CHECK_ARG(ptrs && ws && t && depth > 0, "text_forward: null pointer");
    const TextWs W = make_text_ws(B, L, D);
    TDM_CHECK_ARG(ws_bytes >= W.total, "text_forward: workspace too small");
    const int heads = 4, HD = D / heads;
    const float* tw = reinterpret_cast<const float*>(ptrs[depth * L_COUNT]);
    const float* tb = reinterpret_cast<const float*>(ptrs[depth * L_COUNT + 1]);
    for (int li = 0; li < depth; ++li) {
        const void* const* p = ptrs + li * L_COUNT;
        GemmArgs g{};
        // QKV
        g.a = ws + W.h16; g.a_ps = W.ps; g.w = (const uint8_t*)p[LW_QKV]; g.w_ps = (int64_t)3 * D * 16;
        g.bias = (const float*)p[LB_QKV]; g.M = W.M; g.Mp = W.Mp; g.N = 3 * D; g.n_valid = 3 * D; g.K = D;
        g.out_bf16 = ws + W.qkv; g.ob_ps = W.ps;
        // enough row tiles to fill the GPU: one item per row tile walks all three column tiles with its A tile resident
        // (gemm_tc.cuh); fewer: one item per (row tile, column tile) for parallelism
        g.nsplit = (D == kBN && W.Mp / kBM >= num_sms()) ? 1 : g.N / kBN;
        if ((rc = launch_gemm<GE_BF16>(g, st, "gemm_qkv"))) return rc;
        // attention
        if (L == 64 && HD == 64) rc = launch_attn<64, 64>(ws + W.qkv, W.ps, D, heads, B, ws + W.att, st);
        else if (L == 128 && HD == 64) rc = launch_attn<128, 64>(ws + W.qkv, W.ps, D, heads, B, ws + W.att, st);
        else rc = launch_attn<64, 512>(ws + W.qkv, W.ps, D, heads, B, ws + W.att, st);
        if (rc) return rc;
        // out-proj + residual -> pre-LN
        g = GemmArgs{};
        g.a = ws + W.att; g.a_ps = W.ps; g.w = (const uint8_t*)p[LW_O]; g.w_ps = (int64_t)D * 16;
        g.bias = (const float*)p[LB_O]; g.M = W.M; g.Mp = W.Mp; g.N = D; g.n_valid = D; g.K = D;
        g.res = ws + W.h32; g.res_ps = W.ps; g.nsplit = g.N / kBN;
        LnArgs ln{};
        if (D == kBN) {
            // width 256: the row fits one tile -> LayerNorm fused into the epilogue (in place on h32: each
            // thread reads its own row's residual before it writes the same row)
            g.gamma = (const float*)p[LN1_G]; g.beta = (const float*)p[LN1_B]; g.ln_eps = 1e-5f;
            g.out_f32 = ws + W.h32; g.of_ps = W.ps; g.out_bf16 = ws + W.h16; g.ob_ps = W.ps;
            if ((rc = launch_gemm<GE_RES_LN>(g, st, "gemm_out_proj_ln"))) return rc;
        } else {
            g.out_f32 = ws + W.pre; g.of_ps = W.ps;
            if ((rc = launch_gemm<GE_RES_F32>(g, st, "gemm_out_proj"))) return rc;
            ln.in = ws + W.pre; ln.gamma = (const float*)p[LN1_G]; ln.beta = (const float*)p[LN1_B]; ln.eps = 1e-5f;
            ln.out32 = ws + W.h32; ln.out16 = ws + W.h16; ln.M = W.M; ln.Mp = W.Mp; ln.D = D; ln.L = L;
            launch_pdl(layernorm_kernel, dim3(W.Mp / 32), dim3(256), 0, st, ln);
            TDM_CHECK_LAUNCH("layernorm1");
        }
        if (D == kFfnD) {
            // width 256: the whole feed-forward half (FFN1, ReLU, FFN2, residual, LayerNorm and, on the last
            // layer of a sampling step, the reverse-step update) is one kernel; in place on h32 / h16
            FfnArgs f{};
            f.a = ws + W.h16; f.w1 = (const uint8_t*)p[LW_1]; f.b1 = (const float*)p[LB_1];
            f.w2 = (const uint8_t*)p[LW_2]; f.b2 = (const float*)p[LB_2]; f.res = ws + W.h32;
            f.gamma = (const float*)p[LN2_G]; f.beta = (const float*)p[LN2_B]; f.ln_eps = 1e-5f;
            f.out_f32 = ws + W.h32; f.out_bf16 = ws + W.h16; f.ps = W.ps; f.M = W.M; f.Mp = W.Mp; f.L = L;
            if (li == depth - 1 && sa.fuse_step) {
                f.fuse_step = 1; f.state = ws + W.state; f.t = t; f.z = sa.z; f.betas = sa.betas; f.alphas = sa.alphas;
                f.sqrt_om = sa.sqrt_om; f.tw = tw; f.tb = tb; f.seed = sa.seed; f.sample_offset = sa.sample_offset;
                f.step_id = sa.step_id;
            }
            if ((rc = launch_ffn(f, st))) return rc;
            continue;
        }
        // FFN1 + ReLU
        g = GemmArgs{};
        g.a = ws + W.h16; g.a_ps = W.ps; g.w = (const uint8_t*)p[LW_1]; g.w_ps = (int64_t)kFF * 16;
        g.bias = (const float*)p[LB_1]; g.M = W.M; g.Mp = W.Mp; g.N = kFF; g.n_valid = kFF; g.K = D; g.relu = 1;
        g.out_bf16 = ws + W.ffn; g.ob_ps = W.ps; g.nsplit = g.N / kBN;
        if ((rc = launch_gemm<GE_BF16>(g, st, "gemm_ffn1"))) return rc;
        // FFN2 + residual -> pre-LN
        g = GemmArgs{};
        g.a = ws + W.ffn; g.a_ps = W.ps; g.w = (const uint8_t*)p[LW_2]; g.w_ps = (int64_t)D * 16;
        g.bias = (const float*)p[LB_2]; g.M = W.M; g.Mp = W.Mp; g.N = D; g.n_valid = D; g.K = kFF;
        g.res = ws + W.h32; g.res_ps = W.ps; g.nsplit = g.N / kBN;
        const bool step_here = (li == depth - 1) && sa.fuse_step;
        if (D == kBN && !step_here) {
            g.gamma = (const float*)p[LN2_G]; g.beta = (const float*)p[LN2_B]; g.ln_eps = 1e-5f;
            g.out_f32 = ws + W.h32; g.of_ps = W.ps; g.out_bf16 = ws + W.h16; g.ob_ps = W.ps;
            if ((rc = launch_gemm<GE_RES_LN>(g, st, "gemm_ffn2_ln"))) return rc;
            continue;
        }
        g.out_f32 = ws + W.pre; g.of_ps = W.ps;
        if ((rc = launch_gemm<GE_RES_F32>(g, st, "gemm_ffn2"))) return rc;
        // LN2 (last layer: + reverse step + next time embedding)
        ln = LnArgs{};
        ln.in = ws + W.pre; ln.gamma = (const float*)p[LN2_G]; ln.beta = (const float*)p[LN2_B]; ln.eps = 1e-5f;
        ln.out32 = ws + W.h32; ln.out16 = ws + W.h16; ln.M = W.M; ln.Mp = W.Mp; ln.D = D; ln.L = L;
        if (li == depth - 1 && sa.fuse_step) {
            ln.fuse_step = 1; ln.state = ws + W.state; ln.t = t; ln.z = sa.z; ln.betas = sa.betas;
            ln.alphas = sa.alphas; ln.sqrt_om = sa.sqrt_om; ln.tw = tw; ln.tb = tb; ln.seed = sa.seed;
            ln.sample_offset = sa.sample_offset; ln.step_id = sa.step_id;
        }
        launch_pdl(layernorm_kernel, dim3(W.Mp / 32), dim3(256), 0, st, ln);
        TDM_CHECK_LAUNCH("layernorm2");
    }
    return TDM_OK;
}

}  // namespace tdm

using namespace tdm;

extern "C" int64_t tdm_text_workspace_bytes(int64_t batch, int seq_len, int dim) {
    if (batch <= 0 || seq_len <= 0 || dim <= 0 || dim % 8) return 0;
    return make_text_ws(batch, seq_len, dim).total;
}

extern "C" int tdm_pack_linear(const float* w, int n, int k, int n_padded, void* out_planes, void* stream) {
    TDM_CHECK_ARG(w && out_planes && n > 0 && k > 0 && k % 8 == 0 && n_padded >= n, "tdm_pack_linear: bad arguments");
    const int64_t total = (int64_t)(k / 8) * n_padded;
    const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    pack_linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, n, k, n_padded, reinterpret_cast<__nv_bfloat16*>(out_planes));
    TDM_CHECK_LAUNCH("tdm_pack_linear");
    return TDM_OK;
}

extern "C" int tdm_pack_ffn_weights(const float* w1, const float* w2, void* out1, void* out2, void* stream) {
    TDM_CHECK_ARG(w1 && w2 && out1 && out2, "tdm_pack_ffn_weights: null pointer");
    constexpr int kRows = 16 * 4 * 8 * 128 + 16 * 2 * 8 * 256;
    pack_ffn_weights_kernel<<<(kRows + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w1, w2, reinterpret_cast<__nv_bfloat16*>(out1),
                                                                                 reinterpret_cast<__nv_bfloat16*>(out2));
    TDM_CHECK_LAUNCH("tdm_pack_ffn_weights");
    return TDM_OK;
}

extern "C" int tdm_text_load_state(const float* x_rows, const int64_t* t, const float* time_w, const float* time_b,
                                   void* workspace, int64_t workspace_bytes, int64_t batch, int seq_len, int dim,
                                   void* stream) {
    int rc;
    if ((rc = check_text_shape(batch, seq_len, dim, "tdm_text_load_state"))) return rc;
    TDM_CHECK_ARG(x_rows && t && time_w && time_b && workspace, "tdm_text_load_state: null pointer");
    const TextWs W = make_text_ws(batch, seq_len, dim);
    TDM_CHECK_ARG(workspace_bytes >= W.total, "tdm_text_load_state: workspace too small");
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    rows_to_planes_kernel<<<W.Mp / 32, 256, 0, (cudaStream_t)stream>>>(x_rows, W.M, W.Mp, dim, seq_len, t, time_w,
                                                                       time_b, ws + W.state, ws + W.h32, ws + W.h16,
                                                                       nullptr);
    TDM_CHECK_LAUNCH("tdm_text_load_state");
    return TDM_OK;
}

extern "C" int tdm_text_read(const void* workspace, int64_t workspace_bytes, int which, float* out_rows,
                             int64_t batch, int seq_len, int dim, void* stream) {
    int rc;
    if ((rc = check_text_shape(batch, seq_len, dim, "tdm_text_read"))) return rc;
    TDM_CHECK_ARG(workspace && out_rows && (which == 0 || which == 1), "tdm_text_read: bad arguments");
    const TextWs W = make_text_ws(batch, seq_len, dim);
    TDM_CHECK_ARG(workspace_bytes >= W.total, "tdm_text_read: workspace too small");
    const uint8_t* ws = reinterpret_cast<const uint8_t*>(workspace);
    planes_to_rows_kernel<<<W.Mp / 32, 256, 0, (cudaStream_t)stream>>>(ws + (which == 0 ? W.state : W.h32), W.M, W.Mp,
                                                                       dim, out_rows);
    TDM_CHECK_LAUNCH("tdm_text_read");
    return TDM_OK;
}

extern "C" int tdm_text_forward(const void* const* host_ptrs, int depth, void* workspace, int64_t workspace_bytes,
                                const int64_t* t, int64_t batch, int seq_len, int dim, void* stream) {
    TextStep sa;
    return text_forward_impl(host_ptrs, depth, reinterpret_cast<uint8_t*>(workspace), workspace_bytes, t, batch,
                             seq_len, dim, sa, (cudaStream_t)stream);
}

extern "C" int tdm_text_p_sample(const void* const* host_ptrs, int depth, void* workspace, int64_t workspace_bytes,
                                 const int64_t* t, const float* z_rows, const float* betas, const float* alphas,
                                 const float* sqrt_om_acp, int64_t batch, int seq_len, int dim, uint64_t seed,
                                 uint64_t sample_offset, uint32_t step_id, void* stream) {
    TDM_CHECK_ARG(betas && alphas && sqrt_om_acp, "tdm_text_p_sample: null schedule table");
    TextStep sa;
    sa.fuse_step = 1; sa.z = z_rows; sa.betas = betas; sa.alphas = alphas; sa.sqrt_om = sqrt_om_acp;
    sa.seed = seed; sa.sample_offset = sample_offset; sa.step_id = step_id;
    return text_forward_impl(host_ptrs, depth, reinterpret_cast<uint8_t*>(workspace), workspace_bytes, t, batch,
                             seq_len, dim, sa, (cudaStream_t)stream);
}

extern "C" int64_t tdm_round_workspace_bytes(int64_t rows, int dim, int64_t vocab) {
    if (rows <= 0 || dim <= 0 || vocab <= 0) return 0;
    const int64_t Mp = (rows + 127) / 128 * 128;
    const int64_t n_tiles = (vocab + kBN - 1) / kBN;
    const int64_t nsplit = n_tiles < 148 ? n_tiles : 148;
    int64_t o = (dim / 8) * Mp * 16;              // bf16 planes of x
    o = (o + 255) / 256 * 256;
    o += Mp * 4;                                  // row norms
    o = (o + 255) / 256 * 256;
    o += 2 * nsplit * Mp * (4 + 8);               // partial (value, index)
    return o + 512;
}

extern "C" int tdm_round_argmax(const float* x_rows, int64_t rows, int dim, const void* w_planes, int64_t vocab,
                                int64_t vocab_padded, const float* bias, int cosine, const float* ar_logits,
                                int64_t ar_ld, float alpha, float temperature, int64_t* out_idx, float* out_val,
                                void* workspace, int64_t workspace_bytes, void* stream) {
    TDM_CHECK_ARG(x_rows && w_planes && out_idx && workspace, "tdm_round_argmax: null pointer");
    TDM_CHECK_ARG(rows > 0 && dim % 64 == 0 && dim % 128 == 0 && vocab > 0 && vocab_padded % kBN == 0 && vocab_padded >= vocab,
                  "tdm_round_argmax: bad shape rows=%lld dim=%d vocab=%lld/%lld", (long long)rows, dim,
                  (long long)vocab, (long long)vocab_padded);
    TDM_CHECK_ARG(workspace_bytes >= tdm_round_workspace_bytes(rows, dim, vocab), "tdm_round_argmax: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int M = (int)rows, Mp = (M + 127) / 128 * 128;
    const int n_tiles = (int)(vocab_padded / kBN);
    const int m_tiles = Mp / kBM;
    // Work items = m_tiles x nsplit (each item: one 128-row tile against 1/nsplit of the vocabulary).  Few row tiles:
    // split the vocabulary over all SMs.  Many row tiles: pick the split whose item count fills whole waves of CTAs
    // (256 row tiles unsplit are 1.73 waves - 14 % of the SM-time idle; split 15 ways they are 25.95 waves).
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
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    int64_t o = 0;
    uint8_t* xp = ws;
    o = ((int64_t)(dim / 8) * Mp * 16 + 255) / 256 * 256;
    float* rnorm = reinterpret_cast<float*>(ws + o);
    o = (o + (int64_t)Mp * 4 + 255) / 256 * 256;
    float* pv = reinterpret_cast<float*>(ws + o);
    int64_t* pi = reinterpret_cast<int64_t*>(ws + o + (int64_t)2 * nsplit * Mp * 4);
    // 8-byte alignment of pi: 2*nsplit*Mp*4 is a multiple of 8 since Mp % 128 == 0
    rows_to_planes_kernel<<<Mp / 32, 256, 0, st>>>(x_rows, M, Mp, dim, 1, nullptr, nullptr, nullptr, nullptr, nullptr, xp,
                                                   rnorm);
    TDM_CHECK_LAUNCH("round: rows_to_planes");
    GemmArgs g{};
    g.a = xp; g.a_ps = (int64_t)Mp * 16; g.w = reinterpret_cast<const uint8_t*>(w_planes); g.w_ps = vocab_padded * 16;
    g.bias = bias; g.M = M; g.Mp = Mp; g.N = (int)vocab_padded; g.n_valid = (int)vocab; g.K = dim;
    // cosine similarity = x.e / (||x|| ||e||): e is pre-normalised in w_planes; the 1/||x|| row scale only
    // matters when the similarities are mixed with AR logits (a positive row scale cannot move an argmax)
    g.row_scale = (cosine && ar_logits) ? rnorm : nullptr;
    g.ar = ar_logits; g.ar_ld = ar_ld; g.alpha = alpha; g.inv_temp = 1.0f / temperature;
    g.part_val = pv; g.part_idx = pi; g.nsplit = nsplit;
    int rc;
    if ((rc = launch_gemm<GE_ARGMAX>(g, st, "round_argmax"))) return rc;
    argmax_merge_kernel<<<(M + 127) / 128, 128, 0, st>>>(pv, pi, 2 * nsplit, M, Mp, out_idx, out_val);
    TDM_CHECK_LAUNCH("argmax_merge");
    return TDM_OK;
}

// ---------------------------------------------------------------------------------------------
// LearnedRounding.forward / LearnedEmbedding.forward (src/shakespeare.py:71-80, 93-102): the two module calls
// the samplers never make (they fuse the GEMM with the argmax) but code written against the reference does
// ---------------------------------------------------------------------------------------------
extern "C" int tdm_linear_logits(const float* x_rows, int64_t rows, int dim, const void* w_planes, int64_t vocab,
                                 int64_t vocab_padded, const float* bias, int cosine, float* out_logits, int64_t out_ld,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
    TDM_CHECK_ARG(x_rows && w_planes && out_logits && workspace, "tdm_linear_logits: null pointer");
    TDM_CHECK_ARG(rows > 0 && dim % 128 == 0 && vocab > 0 && vocab_padded % kBN == 0 && vocab_padded >= vocab && out_ld >= vocab,
                  "tdm_linear_logits: bad shape rows=%lld dim=%d vocab=%lld/%lld ld=%lld", (long long)rows, dim,
                  (long long)vocab, (long long)vocab_padded, (long long)out_ld);
    TDM_CHECK_ARG(workspace_bytes >= tdm_round_workspace_bytes(rows, dim, vocab), "tdm_linear_logits: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int M = (int)rows, Mp = (M + 127) / 128 * 128;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    uint8_t* xp = ws;
    const int64_t o = ((int64_t)(dim / 8) * Mp * 16 + 255) / 256 * 256;
    float* rnorm = reinterpret_cast<float*>(ws + o);
    rows_to_planes_kernel<<<Mp / 32, 256, 0, st>>>(x_rows, M, Mp, dim, 1, nullptr, nullptr, nullptr, nullptr, nullptr, xp,
                                                   rnorm);
    TDM_CHECK_LAUNCH("logits: rows_to_planes");
    GemmArgs g{};
    g.a = xp; g.a_ps = (int64_t)Mp * 16; g.w = reinterpret_cast<const uint8_t*>(w_planes); g.w_ps = vocab_padded * 16;
    g.bias = bias; g.M = M; g.Mp = Mp; g.N = (int)vocab_padded; g.n_valid = (int)vocab; g.K = dim;
    g.row_scale = cosine ? rnorm : nullptr;   // cosine similarity: the table in w_planes is pre-normalised
    g.nsplit = (int)(vocab_padded / kBN);
    g.logits = out_logits; g.logits_ld = out_ld;
    return launch_gemm<GE_LOGITS>(g, st, "linear_logits");
}

__global__ void embedding_gather_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids, int64_t n,
                                        int64_t vocab, int dim4, float* __restrict__ out, int* __restrict__ bad) {
    const float4* t4 = reinterpret_cast<const float4*>(table);
    float4* o4 = reinterpret_cast<float4*>(out);
    for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
        const int64_t id = ids[row];
        if (id < 0 || id >= vocab) {   // nn.Embedding raises: report it, write zeros
            if (threadIdx.x == 0 && bad) atomicExch(bad, 1);
            for (int j = threadIdx.x; j < dim4; j += blockDim.x) o4[row * dim4 + j] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        for (int j = threadIdx.x; j < dim4; j += blockDim.x) o4[row * dim4 + j] = __ldg(t4 + id * dim4 + j);
    }
}

extern "C" int tdm_embedding_gather(const float* table, int64_t vocab, int dim, const int64_t* ids, int64_t n,
                                    float* out_rows, int* bad_flag, void* stream) {
    TDM_CHECK_ARG(table && ids && out_rows && vocab > 0 && dim > 0 && dim % 4 == 0 && n >= 0,
                  "tdm_embedding_gather: bad arguments");
    if (n == 0) return TDM_OK;
    const unsigned grid = (unsigned)(n < 148 * 16 ? n : 148 * 16);
    embedding_gather_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(table, ids, n, vocab, dim / 4, out_rows, bad_flag);
    TDM_CHECK_LAUNCH("tdm_embedding_gather");
    return TDM_OK;
}

#ifdef TDM_EXP_TIMELINE
extern "C" int tdm_debug_gemm_timeline(unsigned long long* host16) {
    return cudaMemcpyFromSymbol(host16, tdm::g_gemm_tl, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 2;
}
extern "C" int tdm_debug_ffn_timeline(unsigned long long* host128) {
    return cudaMemcpyFromSymbol(host128, tdm::g_ffn_tl, sizeof(unsigned long long) * 144) == cudaSuccess ? 0 : 2;
}
#endif
