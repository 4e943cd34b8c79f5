// unet_fwd.cu — SimpleUNet.forward (src/mnist.py:76-87) and the fused p_sample
// (src/mnist.py:167-180) as nine launches:
//
//   k1  rb1.conv1   1->32  @28   CUDA cores (K = 9 is too thin for the tensor pipe), fp32 in
//   k2  rb1.conv2   32->32 @28   tcgen05, epilogue relu + 1x1 skip of x           -> h1 (cat[8:12])
//   k3  avg_pool2                                                                  -> p1
//   k4  rb2.conv1   32->64 @14   tcgen05 (+ 1x1 skip GEMM into a 2nd accumulator)  -> t2, s2
//   k5  rb2.conv2   64->64 @14   tcgen05, epilogue relu + s2                       -> h2
//   k6  rb3.conv1   64->64 @14   tcgen05                                           -> t3
//   k7  rb3.conv2   64->64 @14   tcgen05, epilogue relu + h2, nearest x2 scatter   -> cat[0:8]
//   k8  rb4.conv1   96->32 @28   tcgen05 (+ 1x1 skip GEMM)                         -> t4, s4
//   k9  rb4.conv2   32->32 @28   tcgen05, epilogue relu + s4, 1x1 out conv, and (p_sample) the
//                                reverse-step update with in-kernel Philox noise   -> eps | x_{t-1}
//
// Every tensor-core conv is the same persistent warp-specialised kernel: warp 0 streams input
// tiles (one bulk async copy per 8-channel plane, halo included) into a ring of smem stages,
// warp 1 issues 9*Cin/16 tcgen05.mma (M=128 positions, N=Cout, K=16) per tile whose A operand is
// the *same* smem tile addressed at nine different row offsets, warps 2-5 drain the TMEM
// accumulator (double buffered) through the fused epilogue.  Weights stay resident in smem.
#include "common.cuh"
#include "diffusion_math.cuh"
#include "tc05.cuh"
#include "unet_layout.cuh"

namespace tdm {

// ---------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------
struct UNetWs {
    int64_t nt28, nt14, ps28, ps14;  // tiles, plane stride in bytes
    int64_t t1, cat, p1, t2, s2, h2, t3, t4, s4, total;
};

static UNetWs make_ws(int64_t batch) {
    UNetWs w;
    w.nt28 = num_tiles(batch, Geo<28>::S);
    w.nt14 = num_tiles(batch, Geo<14>::S);
    w.ps28 = plane_rows(batch, Geo<28>::S, Geo<28>::HALO) * 16;
    w.ps14 = plane_rows(batch, Geo<14>::S, Geo<14>::HALO) * 16;
    int64_t o = 0;
    auto take = [&](int64_t planes, int64_t ps) {
        int64_t at = o;
        o += (planes * ps + 255) / 256 * 256;
        return at;
    };
    w.t1 = take(4, w.ps28);
    w.cat = take(12, w.ps28);
    w.p1 = take(4, w.ps14);
    w.t2 = take(8, w.ps14);
    w.s2 = take(8, w.ps14);
    w.h2 = take(8, w.ps14);
    w.t3 = take(8, w.ps14);
    w.t4 = take(4, w.ps28);
    w.s4 = take(4, w.ps28);
    w.total = o;
    return w;
}

// ---------------------------------------------------------------------------------------------
// weight packing: flat fp32 (state_dict order, OIHW) -> bf16 [tap][Cin/8][Cout][8]
// ---------------------------------------------------------------------------------------------
struct PackJob {
    int src;       // offset into flat params
    int64_t dst;   // byte offset into wpack
    int cin, cout, taps;
};
__constant__ PackJob c_jobs[9] = {
    {P::rb1_c2w, WP::rb1_c2, 32, 32, 9}, {P::rb2_c1w, WP::rb2_c1, 32, 64, 9},
    {P::rb2_sw, WP::rb2_sk, 32, 64, 1},  {P::rb2_c2w, WP::rb2_c2, 64, 64, 9},
    {P::rb3_c1w, WP::rb3_c1, 64, 64, 9}, {P::rb3_c2w, WP::rb3_c2, 64, 64, 9},
    {P::rb4_c1w, WP::rb4_c1, 96, 32, 9}, {P::rb4_sw, WP::rb4_sk, 96, 32, 1},
    {P::rb4_c2w, WP::rb4_c2, 32, 32, 9}};

__global__ void pack_weights_kernel(const float* __restrict__ flat, uint8_t* __restrict__ wpack) {
    const PackJob j = c_jobs[blockIdx.y];
    const int n = j.taps * j.cin * j.cout;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(wpack + j.dst);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        // i enumerates the destination: (((tap*(cin/8) + cp)*cout + co)*8 + k)
        const int k = i & 7;
        int r = i >> 3;
        const int co = r % j.cout;
        r /= j.cout;
        const int cp = r % (j.cin / 8);
        const int tap = r / (j.cin / 8);
        const int ci = cp * 8 + k;
        // OIHW source: ((co*cin + ci)*taps + tap)
        dst[i] = __float2bfloat16_rn(flat[j.src + (co * j.cin + ci) * j.taps + tap]);
    }
    if (blockIdx.y == 0) {
        float* f = reinterpret_cast<float*>(wpack + WP::flat);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P::count; i += gridDim.x * blockDim.x)
            f[i] = flat[i];
    }
}

// ---------------------------------------------------------------------------------------------
// k1: rb1.conv1 (1 -> 32), relu, + time bias.  One thread per position, fp32 math.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
rb1_conv1_kernel(const float* __restrict__ x, const int64_t* __restrict__ t,
                 const float* __restrict__ fp, uint8_t* __restrict__ out, int64_t out_ps, int batch) {
    using G = Geo<28>;
    __shared__ float s_w[32 * 9], s_b[32], s_tw[32], s_tb[32];
    for (int i = threadIdx.x; i < 288; i += 128) s_w[i] = fp[P::rb1_c1w + i];
    if (threadIdx.x < 32) {
        s_b[threadIdx.x] = fp[P::rb1_c1b + threadIdx.x];
        s_tw[threadIdx.x] = fp[P::rb1_tw + threadIdx.x];
        s_tb[threadIdx.x] = fp[P::rb1_tb + threadIdx.x];
    }
    __syncthreads();
    const int64_t pos = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int b = (int)(pos / G::S);
    const int rem = (int)(pos - (int64_t)b * G::S);
    const int r = rem / G::Wp, c = rem - r * G::Wp;
    const bool valid = b < batch && r >= 1 && c < G::W;
    uint4 o[4] = {};
    if (valid) {
        const int y = r - 1;
        const float* img = x + (int64_t)b * 784;
        float v[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int yy = y + ky - 1, xx = c + kx - 1;
                v[ky * 3 + kx] = (yy >= 0 && yy < 28 && xx >= 0 && xx < 28) ? __ldg(img + yy * 28 + xx) : 0.f;
            }
        const float ts = (float)__ldg(t + b) / 1000.0f;  // src/mnist.py:77
        uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
        for (int co = 0; co < 32; co += 2) {
            float a0 = s_b[co], a1 = s_b[co + 1];
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                a0 = fmaf(s_w[co * 9 + k], v[k], a0);
                a1 = fmaf(s_w[(co + 1) * 9 + k], v[k], a1);
            }
            a0 = fmaxf(a0, 0.f) + fmaf(s_tw[co], ts, s_tb[co]);
            a1 = fmaxf(a1, 0.f) + fmaf(s_tw[co + 1], ts, s_tb[co + 1]);
            ow[co / 2] = pack_bf16x2(a0, a1);
        }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p)
        *reinterpret_cast<uint4*>(out + p * out_ps + (pos + G::HALO) * 16) = o[p];
}

// ---------------------------------------------------------------------------------------------
// k3: 2x2 average pool, 28-geometry planes -> 14-geometry planes (src/mnist.py:80)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
avgpool_kernel(const uint8_t* __restrict__ in, int64_t in_ps, uint8_t* __restrict__ out,
               int64_t out_ps, int batch) {
    using GI = Geo<28>;
    using GO = Geo<14>;
    const int64_t pos = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int plane = blockIdx.y;
    const int b = (int)(pos / GO::S);
    const int rem = (int)(pos - (int64_t)b * GO::S);
    const int r = rem / GO::Wp, c = rem - r * GO::Wp;
    const bool valid = b < batch && r >= 1 && c < GO::W;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (valid) {
        const int y = r - 1;
        const int64_t p00 = (int64_t)b * GI::S + (2 * y + 1) * GI::Wp + 2 * c;
        const uint8_t* src = in + plane * in_ps + (p00 + GI::HALO) * 16;
        const uint4 q0 = *reinterpret_cast<const uint4*>(src);
        const uint4 q1 = *reinterpret_cast<const uint4*>(src + 16);
        const uint4 q2 = *reinterpret_cast<const uint4*>(src + GI::Wp * 16);
        const uint4 q3 = *reinterpret_cast<const uint4*>(src + GI::Wp * 16 + 16);
        const uint32_t* a0 = &q0.x; const uint32_t* a1 = &q1.x;
        const uint32_t* a2 = &q2.x; const uint32_t* a3 = &q3.x;
        uint32_t* ow = &o.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f0 = unpack_bf16x2(a0[k]), f1 = unpack_bf16x2(a1[k]);
            const float2 f2 = unpack_bf16x2(a2[k]), f3 = unpack_bf16x2(a3[k]);
            ow[k] = pack_bf16x2((f0.x + f1.x + f2.x + f3.x) * 0.25f, (f0.y + f1.y + f2.y + f3.y) * 0.25f);
        }
    }
    *reinterpret_cast<uint4*>(out + plane * out_ps + (pos + GO::HALO) * 16) = o;
}

// ---------------------------------------------------------------------------------------------
// tensor-core 3x3 convolution
// ---------------------------------------------------------------------------------------------
enum : int { EPI_CONV1 = 0, EPI_RES = 1, EPI_RES_X = 2, EPI_RES_UP = 3, EPI_FINAL = 4 };

struct ConvArgs {
    const uint8_t* in;     // input planes: row -HALO of plane 0
    int64_t in_ps;         // plane stride (bytes)
    const uint8_t* w;      // packed bf16 weights: conv, then (SKIPG) the 1x1 skip
    const float* bias;     // [COUT] conv bias
    const float* tw;       // [COUT] time_emb.weight   (EPI_CONV1)
    const float* tb;       // [COUT] time_emb.bias     (EPI_CONV1)
    const float* sbias;    // [COUT] skip bias         (SKIPG)
    const int64_t* t;      // [B]
    uint8_t* out;          // output planes
    int64_t out_ps;
    uint8_t* out2;         // skip output planes (SKIPG)
    int64_t out2_ps;
    const uint8_t* res;    // residual planes (EPI_RES / EPI_RES_UP / EPI_FINAL)
    int64_t res_ps;
    const float* x;        // [B,784] fp32: rb1 skip input (EPI_RES_X) or x_t (EPI_FINAL + step)
    const float* aux_w;    // [32]: rb1.skip.weight (EPI_RES_X) or out.weight (EPI_FINAL)
    const float* aux_b;    // [32] rb1.skip.bias / [1] out.bias
    float* fout;           // EPI_FINAL: eps or x_{t-1}, [B,784] fp32
    const float* z;        // injected noise or null (Philox)
    const float* betas;
    const float* alphas;
    const float* sqrt_om;
    uint64_t seed;
    uint64_t sample_offset;
    uint32_t step_id;
    int fuse_step;
    int nt;
    int batch;
};

constexpr int kEpiGroups = 2;  // epilogue warp groups == TMEM accumulator stages

template <int W, int CIN, int COUT, bool SKIPG>
struct ConvCfg {
    using G = Geo<W>;
    static constexpr int NPL = CIN / 8;
    static constexpr int STAGE_BYTES = NPL * G::RT * 16;
    static constexpr int WCONV_BYTES = 9 * CIN * COUT * 2;
    static constexpr int W_BYTES = WCONV_BYTES + (SKIPG ? CIN * COUT * 2 : 0);
    static constexpr int PARAM_BYTES = 5 * 64 * 4;
    static constexpr int MAX_SMEM = 227 * 1024;
    static constexpr int AVAIL = MAX_SMEM - W_BYTES - PARAM_BYTES - 256;
    static constexpr int NSTAGE = (AVAIL / STAGE_BYTES) > 4 ? 4 : (AVAIL / STAGE_BYTES);
    static_assert(NSTAGE >= 2, "need at least two input stages");
    static constexpr int NACC = kEpiGroups;
    static constexpr int ACC_COLS = SKIPG ? 2 * COUT : COUT;
    static constexpr int TMEM_COLS = (NACC * ACC_COLS <= 32) ? 32 : (NACC * ACC_COLS <= 64) ? 64
                                   : (NACC * ACC_COLS <= 128) ? 128 : (NACC * ACC_COLS <= 256) ? 256 : 512;
    static_assert(NACC * ACC_COLS <= 512, "accumulators exceed TMEM");
    static constexpr int SMEM_BYTES = W_BYTES + NSTAGE * STAGE_BYTES + PARAM_BYTES + 256;
    // warp 0 producer, warp 1 MMA issuer, then NACC groups of 4 epilogue warps
    static constexpr int THREADS = 64 + 128 * NACC;
};

template <int W, int CIN, int COUT, int EPI, bool SKIPG>
__global__ void __launch_bounds__(64 + 128 * kEpiGroups, 1) conv3x3_tc_kernel(const ConvArgs a) {
    using C = ConvCfg<W, CIN, COUT, SKIPG>;
    using G = Geo<W>;
    static_assert(COUT == 32 || COUT == 64, "COUT");
    static_assert(EPI != EPI_FINAL || COUT == 32, "final epilogue expects 32 channels");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_in = smem + C::W_BYTES;
    float* s_par = reinterpret_cast<float*>(s_in + C::NSTAGE * C::STAGE_BYTES);
    float* s_bias = s_par;
    float* s_tw = s_par + 64;
    float* s_tb = s_par + 128;
    float* s_sbias = s_par + 192;
    float* s_aux = s_par + 256;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_par + 320);
    uint64_t* bar_w = bars;
    uint64_t* bar_full = bars + 1;
    uint64_t* bar_empty = bar_full + C::NSTAGE;
    uint64_t* bar_accf = bar_empty + C::NSTAGE;
    uint64_t* bar_acce = bar_accf + C::NACC;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acce + C::NACC);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- setup -----------------------------------------------------------------------------
    if (threadIdx.x < COUT) {
        const int c = threadIdx.x;
        s_bias[c] = a.bias[c];
        s_tw[c] = (EPI == EPI_CONV1) ? a.tw[c] : 0.f;
        s_tb[c] = (EPI == EPI_CONV1) ? a.tb[c] : 0.f;
        s_sbias[c] = SKIPG ? a.sbias[c] : 0.f;
        s_aux[c] = (EPI == EPI_RES_X || EPI == EPI_FINAL) ? a.aux_w[c] : 0.f;
        if (EPI == EPI_RES_X) s_aux[32 + c] = a.aux_b[c];
        if (EPI == EPI_FINAL && c == 0) s_aux[32] = a.aux_b[0];
    }
    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < C::NSTAGE; ++i) {
            mbar_init(bar_full + i, 1);
            mbar_init(bar_empty + i, 1);
        }
        for (int i = 0; i < C::NACC; ++i) {
            mbar_init(bar_accf + i, 1);
            mbar_init(bar_acce + i, 4);  // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<C::TMEM_COLS>(s_tmem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;

    // ---- roles -----------------------------------------------------------------------------
    if (warp == 0) {
        // ===== producer: weights once, then one input tile per iteration =====
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_w, C::W_BYTES);
            constexpr int CH = 16384;
            for (int off = 0; off < C::W_BYTES; off += CH) {
                const int n = (C::W_BYTES - off) < CH ? (C::W_BYTES - off) : CH;
                bulk_g2s(s_w + off, a.w + off, n, bar_w);
            }
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < a.nt; tile += gridDim.x, ++it) {
            const int s = it % C::NSTAGE;
            const uint32_t ph = (it / C::NSTAGE) & 1;
            if (lane == 0) {
                mbar_wait(bar_empty + s, ph ^ 1);
                mbar_arrive_expect_tx(bar_full + s, C::STAGE_BYTES);
            }
            __syncwarp();
            if (lane < C::NPL) {
                bulk_g2s(s_in + s * C::STAGE_BYTES + lane * (G::RT * 16),
                         a.in + lane * a.in_ps + (int64_t)tile * (kTile * 16), G::RT * 16,
                         bar_full + s);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, COUT);
            mbar_wait(bar_w, 0);
            const uint32_t w_addr = smem_u32(s_w);
            int it = 0;
            for (int tile = blockIdx.x; tile < a.nt; tile += gridDim.x, ++it) {
                const int s = it % C::NSTAGE;
                const uint32_t ph = (it / C::NSTAGE) & 1;
                const int acc = it % C::NACC;
                const uint32_t aph = (it / C::NACC) & 1;
                mbar_wait(bar_acce + acc, aph ^ 1);
                mbar_wait(bar_full + s, ph);
                tc_fence_after_sync();
                const uint32_t in_addr = smem_u32(s_in + s * C::STAGE_BYTES);
                const uint32_t d = tmem_base + acc * C::ACC_COLS;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int off = (tap / 3 - 1) * G::Wp + (tap % 3 - 1);
#pragma unroll
                    for (int ks = 0; ks < CIN / 16; ++ks) {
                        const uint64_t ad = make_smem_desc(
                            in_addr + (2 * ks) * (G::RT * 16) + (G::HALO + off) * 16, G::RT * 16, 128);
                        const uint64_t bd = make_smem_desc(
                            w_addr + ((tap * C::NPL + 2 * ks) * COUT) * 16, COUT * 16, 128);
                        umma_bf16(d, ad, bd, idesc, (tap | ks) != 0);
                    }
                }
                if constexpr (SKIPG) {
#pragma unroll
                    for (int ks = 0; ks < CIN / 16; ++ks) {
                        const uint64_t ad = make_smem_desc(
                            in_addr + (2 * ks) * (G::RT * 16) + G::HALO * 16, G::RT * 16, 128);
                        const uint64_t bd = make_smem_desc(
                            w_addr + C::WCONV_BYTES + ((2 * ks) * COUT) * 16, COUT * 16, 128);
                        umma_bf16(d + COUT, ad, bd, idesc, ks != 0);
                    }
                }
                umma_commit(bar_empty + s);   // smem stage reusable once these MMAs retire
                umma_commit(bar_accf + acc);  // accumulator complete
            }
        }
    } else {
        // ===== epilogue: group g = (warp-2)/4 owns accumulator stage g (tiles it = g, g+NACC, ..),
        //       so the epilogues of consecutive tiles overlap; TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * C::ACC_COLS;
        int n = 0;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < a.nt; tile += C::NACC * gridDim.x, ++n) {
            const uint32_t aph = n & 1;
            // ---- phase A: everything that does not need the accumulator (overlaps the MMAs) ----
            const int64_t pos = (int64_t)tile * kTile + q * 32 + lane;
            const int b = (int)(pos / G::S);
            const int rem = (int)(pos - (int64_t)b * G::S);
            const int r = rem / G::Wp, c = rem - r * G::Wp;
            const bool valid = b < a.batch && r >= 1 && c < G::W;
            const int y = r - 1;

            float ts = 0.f;
            if (EPI == EPI_CONV1 && valid) ts = (float)__ldg(a.t + b) / 1000.0f;
            float xin = 0.f;
            if ((EPI == EPI_RES_X || EPI == EPI_FINAL) && valid && a.x)
                xin = __ldg(a.x + (int64_t)b * 784 + y * 28 + c);
            constexpr bool kHasRes = (EPI == EPI_RES || EPI == EPI_RES_UP || EPI == EPI_FINAL);
            uint4 rv[kHasRes ? COUT / 8 : 1];
            if constexpr (kHasRes) {
#pragma unroll
                for (int pl = 0; pl < COUT / 8; ++pl) {
                    rv[pl] = make_uint4(0, 0, 0, 0);
                    if (valid)  // residual planes share this kernel's position geometry
                        rv[pl] = *reinterpret_cast<const uint4*>(a.res + pl * a.res_ps + (pos + G::HALO) * 16);
                }
            }
            StepCoef sc{};
            float zz = 0.f;
            bool add_noise = false;
            if constexpr (EPI == EPI_FINAL) {
                if (a.fuse_step && valid) {
                    add_noise = __ldg(a.t) != 0;  // src/mnist.py:176
                    const int64_t tb = __ldg(a.t + b);
                    sc = step_coef(tb, a.betas, a.alphas, a.sqrt_om);
                    if (add_noise) {
                        const int e = y * 28 + c;
                        if (a.z) {
                            zz = __ldg(a.z + (int64_t)b * 784 + e);
                        } else {
                            const float4 n4 = philox_normal4(a.seed, a.sample_offset + (uint64_t)b,
                                                             (uint32_t)(e >> 2), a.step_id + (uint32_t)tb,
                                                             kDomainReverse);
                            const int k = e & 3;
                            zz = k == 0 ? n4.x : k == 1 ? n4.y : k == 2 ? n4.z : n4.w;
                        }
                    }
                }
            }

            // ---- phase B: drain the accumulator ----
            mbar_wait(bar_accf + grp, aph);
            tc_fence_after_sync();
            float dot = 0.f;

#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t r1[32];
                tmem_ld32(taddr + c0, r1);
                uint32_t r2[32];
                if constexpr (SKIPG) tmem_ld32(taddr + COUT + c0, r2);
                tmem_ld_wait();
                if (c0 + 32 >= COUT) {
                    // all TMEM reads of this accumulator are done: hand it back to the MMA warp
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acce + grp);
                }
#pragma unroll
                for (int pj = 0; pj < 4; ++pj) {
                    const int plane = c0 / 8 + pj;
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int ch = c0 + pj * 8 + k;
                        v[k] = fmaxf(__uint_as_float(r1[pj * 8 + k]) + s_bias[ch], 0.f);
                    }
                    if constexpr (EPI == EPI_CONV1) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int ch = c0 + pj * 8 + k;
                            v[k] += fmaf(s_tw[ch], ts, s_tb[ch]);
                        }
                    } else if constexpr (EPI == EPI_RES_X) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int ch = c0 + pj * 8 + k;
                            v[k] += fmaf(s_aux[ch], xin, s_aux[32 + ch]);
                        }
                    } else {
                        const uint32_t* rw = &rv[plane].x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 f = unpack_bf16x2(rw[k]);
                            v[2 * k] += f.x;
                            v[2 * k + 1] += f.y;
                        }
                    }
                    if constexpr (EPI == EPI_FINAL) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) dot = fmaf(s_aux[c0 + pj * 8 + k], v[k], dot);
                    } else {
                        uint4 o;
                        o.x = valid ? pack_bf16x2(v[0], v[1]) : 0u;
                        o.y = valid ? pack_bf16x2(v[2], v[3]) : 0u;
                        o.z = valid ? pack_bf16x2(v[4], v[5]) : 0u;
                        o.w = valid ? pack_bf16x2(v[6], v[7]) : 0u;
                        if constexpr (EPI == EPI_RES_UP) {
                            // nearest x2 upsample (src/mnist.py:83): one 14x14 pixel -> 2x2 block
                            // of the 28x28 geometry; pad positions there keep their initial zeros.
                            if (valid) {
                                using GU = Geo<28>;
                                const int64_t p00 = (int64_t)b * GU::S + (2 * y + 1) * GU::Wp + 2 * c;
                                uint8_t* dst = a.out + plane * a.out_ps + (p00 + GU::HALO) * 16;
                                *reinterpret_cast<uint4*>(dst) = o;
                                *reinterpret_cast<uint4*>(dst + 16) = o;
                                *reinterpret_cast<uint4*>(dst + GU::Wp * 16) = o;
                                *reinterpret_cast<uint4*>(dst + GU::Wp * 16 + 16) = o;
                            }
                        } else {
                            *reinterpret_cast<uint4*>(a.out + plane * a.out_ps + (pos + G::HALO) * 16) = o;
                        }
                    }
                    if constexpr (SKIPG) {
                        uint4 o2;
                        uint32_t* ow = &o2.x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int ch = c0 + pj * 8 + 2 * k;
                            const float s0 = __uint_as_float(r2[pj * 8 + 2 * k]) + s_sbias[ch];
                            const float s1 = __uint_as_float(r2[pj * 8 + 2 * k + 1]) + s_sbias[ch + 1];
                            ow[k] = valid ? pack_bf16x2(s0, s1) : 0u;
                        }
                        *reinterpret_cast<uint4*>(a.out2 + plane * a.out2_ps + (pos + G::HALO) * 16) = o2;
                    }
                }
            }
            if constexpr (EPI == EPI_FINAL) {
                if (valid) {
                    const float eps = dot + s_aux[32];  // out conv bias (src/mnist.py:87)
                    const int64_t oi = (int64_t)b * 784 + y * 28 + c;
                    a.fout[oi] = a.fuse_step ? rstep1(sc, xin, eps, zz, add_noise) : eps;
                }
            }
        }
    }

    // ---- teardown --------------------------------------------------------------------------
    __syncwarp();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int W, int CIN, int COUT, int EPI, bool SKIPG>
static int launch_conv(const ConvArgs& a, cudaStream_t st, const char* name) {
    using C = ConvCfg<W, CIN, COUT, SKIPG>;
    auto kern = conv3x3_tc_kernel<W, CIN, COUT, EPI, SKIPG>;
    static bool configured = false;
    if (!configured) {
        TDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        configured = true;
    }
    const int grid = a.nt < num_sms() ? a.nt : num_sms();
    kern<<<grid, C::THREADS, C::SMEM_BYTES, st>>>(a);
    TDM_CHECK_LAUNCH(name);
    return TDM_OK;
}

// ---------------------------------------------------------------------------------------------
// the nine-launch forward
// ---------------------------------------------------------------------------------------------
struct StepArgs {
    int fuse_step = 0;
    const float* z = nullptr;
    const float* betas = nullptr;
    const float* alphas = nullptr;
    const float* sqrt_om = nullptr;
    uint64_t seed = 0, sample_offset = 0;
    uint32_t step_id = 0;
};

// optional per-kernel timing (bench.py's roofline): events recorded between the nine launches
static thread_local cudaEvent_t* g_prof = nullptr;
#define TDM_PROF(i)                                   \
    do {                                              \
        if (g_prof) cudaEventRecord(g_prof[i], st);   \
    } while (0)

static int unet_forward_impl(const uint8_t* wp, const float* x, const int64_t* t, float* fout,
                             uint8_t* ws, int64_t ws_bytes, int64_t batch, const StepArgs& sa,
                             cudaStream_t st) {
    TDM_CHECK_ARG(wp && x && t && fout && ws, "unet_forward: null pointer");
    TDM_CHECK_ARG(batch > 0 && batch <= (1 << 20), "unet_forward: batch %lld out of range", (long long)batch);
    TDM_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0 && (reinterpret_cast<uintptr_t>(wp) & 255) == 0,
                  "unet_forward: workspace and wpack must be 256-byte aligned");
    const UNetWs L = make_ws(batch);
    TDM_CHECK_ARG(ws_bytes >= L.total, "unet_forward: workspace too small (%lld < %lld)",
                  (long long)ws_bytes, (long long)L.total);
    const float* fp = reinterpret_cast<const float*>(wp + WP::flat);
    const int B = (int)batch;
    const int nt28 = (int)L.nt28, nt14 = (int)L.nt14;
    int rc;

    // k1
    TDM_PROF(0);
    rb1_conv1_kernel<<<nt28, 128, 0, st>>>(x, t, fp, ws + L.t1, L.ps28, B);
    TDM_CHECK_LAUNCH("rb1_conv1");

    TDM_PROF(1);
    ConvArgs a{};
    a.t = t;
    a.batch = B;
    // k2: rb1.conv2 -> h1 = cat planes 8..11
    a.in = ws + L.t1; a.in_ps = L.ps28; a.w = wp + WP::rb1_c2; a.bias = fp + P::rb1_c2b;
    a.out = ws + L.cat + 8 * L.ps28; a.out_ps = L.ps28;
    a.x = x; a.aux_w = fp + P::rb1_sw; a.aux_b = fp + P::rb1_sb; a.nt = nt28;
    if ((rc = launch_conv<28, 32, 32, EPI_RES_X, false>(a, st, "rb1_conv2"))) return rc;

    // k3: pool h1 -> p1
    TDM_PROF(2);
    avgpool_kernel<<<dim3(nt14, 4), 128, 0, st>>>(ws + L.cat + 8 * L.ps28, L.ps28, ws + L.p1, L.ps14, B);
    TDM_CHECK_LAUNCH("avgpool");

    // k4: rb2.conv1 (+skip) -> t2, s2
    TDM_PROF(3);
    a = ConvArgs{}; a.t = t; a.batch = B; a.nt = nt14;
    a.in = ws + L.p1; a.in_ps = L.ps14; a.w = wp + WP::rb2_c1; a.bias = fp + P::rb2_c1b;
    a.tw = fp + P::rb2_tw; a.tb = fp + P::rb2_tb; a.sbias = fp + P::rb2_sb;
    a.out = ws + L.t2; a.out_ps = L.ps14; a.out2 = ws + L.s2; a.out2_ps = L.ps14;
    if ((rc = launch_conv<14, 32, 64, EPI_CONV1, true>(a, st, "rb2_conv1"))) return rc;

    // k5: rb2.conv2 + s2 -> h2
    TDM_PROF(4);
    a = ConvArgs{}; a.t = t; a.batch = B; a.nt = nt14;
    a.in = ws + L.t2; a.in_ps = L.ps14; a.w = wp + WP::rb2_c2; a.bias = fp + P::rb2_c2b;
    a.res = ws + L.s2; a.res_ps = L.ps14; a.out = ws + L.h2; a.out_ps = L.ps14;
    if ((rc = launch_conv<14, 64, 64, EPI_RES, false>(a, st, "rb2_conv2"))) return rc;

    // k6: rb3.conv1 -> t3
    TDM_PROF(5);
    a = ConvArgs{}; a.t = t; a.batch = B; a.nt = nt14;
    a.in = ws + L.h2; a.in_ps = L.ps14; a.w = wp + WP::rb3_c1; a.bias = fp + P::rb3_c1b;
    a.tw = fp + P::rb3_tw; a.tb = fp + P::rb3_tb; a.out = ws + L.t3; a.out_ps = L.ps14;
    if ((rc = launch_conv<14, 64, 64, EPI_CONV1, false>(a, st, "rb3_conv1"))) return rc;

    // k7: rb3.conv2 + h2 -> upsampled into cat planes 0..7
    TDM_PROF(6);
    a = ConvArgs{}; a.t = t; a.batch = B; a.nt = nt14;
    a.in = ws + L.t3; a.in_ps = L.ps14; a.w = wp + WP::rb3_c2; a.bias = fp + P::rb3_c2b;
    a.res = ws + L.h2; a.res_ps = L.ps14; a.out = ws + L.cat; a.out_ps = L.ps28;
    if ((rc = launch_conv<14, 64, 64, EPI_RES_UP, false>(a, st, "rb3_conv2"))) return rc;

    // k8: rb4.conv1 (+skip) -> t4, s4
    TDM_PROF(7);
    a = ConvArgs{}; a.t = t; a.batch = B; a.nt = nt28;
    a.in = ws + L.cat; a.in_ps = L.ps28; a.w = wp + WP::rb4_c1; a.bias = fp + P::rb4_c1b;
    a.tw = fp + P::rb4_tw; a.tb = fp + P::rb4_tb; a.sbias = fp + P::rb4_sb;
    a.out = ws + L.t4; a.out_ps = L.ps28; a.out2 = ws + L.s4; a.out2_ps = L.ps28;
    if ((rc = launch_conv<28, 96, 32, EPI_CONV1, true>(a, st, "rb4_conv1"))) return rc;

    // k9: rb4.conv2 + s4, out conv, optional reverse step
    TDM_PROF(8);
    a = ConvArgs{}; a.t = t; a.batch = B; a.nt = nt28;
    a.in = ws + L.t4; a.in_ps = L.ps28; a.w = wp + WP::rb4_c2; a.bias = fp + P::rb4_c2b;
    a.res = ws + L.s4; a.res_ps = L.ps28; a.aux_w = fp + P::out_w; a.aux_b = fp + P::out_b;
    a.fout = fout; a.x = sa.fuse_step ? x : nullptr;
    a.fuse_step = sa.fuse_step; a.z = sa.z; a.betas = sa.betas; a.alphas = sa.alphas;
    a.sqrt_om = sa.sqrt_om; a.seed = sa.seed; a.sample_offset = sa.sample_offset; a.step_id = sa.step_id;
    if ((rc = launch_conv<28, 32, 32, EPI_FINAL, false>(a, st, "rb4_conv2"))) return rc;
    TDM_PROF(9);
    return TDM_OK;
}

}  // namespace tdm

using namespace tdm;

extern "C" int64_t tdm_unet_param_count(void) { return P::count; }
extern "C" int64_t tdm_unet_wpack_bytes(void) { return WP::total; }
extern "C" int64_t tdm_unet_workspace_bytes(int64_t batch, int for_backward) {
    (void)for_backward;
    if (batch <= 0) return 0;
    return make_ws(batch).total;
}

extern "C" int tdm_unet_debug_layout(int64_t batch, int64_t* host_out14) {
    TDM_CHECK_ARG(batch > 0 && host_out14, "tdm_unet_debug_layout: bad arguments");
    const UNetWs w = make_ws(batch);
    const int64_t v[14] = {w.nt28, w.nt14, w.ps28, w.ps14, w.t1, w.cat, w.p1,
                           w.t2,   w.s2,   w.h2,   w.t3,   w.t4, w.s4,  w.total};
    for (int i = 0; i < 14; ++i) host_out14[i] = v[i];
    return TDM_OK;
}

extern "C" int tdm_unet_pack_weights(const float* flat_params, void* wpack, void* stream) {
    TDM_CHECK_ARG(flat_params && wpack, "tdm_unet_pack_weights: null pointer");
    pack_weights_kernel<<<dim3(32, 9), 256, 0, (cudaStream_t)stream>>>(
        flat_params, reinterpret_cast<uint8_t*>(wpack));
    TDM_CHECK_LAUNCH("tdm_unet_pack_weights");
    return TDM_OK;
}

extern "C" int tdm_unet_forward(const void* wpack, const float* x, const int64_t* t, float* eps_out,
                                void* workspace, int64_t workspace_bytes, int64_t batch, void* stream) {
    StepArgs sa;
    return unet_forward_impl(reinterpret_cast<const uint8_t*>(wpack), x, t, eps_out,
                             reinterpret_cast<uint8_t*>(workspace), workspace_bytes, batch, sa,
                             (cudaStream_t)stream);
}

extern "C" int tdm_unet_p_sample(const void* wpack, const float* x_in, const int64_t* t, const float* z,
                                 const float* betas, const float* alphas, const float* sqrt_om_acp,
                                 float* x_out, void* workspace, int64_t workspace_bytes, int64_t batch,
                                 int n_steps, uint64_t seed, uint64_t sample_offset, uint32_t step_id,
                                 void* stream) {
    TDM_CHECK_ARG(betas && alphas && sqrt_om_acp, "tdm_unet_p_sample: null schedule table");
    TDM_CHECK_ARG(n_steps > 0, "tdm_unet_p_sample: bad n_steps");
    StepArgs sa;
    sa.fuse_step = 1; sa.z = z; sa.betas = betas; sa.alphas = alphas; sa.sqrt_om = sqrt_om_acp;
    sa.seed = seed; sa.sample_offset = sample_offset; sa.step_id = step_id;
    return unet_forward_impl(reinterpret_cast<const uint8_t*>(wpack), x_in, t, x_out,
                             reinterpret_cast<uint8_t*>(workspace), workspace_bytes, batch, sa,
                             (cudaStream_t)stream);
}

extern "C" int tdm_unet_profile_p_sample(const void* wpack, const float* x_in, const int64_t* t,
                                         const float* betas, const float* alphas,
                                         const float* sqrt_om_acp, float* x_out, void* workspace,
                                         int64_t workspace_bytes, int64_t batch, uint64_t seed,
                                         float* host_ms9, void* stream) {
    TDM_CHECK_ARG(host_ms9, "tdm_unet_profile_p_sample: null output");
    cudaEvent_t ev[10];
    for (int i = 0; i < 10; ++i) TDM_CHECK_CUDA(cudaEventCreate(&ev[i]));
    StepArgs sa;
    sa.fuse_step = 1; sa.betas = betas; sa.alphas = alphas; sa.sqrt_om = sqrt_om_acp; sa.seed = seed;
    g_prof = ev;
    const int rc = unet_forward_impl(reinterpret_cast<const uint8_t*>(wpack), x_in, t, x_out,
                                     reinterpret_cast<uint8_t*>(workspace), workspace_bytes, batch, sa,
                                     (cudaStream_t)stream);
    g_prof = nullptr;
    int ret = rc;
    if (rc == TDM_OK) {
        cudaError_t e = cudaEventSynchronize(ev[9]);
        if (e != cudaSuccess) {
            set_error("tdm_unet_profile_p_sample: %s", cudaGetErrorString(e));
            ret = TDM_ERR_CUDA;
        } else {
            for (int i = 0; i < 9; ++i) cudaEventElapsedTime(&host_ms9[i], ev[i], ev[i + 1]);
        }
    }
    for (int i = 0; i < 10; ++i) cudaEventDestroy(ev[i]);
    return ret;
}
