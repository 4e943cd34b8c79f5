// ffn_tc.cuh — the whole feed-forward half of a post-LN encoder layer in ONE kernel (width 256):
//
//     out = LayerNorm( h + b2 + relu(h W1^T + b1) W2^T )          nn.TransformerEncoderLayer, FFN 2048
//
// The 2048-wide intermediate never leaves the SM.  Per tile of 128 token rows, for each chunk c of
// 128 hidden units:
//     G1(c)  acc1[c%2] (TMEM, 128 cols)  = A (smem, resident)    x W1[128c:128c+128, :]^T   16 MMAs, N=128
//     E1(c)  P (smem, bf16 K-major planes) = relu(acc1 + b1)     epilogue group c%2
//     G2(c)  acc2 (TMEM, 256 cols)      += P                     x W2[:, 128c:128c+128]^T    8 MMAs, N=256
// W1 / W2 stream through a ring of 32 KB stages (64-wide K blocks, 8 bulk copies each).  Every tile needs
// the same 2 MB of weights, which makes the kernel L2->SMEM bandwidth bound; so CTAs run as clusters of
// two that walk the weight stream in lock step: each CTA issues half of every K block's copies as a
// cluster MULTICAST (both CTAs receive all eight planes) and a stage is released only when both CTAs'
// MMAs have consumed it (commit multicast onto both empty barriers).  The MMA thread
// issues G1 two chunks ahead of G2 so the tensor pipe stays busy while E1 converts.  After the 16th
// chunk epilogue group 0 adds bias + residual, LayerNorms the row (two passes over acc2, the pre-LN
// row parked in TMEM) and — on the last layer of a sampling step — applies the reverse-step update
// and the next timestep's time embedding (src/shakespeare.py:343-352, 116-118).
#pragma once
#include "common.cuh"
#include "diffusion_math.cuh"
#include "tc05.cuh"

namespace tdm {

constexpr int kFfnD = 256;        // model width handled by this kernel
constexpr int kFfnH = 2048;       // hidden width
constexpr int kFfnC = 128;        // hidden chunk
constexpr int kFfnChunks = kFfnH / kFfnC;
constexpr int kFfnStage = 32768;
constexpr int kFfnStages = 4;
constexpr int kFfnA = 128 * kFfnD * 2;          // 64 KB
constexpr int kFfnP = 128 * kFfnC * 2;          // 32 KB
constexpr int kFfnLn = 2 * 128 * 8;             // LayerNorm partial sums exchanged between the two epilogue groups
constexpr int kFfnSmem = kFfnA + kFfnStages * kFfnStage + kFfnP + kFfnLn + 1024;   // = 227 KB exactly
constexpr int kFfnThreads = 64 + 256;
constexpr int kFfnCluster = 2;    // CTAs sharing one multicast weight stream

struct FfnArgs {
    const uint8_t* a;       // bf16 planes [32][Mp][8] (LayerNorm-1 output)
    const uint8_t* w1;      // linear1.weight (2048, 256) as bf16 STAGES [chunk 16][K block 4][plane 8][128 rows][8] (tdm_pack_ffn_weights)
    const float* b1;        // [2048]
    const uint8_t* w2;      // linear2.weight (256, 2048) as bf16 STAGES [chunk 16][K block 2][plane 8][256 rows][8]
    const float* b2;        // [256]
    const uint8_t* res;     // fp32 planes [64][Mp][4] residual (same tensor as `a`, in fp32)
    const float* gamma;     // norm2
    const float* beta;
    float ln_eps;
    uint8_t* out_f32;       // fp32 planes (may alias res: a thread reads its row before writing it)
    uint8_t* out_bf16;      // bf16 planes (may alias a: rows are partitioned by tile and a tile's rows are copied
                            // to smem before its epilogue writes them)
    int64_t ps;             // plane stride of all row-plane tensors (Mp*16)
    int M, Mp, L;
    // fused reverse step (last layer of a sampling step)
    int fuse_step;
    uint8_t* state;
    const int64_t* t;
    const float* z;
    const float* betas;
    const float* alphas;
    const float* sqrt_om;
    const float* tw;
    const float* tb;
    uint64_t seed, sample_offset;
    uint32_t step_id;
};

#ifdef TDM_EXP_TIMELINE
// development aid (tools/ffn_timeline.py): globaltimer stamps of block 0, second tile: [what][chunk]
__device__ unsigned long long g_ffn_tl[9][16];
__device__ __forceinline__ void ffn_stamp(int what, int c) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_ffn_tl[what][c] = t;
}
#define FTL(what, c) do { if (blockIdx.x == 0 && ti == 1) ffn_stamp(what, c); } while (0)
#else
#define FTL(what, c)
#endif

__global__ void __launch_bounds__(kFfnThreads, 1) ffn_tc_kernel(const FfnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sR = smem + kFfnA;
    uint8_t* sP = sR + kFfnStages * kFfnStage;
    float2* s_ln = reinterpret_cast<float2*>(sP + kFfnP);   // [group][row of the tile]: (sum, sum of squares) over the group's columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + kFfnP + kFfnLn);
    uint64_t* bar_full = bars;                       // [stages]
    uint64_t* bar_empty = bars + kFfnStages;         // [stages]
    uint64_t* bar_a_full = bar_empty + kFfnStages;
    uint64_t* bar_a_empty = bar_a_full + 1;
    uint64_t* bar_acc1_full = bar_a_empty + 1;       // [2]
    uint64_t* bar_acc1_empty = bar_acc1_full + 2;    // [2]
    uint64_t* bar_p_full = bar_acc1_empty + 2;
    uint64_t* bar_p_empty = bar_p_full + 1;
    uint64_t* bar_acc2_full = bar_p_empty + 1;
    uint64_t* bar_acc2_empty = bar_acc2_full + 1;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc2_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kFfnStages; ++i) {
            mbar_init(bar_full + i, 1);
            mbar_init(bar_empty + i, kFfnCluster);   // one commit from every CTA of the cluster
        }
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_empty, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_acc1_full + i, 1);
            mbar_init(bar_acc1_empty + i, 4);
        }
        mbar_init(bar_p_full, 4);
        mbar_init(bar_p_empty, 1);
        mbar_init(bar_acc2_full, 1);
        mbar_init(bar_acc2_empty, 8);   // both epilogue groups read the tile accumulator
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(s_tmem);
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();   // every CTA's mbarriers exist before any peer multicasts into / arrives on them
    tc_fence_after_sync();
    pdl_wait();                 // PDL (common.cuh): nothing above touches global memory
    pdl_launch_dependents();
    const uint32_t tmem_base = *s_tmem;
    const uint32_t crank = cluster_ctarank();
    const int cluster_id = blockIdx.x / kFfnCluster, n_clusters = gridDim.x / kFfnCluster;
    constexpr uint16_t kAll = (1u << kFfnCluster) - 1;
    const uint32_t t_acc1 = tmem_base;           // two buffers of 128 columns
    const uint32_t t_acc2 = tmem_base + 256;     // 256 columns
    const int m_tiles = a.Mp / 128;

    if (warp == 0) {
        // ===== producer: A once per tile, then the weight K-blocks in the order the MMA thread consumes them =====
        int kit = 0, ti = 0;
        auto put = [&](bool is_w2, int c, int kb) {
            const int s = kit % kFfnStages;
            const uint32_t ph = (kit / kFfnStages) & 1;
            ++kit;
            if (lane == 0) {
                mbar_wait(bar_empty + s, ph ^ 1);
                mbar_arrive_expect_tx(bar_full + s, is_w2 ? 32768 : 16384);
            }
            __syncwarp();
            uint8_t* st = sR + s * kFfnStage;
            // The weights are stored stage by stage (tdm_pack_ffn_weights): K block kb of chunk c is ONE contiguous 16 KB (W1) or
            // 32 KB (W2) block in exactly the order the ring slot holds it (eight planes), so each CTA of the pair fetches
            // its contiguous half with a single multicast copy.  In the plane layout a stage was eight 2-4 KB copies, 24 per
            // chunk and CTA, and the copy engine's issue rate (one small bulk copy per ~50-150 ns) - not its bandwidth -
            // paced the weight stream.
            if (lane == 0) {
                if (is_w2) {
                    const uint32_t half = 32768 / kFfnCluster;
                    bulk_g2s_multicast(st + crank * half, a.w2 + ((int64_t)(c * 2 + kb) * 32768) + crank * half, half, bar_full + s, kAll);
                } else {
                    const uint32_t half = 16384 / kFfnCluster;
                    bulk_g2s_multicast(st + crank * half, a.w1 + ((int64_t)(c * 4 + kb) * 16384) + crank * half, half, bar_full + s, kAll);
                }
            }
        };
        for (int mt = cluster_id * kFfnCluster + (int)crank; mt < m_tiles; mt += n_clusters * kFfnCluster, ++ti) {
            if (lane == 0) {
                mbar_wait(bar_a_empty, (ti & 1) ^ 1);
                mbar_arrive_expect_tx(bar_a_full, kFfnA);
            }
            __syncwarp();
            bulk_g2s(sA + lane * 2048, a.a + (int64_t)lane * a.ps + (int64_t)mt * 2048, 2048, bar_a_full);
            for (int kb = 0; kb < 4; ++kb) put(false, 0, kb);
            for (int kb = 0; kb < 4; ++kb) put(false, 1, kb);
            for (int c = 0; c < kFfnChunks; ++c) {
                for (int kb = 0; kb < 2; ++kb) put(true, c, kb);
                if (c + 2 < kFfnChunks)
                    for (int kb = 0; kb < 4; ++kb) put(false, c + 2, kb);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected thread runs the whole issue loop (conv_tc.cuh) =====
        constexpr uint32_t idesc1 = make_idesc_bf16(128, kFfnC);
        constexpr uint32_t idesc2 = make_idesc_bf16(128, kFfnD);
        const uint64_t a_base = make_smem_desc(smem_u32(sA), 2048, 128);
        const uint64_t p_base = make_smem_desc(smem_u32(sP), 2048, 128);
        if (elect_one()) {
        int kit = 0, ti = 0;
        auto g1 = [&](int cg) {   // cg: chunk counter across tiles (buffer cg&1, use index cg>>1)
            const int buf = cg & 1;
            mbar_wait(bar_acc1_empty + buf, ((cg >> 1) & 1) ^ 1);
            tc_fence_after_sync();
            for (int kb = 0; kb < 4; ++kb, ++kit) {
                const int s = kit % kFfnStages;
                mbar_wait(bar_full + s, (kit / kFfnStages) & 1);
                tc_fence_after_sync();
                const uint64_t b_base = make_smem_desc(smem_u32(sR + s * kFfnStage), 2048, 128);
                const uint32_t acc_flag = kb != 0;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma_bf16(t_acc1 + buf * kFfnC, desc_add(a_base, (kb * 8 + 2 * ks) * 2048), desc_add(b_base, (2 * ks) * 2048),
                                    idesc1, ks != 0 ? 1u : acc_flag);
                umma_commit_multicast(bar_empty + s, kAll);
            }
            umma_commit(bar_acc1_full + buf);
        };
        for (int mt = cluster_id * kFfnCluster + (int)crank; mt < m_tiles; mt += n_clusters * kFfnCluster, ++ti) {
            mbar_wait(bar_a_full, ti & 1);
            tc_fence_after_sync();
            const int cg0 = ti * kFfnChunks;
            g1(cg0);
            g1(cg0 + 1);
            for (int c = 0; c < kFfnChunks; ++c) {
                const int cg = cg0 + c;
                if (c == 0) mbar_wait(bar_acc2_empty, (ti & 1) ^ 1);
                FTL(0, c);                       // MMA thread: about to wait for P(c)
                mbar_wait(bar_p_full, cg & 1);
                tc_fence_after_sync();
                FTL(1, c);                       // MMA thread: P(c) seen, G2(c) starts
                for (int kb = 0; kb < 2; ++kb, ++kit) {
                    const int s = kit % kFfnStages;
                    mbar_wait(bar_full + s, (kit / kFfnStages) & 1);
                    tc_fence_after_sync();
                    const uint64_t b_base = make_smem_desc(smem_u32(sR + s * kFfnStage), 4096, 128);
                    const uint32_t acc_flag = (c | kb) != 0;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_bf16(t_acc2, desc_add(p_base, (kb * 8 + 2 * ks) * 2048), desc_add(b_base, (2 * ks) * 4096), idesc2,
                                        ks != 0 ? 1u : acc_flag);
                    umma_commit_multicast(bar_empty + s, kAll);
                }
                FTL(2, c);                       // MMA thread: G2(c) issued
                umma_commit(bar_p_empty);                 // P may be overwritten once these MMAs retire
                if (c == kFfnChunks - 1) umma_commit(bar_acc2_full);
                if (c + 2 < kFfnChunks) {
                    g1(cg + 2);
                    FTL(3, c);                   // MMA thread: G1(c+2) issued
                    if (c + 2 == kFfnChunks - 1) umma_commit(bar_a_empty);   // last G1 of the tile: A may be reloaded
                }
            }
        }
        }
        __syncwarp();
    } else {
        // ===== epilogue groups: group g converts chunks c = g (mod 2) and finishes columns [128g, 128g+128) of the tile =====
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        int ti = 0;
        for (int mt = cluster_id * kFfnCluster + (int)crank; mt < m_tiles; mt += n_clusters * kFfnCluster, ++ti) {
            const int row = mt * 128 + q * 32 + lane;
            const bool rvalid = row < a.M;
            for (int c = grp; c < kFfnChunks; c += 2) {
                const int cg = ti * kFfnChunks + c;
                mbar_wait(bar_acc1_full + grp, (cg >> 1) & 1);
                tc_fence_after_sync();
                if (q == 0 && lane == 0) FTL(4, c);   // E1: G1(c) result seen
                uint4 pk[kFfnC / 8];
#pragma unroll
                for (int c0 = 0; c0 < kFfnC; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(t_acc1 + lane_base + grp * kFfnC + c0, r);
                    const float bias_l = __ldg(a.b1 + c * kFfnC + c0 + lane);   // coalesced; broadcast by shuffle below
                    tmem_ld_wait();
                    if (c0 + 32 == kFfnC) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_acc1_empty + grp);
                    }
#pragma unroll
                    for (int pj = 0; pj < 4; ++pj) {
                        float v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            v[k] = fmaxf(__uint_as_float(r[pj * 8 + k]) + __shfl_sync(0xffffffffu, bias_l, pj * 8 + k), 0.f);
                        pk[c0 / 8 + pj] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                     pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    }
                }
                if (q == 0 && lane == 0) FTL(5, c);   // E1: converted, about to wait for P to be free
                mbar_wait(bar_p_empty, (cg & 1) ^ 1);          // G2 of the previous chunk has consumed P
                if (q == 0 && lane == 0) FTL(6, c);   // E1: P free
#pragma unroll
                for (int pl = 0; pl < kFfnC / 8; ++pl)
                    *reinterpret_cast<uint4*>(sP + pl * 2048 + (q * 32 + lane) * 16) = pk[pl];
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_p_full);
                if (q == 0 && lane == 0) FTL(7, c);   // E1: P(c) written
            }
            // ---- tile epilogue: bias + residual, LayerNorm, optional reverse step.  The two groups split the
            //      256 columns (same TMEM lanes, i.e. rows; disjoint columns) and exchange the row statistics
            //      through smem: with one group the last layer's Philox + reverse-step epilogue was 67 us
            //      of a 184 us kernel (profiles/r01_ncu_launches_text_B512.csv). ----
            if (q == 0 && lane == 0) FTL(8, 0 + 8 * grp);   // tile epilogue: E1 work of this group done, waiting for acc2
            mbar_wait(bar_acc2_full, ti & 1);
            tc_fence_after_sync();
            if (q == 0 && lane == 0) FTL(8, 1 + 8 * grp);   // acc2 complete
            const uint32_t taddr = t_acc2 + lane_base;
            const int col_lo = grp * (kFfnD / 2), col_hi = col_lo + kFfnD / 2;
            float ln_sum = 0.f, ln_sq = 0.f;
#pragma unroll 1
            for (int c0 = col_lo; c0 < col_hi; c0 += 32) {
                uint32_t r[32], vb[32];
                tmem_ld32(taddr + c0, r);
                const float bias_l = __ldg(a.b2 + c0 + lane);
                tmem_ld_wait();
#pragma unroll
                for (int pj = 0; pj < 8; ++pj) {
                    float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rvalid) rv = *reinterpret_cast<const float4*>(a.res + (int64_t)(c0 / 4 + pj) * a.ps + (int64_t)row * 16);
                    const float v0 = __uint_as_float(r[pj * 4 + 0]) + __shfl_sync(0xffffffffu, bias_l, pj * 4 + 0) + rv.x;
                    const float v1 = __uint_as_float(r[pj * 4 + 1]) + __shfl_sync(0xffffffffu, bias_l, pj * 4 + 1) + rv.y;
                    const float v2 = __uint_as_float(r[pj * 4 + 2]) + __shfl_sync(0xffffffffu, bias_l, pj * 4 + 2) + rv.z;
                    const float v3 = __uint_as_float(r[pj * 4 + 3]) + __shfl_sync(0xffffffffu, bias_l, pj * 4 + 3) + rv.w;
                    ln_sum += (v0 + v1) + (v2 + v3);
                    ln_sq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, ln_sq))));
                    vb[pj * 4 + 0] = __float_as_uint(v0);
                    vb[pj * 4 + 1] = __float_as_uint(v1);
                    vb[pj * 4 + 2] = __float_as_uint(v2);
                    vb[pj * 4 + 3] = __float_as_uint(v3);
                }
                tmem_st32(taddr + c0, vb);
            }
            tmem_st_wait();
            if (q == 0 && lane == 0) FTL(8, 2 + 8 * grp);   // pass 1 done
            s_ln[grp * 128 + q * 32 + lane] = make_float2(ln_sum, ln_sq);
            asm volatile("bar.sync 1, 256;" ::: "memory");   // the eight epilogue warps
            if (q == 0 && lane == 0) FTL(8, 3 + 8 * grp);   // statistics exchanged
            {
                const float2 o = s_ln[(grp ^ 1) * 128 + q * 32 + lane];
                ln_sum += o.x;
                ln_sq += o.y;
            }
            const float mean = ln_sum * (1.0f / kFfnD);
            const float rstd = rsqrtf(fmaxf(ln_sq * (1.0f / kFfnD) - mean * mean, 0.f) + a.ln_eps);
            StepCoef sc{};
            bool add_noise = false;
            float tsn = 0.f;
            int64_t tb_ = 0;
            const int bidx = row / a.L, l = row - bidx * a.L;
            if (a.fuse_step && rvalid) {
                add_noise = __ldg(a.t) != 0;   // the reference branches on t[0] (src/shakespeare.py:349)
                tb_ = __ldg(a.t + bidx);
                sc = step_coef(tb_, a.betas, a.alphas, a.sqrt_om);
                tsn = (float)(tb_ - 1) / 1000.0f;
            }
#pragma unroll 1
            for (int c0 = col_lo; c0 < col_hi; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                if (c0 + 32 == col_hi) {
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acc2_empty);
                }
                if (!rvalid) continue;
#pragma unroll
                for (int pj = 0; pj < 4; ++pj) {
                    float y[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int n = c0 + pj * 8 + k;
                        y[k] = (__uint_as_float(r[pj * 8 + k]) - mean) * rstd * __ldg(a.gamma + n) + __ldg(a.beta + n);
                    }
                    if (a.fuse_step) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int cc = c0 + pj * 8 + h * 4;
                            float4* sp = reinterpret_cast<float4*>(a.state + (int64_t)(cc / 4) * a.ps + (int64_t)row * 16);
                            const float4 xv = *sp;
                            float4 zz = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (add_noise) {
                                if (a.z) {
                                    const float* zr = a.z + (int64_t)row * kFfnD + cc;
                                    zz = make_float4(__ldg(zr), __ldg(zr + 1), __ldg(zr + 2), __ldg(zr + 3));
                                } else {
                                    zz = philox_normal4(a.seed, a.sample_offset + (uint64_t)bidx,
                                                        (uint32_t)((l * kFfnD + cc) >> 2), a.step_id + (uint32_t)tb_, kDomainReverse);
                                }
                            }
                            float4 xn;
                            xn.x = rstep1(sc, xv.x, y[h * 4 + 0], zz.x, add_noise);
                            xn.y = rstep1(sc, xv.y, y[h * 4 + 1], zz.y, add_noise);
                            xn.z = rstep1(sc, xv.z, y[h * 4 + 2], zz.z, add_noise);
                            xn.w = rstep1(sc, xv.w, y[h * 4 + 3], zz.w, add_noise);
                            *sp = xn;
                            y[h * 4 + 0] = xn.x + fmaf(__ldg(a.tw + cc + 0), tsn, __ldg(a.tb + cc + 0));
                            y[h * 4 + 1] = xn.y + fmaf(__ldg(a.tw + cc + 1), tsn, __ldg(a.tb + cc + 1));
                            y[h * 4 + 2] = xn.z + fmaf(__ldg(a.tw + cc + 2), tsn, __ldg(a.tb + cc + 2));
                            y[h * 4 + 3] = xn.w + fmaf(__ldg(a.tw + cc + 3), tsn, __ldg(a.tb + cc + 3));
                        }
                    }
                    *reinterpret_cast<float4*>(a.out_f32 + (int64_t)(c0 / 4 + 2 * pj) * a.ps + (int64_t)row * 16) = make_float4(y[0], y[1], y[2], y[3]);
                    *reinterpret_cast<float4*>(a.out_f32 + (int64_t)(c0 / 4 + 2 * pj + 1) * a.ps + (int64_t)row * 16) = make_float4(y[4], y[5], y[6], y[7]);
                    *reinterpret_cast<uint4*>(a.out_bf16 + (int64_t)(c0 / 8 + pj) * a.ps + (int64_t)row * 16) =
                        make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
                }
            }
            if (q == 0 && lane == 0) FTL(8, 4 + 8 * grp);   // pass 2 done
        }
    }
    __syncwarp();
    tc_fence_before_sync();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while a peer may still multicast into it or arrive on its barriers
    tc_fence_after_sync();
    if (warp == 2) tmem_dealloc<512>(tmem_base);
}

static int launch_ffn(const FfnArgs& a, cudaStream_t st) {
    TDM_SET_MAX_DYN_SMEM(ffn_tc_kernel, kFfnSmem);
    const int m_tiles = a.Mp / 128;
    TDM_CHECK_ARG(m_tiles % kFfnCluster == 0, "ffn_fused: row tiles (%d) must be a multiple of the cluster size", m_tiles);
    int grid = m_tiles < num_sms() ? m_tiles : num_sms();
    grid -= grid % kFfnCluster;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kFfnThreads);
    cfg.dynamicSmemBytes = kFfnSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kFfnCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    TDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ffn_tc_kernel, a));
    TDM_CHECK_LAUNCH("ffn_fused");
    return TDM_OK;
}

}  // namespace tdm
