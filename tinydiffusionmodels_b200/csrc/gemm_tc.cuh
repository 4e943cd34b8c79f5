// gemm_tc.cuh — persistent warp-specialised tcgen05 GEMM for the text path:
//     C[M, N] = A[M, K] · W[N, K]^T  (+ bias, fused epilogue)
// Operands are bf16 "planes" — [K/8][rows][8], 16-byte rows — the same un-swizzled K-major
// core-matrix layout the UNet kernels use, so a 128x64 A block or a 256x64 W block is eight
// contiguous 1-D bulk async copies (no tensor maps).  Tile 128 x 256 x 64, 4 smem stages,
// two TMEM accumulator stages (2 x 256 columns) each drained by its own group of 4 epilogue warps.
//
// Epilogues (thread = output row, all N columns of the tile in registers 32 at a time):
//   GE_BF16    bias (+ReLU) -> bf16 planes                      (QKV, FFN1 of nn.TransformerEncoderLayer)
//   GE_RES_F32 bias + fp32 residual -> fp32 planes (pre-LN)     (attention out-proj, FFN2)
//   GE_RES_LN  N == 256 only: bias + fp32 residual, then LayerNorm of the whole row in the epilogue
//              (row = one thread; the pre-LN values are parked back in the accumulator's TMEM columns
//              between the statistics pass and the normalise pass) -> fp32 + bf16 planes
//   GE_LOGITS  bias (+ row scale) -> fp32 row-major [M][n_valid]: the logits LearnedRounding.forward returns
//              (src/shakespeare.py:93-102); the samplers never take this path
//   GE_ARGMAX  running (max, argmax) over N, optionally mixed with AR logits — the logits are
//              never written (src/shakespeare.py:389-390, 398-401, 451-467)
// Training (text_train.cu; src/shakespeare.py:221-250) adds a K split for GE_LOGITS (few output tiles, long
// reductions: partial s of an item goes to its own fp32 slab, summed by a small kernel afterwards), ReLU / an fp32
// addend for GE_LOGITS, and the two halves of the fused Linear(dim, V) + cross-entropy:
//   GE_LSE     online (max, sum exp) per row over the item's vocabulary range + the target's logit; logits never written
//   GE_DLOGITS (softmax - onehot) * scale -> bf16 planes (the A operand of the two gradient GEMMs)
#pragma once
#include "common.cuh"
#include "tc05.cuh"

namespace tdm {

enum : int { GE_BF16 = 0, GE_RES_F32 = 1, GE_ARGMAX = 2, GE_RES_LN = 3, GE_LOGITS = 4, GE_LSE = 5, GE_DLOGITS = 6 };

constexpr int kBM = 128, kBN = 256, kBK = 64;
constexpr int kGemmStages = 4;
constexpr int kGemmStageA = kBM * kBK * 2;   // 16 KB
constexpr int kGemmStageB = kBN * kBK * 2;   // 32 KB
constexpr int kGemmStage = kGemmStageA + kGemmStageB;
// GE_ARGMAX with AR logits: each epilogue warp transposes 32 rows x 32 columns of the fp32 AR logits through its own
// padded shared-memory tile (coalesced 128-byte row reads in, one row per lane out)
constexpr int kArTileFloats = 32 * 33;
constexpr int kGemmSmemBase = kGemmStages * kGemmStage + 1024;
constexpr int kGemmSmem = kGemmSmemBase + 8 * kArTileFloats * 4;
constexpr int kGemmThreads = 64 + 2 * 128;

struct GemmArgs {
    const uint8_t* a;      // bf16 planes [K/8][a_rows][8]
    int64_t a_ps;          // plane stride, bytes
    const uint8_t* w;      // bf16 planes [K/8][Np][8]
    int64_t w_ps;
    const float* bias;     // [N] or null
    int M;                 // valid rows
    int Mp;                // rows rounded up to 128 (tiles)
    int N;                 // columns, multiple of 256 (padded)
    int n_valid;           // columns that exist (argmax ignores the padding)
    int K;                 // multiple of 64
    int relu;
    uint8_t* out_bf16;     // GE_BF16: bf16 planes [N/8][Mp][8]
    int64_t ob_ps;
    const uint8_t* res;    // GE_RES_F32: fp32 planes [N/4][Mp][4]
    int64_t res_ps;
    uint8_t* out_f32;      // GE_RES_F32: fp32 planes
    int64_t of_ps;
    // GE_RES_LN: LayerNorm affine + outputs (fp32 planes -> out_f32, bf16 planes -> out_bf16)
    const float* gamma;
    const float* beta;
    float ln_eps;
    // GE_ARGMAX
    const float* row_scale;  // [M] multiplies the dot product (cosine: 1/||x||), or null
    const float* ar;         // [M][ar_ld] fp32 AR logits, or null
    int64_t ar_ld;
    float alpha, inv_temp;
    float* part_val;         // [2*nsplit][Mp]
    int64_t* part_idx;
    int nsplit;              // work items per row tile (== N/256 for the plain epilogues)
    // GE_LOGITS
    float* logits;           // [M][logits_ld] fp32 row-major
    int64_t logits_ld;
    const float* logits_add; // optional fp32 [M][logits_ld] added to the result (gradient of a residual branch)
    // K split (GE_LOGITS only): items = row tiles x nsplit x ksplit; split s > 0 writes its partial sums (no bias, no
    // addend) to logits + s * split_stride
    int ksplit;              // 0 or 1: the whole K in one item
    int64_t split_stride;    // floats
    // GE_LSE / GE_DLOGITS: the fused Linear + cross-entropy (F.cross_entropy of src/shakespeare.py:241)
    const int64_t* target;   // [M] class index per row
    float* part_sum;         // GE_LSE: [2*nsplit][Mp] sum exp(v - part_val); part_val holds the running maxima
    float* tgt_logit;        // GE_LSE: [M] logit of the target class
    const float* lse;        // GE_DLOGITS: [M] log-sum-exp per row
    const float* dev_scale;  // GE_DLOGITS: device scalar multiplying the gradient (the rounding-loss weight), or null
    float scale;             // GE_DLOGITS: host factor (1 / rows)
    int pair;                // set by launch_gemm: CTAs run as clusters of two that share one multicast W stream
    uint8_t* out_bf16_t;     // GE_DLOGITS: the same values transposed, bf16 planes along the ROW index [Mp/8][obt_rows][8]
    int64_t obt_rows;        //             (the A operand of dW = dlogits^T . X), or null
};

// 2^x on the MUFU (ex2.approx.ftz: 2^-inf = +0, no range fix-up code)
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

#ifdef TDM_EXP_TIMELINE
// development aid (tools/gemm_timeline.py): globaltimer stamps of block 0's phases
__device__ unsigned long long g_gemm_tl[16];
__device__ __forceinline__ void tl_stamp(int i) {
    if (blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_gemm_tl[i] = t;
    }
}
#define TL(i) tl_stamp(i)
#define TL0(i) do { if (lane == 0) tl_stamp(i); } while (0)
#else
#define TL(i)
#define TL0(i)
#endif

// one work item of the persistent loop: a row tile, a range of column tiles and a range of K blocks
struct GemmItem {
    int mt, sp, ks, n0, n1, kb0, kb1;
};
__device__ __forceinline__ GemmItem gemm_item(const GemmArgs& a, int item, int n_tiles, int kblocks) {
    GemmItem w;
    const int ksplit = a.ksplit > 1 ? a.ksplit : 1;
    const int rest = item / ksplit;
    w.ks = item - rest * ksplit;
    // column-split major, row tile minor: the CTAs that run at the same time work on DIFFERENT row tiles of the SAME column
    // range, so that range of W (V / nsplit rows: 8.5 MB of the 131 MB rounding matrix at nsplit = 15) stays in L2 while
    // every row tile passes over it.  Row-tile major, the ~10 row tiles in flight streamed all of W concurrently and
    // ncu counted 4.5 GB of DRAM reads for 147 MB of operands (profiles/r02_ncu_text_gemm_metrics.csv).
    const int m_tiles = a.Mp / kBM;
    w.sp = rest / m_tiles;
    w.mt = rest - w.sp * m_tiles;
    w.n0 = (int)((int64_t)w.sp * n_tiles / a.nsplit);
    w.n1 = (int)((int64_t)(w.sp + 1) * n_tiles / a.nsplit);
    w.kb0 = (int)((int64_t)w.ks * kblocks / ksplit);
    w.kb1 = (int)((int64_t)(w.ks + 1) * kblocks / ksplit);
    return w;
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_kernel(const GemmArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kGemmStages * kGemmStage);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + kGemmStages;
    uint64_t* bar_accf = bar_empty + kGemmStages;
    uint64_t* bar_acce = bar_accf + 2;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acce + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) TL(0);
    if (threadIdx.x == 0) {
        for (int i = 0; i < kGemmStages; ++i) {
            mbar_init(bar_full + i, 1);
            mbar_init(bar_empty + i, a.pair ? a.pair : 1);   // clustered: one commit from each CTA of the cluster
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_accf + i, 1);
            mbar_init(bar_acce + i, 4);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<512>(s_tmem);
    tc_fence_before_sync();
    __syncthreads();
    // Paired launch (launch_gemm): the two CTAs of a cluster take consecutive work items - with the column-split-major item
    // order these are two row tiles over the SAME columns and K range - and walk one W stream in lock step: each CTA fetches
    // half of every K block's W planes as a cluster MULTICAST (both receive all eight), and a ring slot is released only
    // when both CTAs' MMAs have consumed it (commit multicast onto both empty barriers).  One L2 read of W feeds two SMs:
    // with A resident a vocabulary tile then costs 64 KB of L2 reads instead of 128 KB (the protocol of ffn_tc.cuh).
    if (a.pair) cluster_sync_all();   // every CTA's mbarriers exist before a peer multicasts into / arrives on them
    tc_fence_after_sync();
    if (threadIdx.x == 0) TL(1);
    pdl_wait();                 // PDL (common.cuh): nothing above touches global memory
    pdl_launch_dependents();
    const uint32_t crank = a.pair ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) TL(2);
    const uint32_t tmem_base = *s_tmem;

    const int m_tiles = a.Mp / kBM;
    const int n_tiles = a.N / kBN;
    const int items = m_tiles * a.nsplit * (a.ksplit > 1 ? a.ksplit : 1);
    const int kblocks = a.K / kBK;
    const bool a_resident = kblocks == kGemmStages && a.ksplit <= 1;

    if (warp == 0) {
        // ===== producer =====
        int kit = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const GemmItem w = gemm_item(a, item, n_tiles, kblocks);
            const int mt = w.mt;
            for (int nt = w.n0; nt < w.n1; ++nt) {
                // K == 4 blocks (width 256) and no K split: k-block kb always lands in ring slot kb, so the item's A tile
                // (128 rows x 256, 64 KB) is loaded once, with the first column tile, and stays in the slots' A halves
                // for every further column tile - 128 KB instead of 192 KB of L2 reads per tile, which is what the
                // vocabulary-sized GEMMs (AI 87 FLOP/B with A re-read) are bound by
                const bool load_a = !a_resident || nt == w.n0;
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++kit) {
                    const int s = kit % kGemmStages;
                    const uint32_t ph = (kit / kGemmStages) & 1;
                    if (lane == 0) {
                        mbar_wait(bar_empty + s, ph ^ 1);
                        mbar_arrive_expect_tx(bar_full + s, load_a ? kGemmStage : kGemmStageB);
                    }
                    __syncwarp();
                    uint8_t* st = smem + s * kGemmStage;
                    if (lane < 8) {
                        if (load_a)
                        bulk_g2s(st + lane * (kBM * 16), a.a + (int64_t)(kb * 8 + lane) * a.a_ps + (int64_t)mt * (kBM * 16),
                                 kBM * 16, bar_full + s);
                    } else if (lane < 16) {
                        const int j = lane - 8;
                        if (!a.pair)
                            bulk_g2s(st + kGemmStageA + j * (kBN * 16),
                                     a.w + (int64_t)(kb * 8 + j) * a.w_ps + (int64_t)nt * (kBN * 16), kBN * 16, bar_full + s);
                        else if ((uint32_t)(j & (a.pair - 1)) == crank)   // this CTA's share of the planes, to all CTAs
                            bulk_g2s_multicast(st + kGemmStageA + j * (kBN * 16),
                                               a.w + (int64_t)(kb * 8 + j) * a.w_ps + (int64_t)nt * (kBN * 16), kBN * 16,
                                               bar_full + s, (uint16_t)((1u << a.pair) - 1));
                    }
                    if (kit == 0) TL0(3);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected thread runs the whole issue loop (conv_tc.cuh: descriptors stay in uniform
        //       registers, no per-tile re-convergence) =====
        constexpr uint32_t idesc = make_idesc_bf16(kBM, kBN);
        if (elect_one()) {
        int kit = 0, it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const GemmItem w = gemm_item(a, item, n_tiles, kblocks);
            for (int nt = w.n0; nt < w.n1; ++nt, ++it) {
                const int acc = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(bar_acce + acc, aph ^ 1);
                const uint32_t d = tmem_base + acc * kBN;
                for (int kb = w.kb0; kb < w.kb1; ++kb, ++kit) {
                    const int s = kit % kGemmStages;
                    const uint32_t ph = (kit / kGemmStages) & 1;
                    mbar_wait(bar_full + s, ph);
                    tc_fence_after_sync();
                    if (kit == 0) TL(4);
                    const uint32_t a_addr = smem_u32(smem + s * kGemmStage);
                    const uint64_t a_base = make_smem_desc(a_addr, kBM * 16, 128);
                    const uint64_t b_base = make_smem_desc(a_addr + kGemmStageA, kBN * 16, 128);
                    const uint32_t acc_flag = kb != w.kb0;
#pragma unroll
                    for (int ks = 0; ks < kBK / 16; ++ks)
                        umma_bf16(d, desc_add(a_base, (2 * ks) * (kBM * 16)), desc_add(b_base, (2 * ks) * (kBN * 16)), idesc,
                                        ks != 0 ? 1u : acc_flag);
                    if (a.pair) umma_commit_multicast(bar_empty + s, (uint16_t)((1u << a.pair) - 1));
                    else umma_commit(bar_empty + s);
                }
                umma_commit(bar_accf + acc);
                if (it == 0) TL(5);
            }
        }
        }
        __syncwarp();
    } else {
        // ===== epilogue groups =====
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * kBN;
        int it = 0, n_mine = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const GemmItem w = gemm_item(a, item, n_tiles, kblocks);
            const int mt = w.mt, sp = w.sp, n0 = w.n0, n1 = w.n1;
            const int row = mt * kBM + q * 32 + lane;
            const bool rvalid = row < a.M;
            // GE_LSE: running maximum and sum of exp over this item's columns; GE_LSE / GE_DLOGITS: the row's target class
            float lse_m = -INFINITY, lse_s = 0.f, tgt_v = 0.f;
            bool tgt_hit = false;
            int tgt = -1;
            float row_lse = 0.f, gscale = 0.f;
            if constexpr (EPI == GE_LSE || EPI == GE_DLOGITS) {
                if (rvalid) tgt = (int)__ldg(a.target + row);
            }
            if constexpr (EPI == GE_DLOGITS) {
                // row_lse: the exponent's per-row constant log2(scale) - lse * log2 e (-inf for padding rows and for a zero
                // scale: every gradient of the row is then exactly zero)
                gscale = a.scale * (a.dev_scale ? __ldg(a.dev_scale) : 1.0f);
                row_lse = (rvalid && gscale > 0.f) ? log2f(gscale) - __ldg(a.lse + row) * 1.4426950408889634f : -INFINITY;
                if (!rvalid) gscale = 0.f;
            }
            // four independent running maxima (columns k % 4): one serial compare/select chain over all
            // 256 columns of a tile is a ~2000-cycle dependency chain per thread
            float best4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            int best4_i[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
            float rs = 1.f;
            if ((EPI == GE_ARGMAX || EPI == GE_LOGITS) && a.row_scale && rvalid) rs = __ldg(a.row_scale + row);
            // GE_ARGMAX with AR logits: the NEXT 32-column chunk of this warp's 32 AR rows is kept in flight in registers
            // (32 coalesced 128-byte row reads per lane) while the current chunk is mixed and compared.  Loaded chunk by
            // chunk on demand, each chunk was four dependent batches of eight loads at HBM latency: the guided mix ran at
            // 0.20 (512 sequences) - 0.55 (128) of the HBM roofline, latency-bound in this epilogue.
            // (two chunks ahead, in two register sets: the chunk loop is unrolled by two for this epilogue so that the
            // sets are addressed statically)
            constexpr int kSub = EPI == GE_ARGMAX ? 2 : 1;
            float nx[kSub][EPI == GE_ARGMAX ? 32 : 1];
            auto ar_load = [&](float (&dst)[EPI == GE_ARGMAX ? 32 : 1], int nt_, int c0_) {
                if constexpr (EPI == GE_ARGMAX) {
                    const int row0 = mt * kBM + q * 32;
                    const int ncol = nt_ * kBN + c0_ + lane;
                    const bool cok = ncol < a.n_valid;
                    const float* p = a.ar + (int64_t)row0 * a.ar_ld + ncol;   // one pointer, stepped by the row pitch
                    if (row0 + 32 <= a.M) {          // warp-uniform: all 32 rows exist (the usual case)
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr, p += a.ar_ld) dst[rr] = cok ? __ldg(p) : 0.f;
                    } else {
#pragma unroll
                        for (int rr = 0; rr < 32; ++rr, p += a.ar_ld) dst[rr] = (cok && row0 + rr < a.M) ? __ldg(p) : 0.f;
                    }
                }
            };
            if (EPI == GE_ARGMAX && a.ar) {
                int nt_first = n0 + ((grp - (it & 1)) & 1);     // this group's first tile of the item
                if (nt_first < n1) {
                    ar_load(nx[0], nt_first, 0);
                    ar_load(nx[kSub - 1], nt_first, 32);
                }
            }
            for (int nt = n0; nt < n1; ++nt, ++it) {
                if ((it & 1) != grp) continue;
                const uint32_t aph = (n_mine++) & 1;
                // The tile's 256 bias values, one 128-byte line per 32-column chunk (lane l <-> column 32 c + l), are fetched
                // BEFORE the accumulator wait: loaded chunk by chunk they put one L2 / HBM latency into every chunk's
                // dependent chain - eight per tile, which was most of a single-tile GEMM's 16-18 us.  The chunk loop is
                // not unrolled, so the eight registers are consumed by rotation (static indices).
                float bl[8];
                {
                    const bool with_bias = a.bias && (EPI != GE_LOGITS || w.ks == 0);
#pragma unroll
                    for (int c = 0; c < 8; ++c) bl[c] = with_bias ? __ldg(a.bias + nt * kBN + c * 32 + lane) : 0.f;
                }
                mbar_wait(bar_accf + grp, aph);
                tc_fence_after_sync();
                if (warp == 2 && n_mine == 1) TL0(6);
                float ln_sum = 0.f, ln_sq = 0.f;
#pragma unroll 1
                for (int c00 = 0; c00 < kBN; c00 += 32 * kSub) {
#pragma unroll
                for (int hsub = 0; hsub < kSub; ++hsub) {
                    const int c0 = c00 + 32 * hsub;
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    // lane l holds the bias of column c0 + l; columns get it by shuffle
                    const float bias_l = bl[0];
#pragma unroll
                    for (int c = 0; c < 7; ++c) bl[c] = bl[c + 1];
                    tmem_ld_wait();
                    if (warp == 2 && n_mine == 1 && c0 == 0) TL0(11);
                    if (warp == 2 && n_mine == 1 && c0 == 32) TL0(13);
                    if (warp == 2 && n_mine == 1 && c0 == 64) TL0(3);
                    if (EPI != GE_RES_LN && c0 + 32 == kBN) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_acce + grp);
                    }
                    const int nb = nt * kBN + c0;
                    if (warp == 2 && n_mine == 1 && c0 == 0) TL0(12);
                    if (warp == 2 && n_mine == 1 && c0 == 32) TL0(14);
                    if constexpr (EPI == GE_BF16) {
                        float vv[32];   // shuffles in straight-line code, stores under ONE row-validity branch (see GE_LOGITS)
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            vv[k] = __uint_as_float(r[k]) + __shfl_sync(0xffffffffu, bias_l, k);
                            if (a.relu) vv[k] = fmaxf(vv[k], 0.f);
                        }
                        if (rvalid) {
#pragma unroll
                            for (int pj = 0; pj < 4; ++pj) {
                                const float* v = vv + pj * 8;
                                uint4 o = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                     pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                                *reinterpret_cast<uint4*>(a.out_bf16 + (int64_t)(nb / 8 + pj) * a.ob_ps + (int64_t)row * 16) = o;
                            }
                        }
                    } else if constexpr (EPI == GE_RES_F32) {
                        float vv[32];   // shuffles in straight-line code; residual loads batched under one branch
#pragma unroll
                        for (int k = 0; k < 32; ++k) vv[k] = __uint_as_float(r[k]) + __shfl_sync(0xffffffffu, bias_l, k);
                        if (rvalid) {
                            float4 rv[8];
#pragma unroll
                            for (int pj = 0; pj < 8; ++pj)
                                rv[pj] = *reinterpret_cast<const float4*>(a.res + (int64_t)(nb / 4 + pj) * a.res_ps + (int64_t)row * 16);
#pragma unroll
                            for (int pj = 0; pj < 8; ++pj)
                                *reinterpret_cast<float4*>(a.out_f32 + (int64_t)(nb / 4 + pj) * a.of_ps + (int64_t)row * 16) =
                                    make_float4(vv[pj * 4 + 0] + rv[pj].x, vv[pj * 4 + 1] + rv[pj].y, vv[pj * 4 + 2] + rv[pj].z,
                                                vv[pj * 4 + 3] + rv[pj].w);
                        }
                    } else if constexpr (EPI == GE_RES_LN) {
                        // pass 1 of 2: v = acc + bias + residual; row statistics; park v in TMEM
                        // (shuffles in straight-line code, the eight residual loads issued together)
                        uint32_t vb[32];
                        float4 rv[8];
#pragma unroll
                        for (int pj = 0; pj < 8; ++pj) {
                            rv[pj] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (rvalid) rv[pj] = *reinterpret_cast<const float4*>(a.res + (int64_t)(nb / 4 + pj) * a.res_ps + (int64_t)row * 16);
                        }
#pragma unroll
                        for (int pj = 0; pj < 8; ++pj) {
                            const float v0 = __uint_as_float(r[pj * 4 + 0]) + __shfl_sync(0xffffffffu, bias_l, pj * 4 + 0) + rv[pj].x;
                            const float v1 = __uint_as_float(r[pj * 4 + 1]) + __shfl_sync(0xffffffffu, bias_l, pj * 4 + 1) + rv[pj].y;
                            const float v2 = __uint_as_float(r[pj * 4 + 2]) + __shfl_sync(0xffffffffu, bias_l, pj * 4 + 2) + rv[pj].z;
                            const float v3 = __uint_as_float(r[pj * 4 + 3]) + __shfl_sync(0xffffffffu, bias_l, pj * 4 + 3) + rv[pj].w;
                            ln_sum += (v0 + v1) + (v2 + v3);
                            ln_sq = fmaf(v0, v0, fmaf(v1, v1, fmaf(v2, v2, fmaf(v3, v3, ln_sq))));
                            vb[pj * 4 + 0] = __float_as_uint(v0);
                            vb[pj * 4 + 1] = __float_as_uint(v1);
                            vb[pj * 4 + 2] = __float_as_uint(v2);
                            vb[pj * 4 + 3] = __float_as_uint(v3);
                        }
                        tmem_st32(taddr + c0, vb);
                    } else if constexpr (EPI == GE_LOGITS) {
                        // all 32 shuffles first, in straight-line code: interleaved with the row / column validity branches
                        // every group of four sat inside its own convergence region (BSSY / BRA.DIV / BSYNC around each
                        // SHFL quartet) and one chunk of one warp took ~1,900 cycles
                        float v[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            v[k] = fmaf(__uint_as_float(r[k]), rs, __shfl_sync(0xffffffffu, bias_l, k));
                            if (a.relu) v[k] = fmaxf(v[k], 0.f);
                        }
                        if (warp == 2 && n_mine == 1 && c0 == 32 && __float_as_uint(v[31]) != 0x7fc12345u) TL0(15);
                        float* orow = a.logits + (int64_t)w.ks * a.split_stride + (int64_t)row * a.logits_ld + nb;
                        const float* arow = (a.logits_add && w.ks == 0) ? a.logits_add + (int64_t)row * a.logits_ld + nb : nullptr;
                        const bool vec = (a.logits_ld & 3) == 0 && nb + 32 <= a.n_valid;   // 16-byte aligned, whole chunk valid
                        if (rvalid) {
                            if (vec) {
                                if (arow) {
                                    float4 ad[8];
#pragma unroll
                                    for (int k4 = 0; k4 < 8; ++k4) ad[k4] = __ldg(reinterpret_cast<const float4*>(arow) + k4);
#pragma unroll
                                    for (int k4 = 0; k4 < 8; ++k4) {
                                        v[4 * k4] += ad[k4].x; v[4 * k4 + 1] += ad[k4].y; v[4 * k4 + 2] += ad[k4].z; v[4 * k4 + 3] += ad[k4].w;
                                    }
                                }
#pragma unroll
                                for (int k4 = 0; k4 < 8; ++k4)
                                    reinterpret_cast<float4*>(orow)[k4] = make_float4(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3]);
                            } else {
#pragma unroll
                                for (int k = 0; k < 32; ++k)
                                    if (nb + k < a.n_valid) orow[k] = arow ? v[k] + __ldg(arow + k) : v[k];
                            }
                        }
                        if (warp == 2 && n_mine == 1 && c0 == 32) TL0(10);
                    } else if constexpr (EPI == GE_LSE) {
                        // logits of this row x these 32 columns in the base-2 domain (v2 = (acc + bias) * log2 e): fold into
                        // the running (max, sum 2^(v2 - max)) pair.  Padding columns (last vocabulary tile only) and the
                        // target's logit are handled on rarely-taken branches, not per element.
                        constexpr float kLog2e = 1.4426950408889634f;
                        const float bias2_l = bias_l * kLog2e;
                        float v[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k) v[k] = fmaf(__uint_as_float(r[k]), kLog2e, __shfl_sync(0xffffffffu, bias2_l, k));
                        if (nb + 32 > a.n_valid) {
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (nb + k >= a.n_valid) v[k] = -INFINITY;
                        }
                        if ((unsigned)(tgt - nb) < 32u) {
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (nb + k == tgt) tgt_v = v[k];
                            tgt_hit = true;
                        }
                        float c4[4] = {v[0], v[1], v[2], v[3]};
#pragma unroll
                        for (int k = 4; k < 32; ++k) c4[k & 3] = fmaxf(c4[k & 3], v[k]);
                        const float cmax = fmaxf(fmaxf(c4[0], c4[1]), fmaxf(c4[2], c4[3]));
                        if (cmax > -INFINITY) {
                            if (cmax > lse_m) {
                                lse_s *= ex2_approx(lse_m - cmax);   // 2^(-inf) = 0 on the first chunk
                                lse_m = cmax;
                            }
                            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                            for (int k = 0; k < 32; ++k) s4[k & 3] += ex2_approx(v[k] - lse_m);
                            lse_s += (s4[0] + s4[1]) + (s4[2] + s4[3]);
                        }
                    } else if constexpr (EPI == GE_DLOGITS) {
                        // d loss / d logits = (softmax - onehot) * scale, zero for padding rows and columns (both are
                        // reduction indices of the gradient GEMMs that read these planes).  softmax * scale =
                        // 2^((acc + bias) log2 e - lse log2 e + log2 scale): one FMA, one add and one MUFU per element; the
                        // onehot and the padding columns are fixed up on rarely-taken branches.
                        constexpr float kLog2e = 1.4426950408889634f;
                        const float bias2_l = bias_l * kLog2e;
                        const bool full = nb + 32 <= a.n_valid;
                        float* art = reinterpret_cast<float*>(smem + kGemmSmemBase) + (warp - 2) * kArTileFloats;
                        if (a.out_bf16_t) __syncwarp();   // the previous chunk's reads of the tile are done
                        float gv[32];
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            gv[k] = ex2_approx(fmaf(__uint_as_float(r[k]), kLog2e, __shfl_sync(0xffffffffu, bias2_l, k)) + row_lse);
                        if ((unsigned)(tgt - nb) < 32u) {
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (nb + k == tgt) gv[k] -= gscale;
                        }
                        if (!full) {
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (nb + k >= a.n_valid) gv[k] = 0.f;
                        }
#pragma unroll
                        for (int pj = 0; pj < 4; ++pj) {
                            const float* gk = gv + pj * 8;
                            const uint4 o = make_uint4(pack_bf16x2(gk[0], gk[1]), pack_bf16x2(gk[2], gk[3]),
                                                       pack_bf16x2(gk[4], gk[5]), pack_bf16x2(gk[6], gk[7]));
                            *reinterpret_cast<uint4*>(a.out_bf16 + (int64_t)(nb / 8 + pj) * a.ob_ps + (int64_t)row * 16) = o;
                            if (a.out_bf16_t) {
#pragma unroll
                                for (int k = 0; k < 8; ++k) art[lane * 33 + pj * 8 + k] = gk[k];
                            }
                        }
                        if (a.out_bf16_t) {
                            // the same 32 x 32 block with lane = column: eight consecutive rows (tokens) of one column are
                            // one 16-byte plane row; the warp's 32 columns make each store 512 contiguous bytes
                            __syncwarp();
                            const int64_t rp0 = (int64_t)(mt * kBM + q * 32) / 8;
#pragma unroll
                            for (int g8 = 0; g8 < 4; ++g8) {
                                float tv[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) tv[j] = art[(g8 * 8 + j) * 33 + lane];
                                *reinterpret_cast<uint4*>(a.out_bf16_t + ((rp0 + g8) * a.obt_rows + nb + lane) * 16) =
                                    make_uint4(pack_bf16x2(tv[0], tv[1]), pack_bf16x2(tv[2], tv[3]), pack_bf16x2(tv[4], tv[5]),
                                               pack_bf16x2(tv[6], tv[7]));
                            }
                        }
                    } else {
                        const bool full = nb + 32 <= a.n_valid;   // tile-uniform: only the last tile of a padded vocabulary is partial
                        // AR logits of this warp's 32 rows x these 32 columns.  Read row-wise - thread = row, 32 consecutive
                        // floats each - every load instruction touches 32 different 128-byte lines (1,024 LSU wavefronts per
                        // chunk; the guided mix ran at 0.15 of the HBM roofline on exactly this).  Instead the warp reads one
                        // ROW per instruction (lane = column: one coalesced 128-byte request) and turns the tile through
                        // shared memory (row stride 33 words: conflict-free both ways).
                        float* art = reinterpret_cast<float*>(smem + kGemmSmemBase) + (warp - 2) * kArTileFloats;
                        if (a.ar) {
                            __syncwarp();   // the previous chunk's reads of the tile are done
                            // lane = column here: the column's bias term (and -inf for the padding columns of the last
                            // vocabulary tile) goes in with the AR logit, so the row-wise pass needs no shuffle and no
                            // validity select per element
                            const float cb = (full || nb + lane < a.n_valid) ? a.alpha * a.inv_temp * bias_l : -INFINITY;
                            const float car = (1.0f - a.alpha) * a.inv_temp;
#pragma unroll
                            for (int rr = 0; rr < 32; ++rr) art[rr * 33 + lane] = fmaf(car, nx[hsub][rr], cb);
                            __syncwarp();
                            // this register set is free: the chunk after the next one (of this tile, or of this group's next tile)
                            if (c0 + 64 < kBN) ar_load(nx[hsub], nt, c0 + 64);
                            else if (nt + 2 < n1) ar_load(nx[hsub], nt + 2, c0 + 64 - kBN);
                        }
                        // src/shakespeare.py:449-466: both logit sets divided by the temperature, then mixed
                        // (1-alpha)*ar + alpha*diff.  The constants are folded per thread (row) and per chunk (bias):
                        //   v = c_dot * dot + (c_ar * ar + bias_s),  c_ar = (1-alpha)/T,  c_dot = alpha*rs/T,  bias_s = alpha*bias/T
                        // (the bracket is formed when the AR chunk is turned through shared memory) - two FMAs per element
                        // instead of six multiplies and adds; the value differs from the reference's op order in the last
                        // bit, which the token parity bar (exact where the top-2 margin allows) absorbs
                        if (a.ar) {   // (kernel-uniform: two copies of the loop, no select per element)
                            const float c_dot = a.alpha * a.inv_temp * rs;
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                const float v = fmaf(c_dot, __uint_as_float(r[k]), art[lane * 33 + k]);
                                if (v > best4[k & 3]) {   // strict: first (lowest) index wins ties, as torch.argmax
                                    best4[k & 3] = v;
                                    best4_i[k & 3] = nb + k;
                                }
                            }
                        } else {
                            const float bias_s = (full || nb + lane < a.n_valid) ? bias_l : -INFINITY;
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                const float v = fmaf(rs, __uint_as_float(r[k]), __shfl_sync(0xffffffffu, bias_s, k));
                                if (v > best4[k & 3]) {
                                    best4[k & 3] = v;
                                    best4_i[k & 3] = nb + k;
                                }
                            }
                        }
                    }
                }
                }
                if constexpr (EPI == GE_RES_LN) {
                    // pass 2 of 2: normalise (nn.LayerNorm: biased variance, eps inside the sqrt), write both formats
                    tmem_st_wait();
                    const float mean = ln_sum * (1.0f / kBN);
                    const float rstd = rsqrtf(fmaxf(ln_sq * (1.0f / kBN) - mean * mean, 0.f) + a.ln_eps);
#pragma unroll 1
                    for (int c0 = 0; c0 < kBN; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(taddr + c0, r);
                        tmem_ld_wait();
                        if (c0 + 32 == kBN) {
                            tc_fence_before_sync();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_acce + grp);
                        }
                        if (rvalid) {
#pragma unroll
                            for (int pj = 0; pj < 4; ++pj) {
                                float y[8];
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    const int n = c0 + pj * 8 + k;
                                    y[k] = (__uint_as_float(r[pj * 8 + k]) - mean) * rstd * __ldg(a.gamma + n) + __ldg(a.beta + n);
                                }
                                *reinterpret_cast<float4*>(a.out_f32 + (int64_t)(c0 / 4 + 2 * pj) * a.of_ps + (int64_t)row * 16) = make_float4(y[0], y[1], y[2], y[3]);
                                *reinterpret_cast<float4*>(a.out_f32 + (int64_t)(c0 / 4 + 2 * pj + 1) * a.of_ps + (int64_t)row * 16) = make_float4(y[4], y[5], y[6], y[7]);
                                *reinterpret_cast<uint4*>(a.out_bf16 + (int64_t)(c0 / 8 + pj) * a.ob_ps + (int64_t)row * 16) =
                                    make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
                            }
                        }
                    }
                }
            }
            if constexpr (EPI == GE_LSE) {
                const int64_t slot = (int64_t)(sp * 2 + grp) * a.Mp + row;
                a.part_val[slot] = lse_m * 0.6931471805599453f;   // back to natural-log units (the sum is unit-free)
                a.part_sum[slot] = lse_s;
                if (tgt_hit) a.tgt_logit[row] = tgt_v * 0.6931471805599453f;
            }
            if constexpr (EPI == GE_ARGMAX) {
                float best = best4[0];
                int bi = best4_i[0];
#pragma unroll
                for (int j = 1; j < 4; ++j) {
                    if (best4[j] > best || (best4[j] == best && best4_i[j] < bi)) {
                        best = best4[j];
                        bi = best4_i[j];
                    }
                }
                const int64_t slot = (int64_t)(sp * 2 + grp) * a.Mp + row;
                a.part_val[slot] = best;
                a.part_idx[slot] = bi == 0x7fffffff ? INT64_MAX : (int64_t)bi;
            }
        }
    }
    if (warp == 2) TL0(7);
    __syncwarp();
    tc_fence_before_sync();
    __syncthreads();
    if (a.pair) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it or arrive on its barriers
    tc_fence_after_sync();
    if (threadIdx.x == 0) TL(8);
    if (warp == 2) tmem_dealloc<512>(tmem_base);
    if (warp == 2) TL0(9);
}

template <int EPI>
static int launch_gemm(const GemmArgs& a_in, cudaStream_t st, const char* name) {
    GemmArgs a = a_in;
    auto kern = gemm_tc_kernel<EPI>;
    TDM_SET_MAX_DYN_SMEM(kern, kGemmSmem);
    TDM_CHECK_ARG(EPI != GE_RES_LN || a.N == kBN, "%s: fused LayerNorm needs N == 256", name);
    TDM_CHECK_ARG(a.Mp % kBM == 0 && a.N % kBN == 0 && a.K % kBK == 0 && a.K > 0 && a.nsplit > 0,
                  "%s: bad GEMM shape M=%d Mp=%d N=%d K=%d", name, a.M, a.Mp, a.N, a.K);
    TDM_CHECK_ARG(a.ksplit <= 1 || EPI == GE_LOGITS, "%s: the K split exists for the fp32 row-major epilogue only", name);
    const int m_tiles = a.Mp / kBM;
    const int items = m_tiles * a.nsplit * (a.ksplit > 1 ? a.ksplit : 1);
    int grid = items < num_sms() ? items : num_sms();
    // pairs: consecutive items must be two row tiles of one column range (even row-tile count, no K split), and both CTAs
    // of a cluster must run the same number of items (even item count and grid)
    static const bool pair_ok = [] {
        const char* e = std::getenv("TDM_NO_PAIR");
        return !(e && e[0] == '1');
    }();
    static const int csize_env = [] {
        const char* e = std::getenv("TDM_GEMM_CLUSTER");
        return e ? std::atoi(e) : 2;
    }();
    int cs = (csize_env == 4 && m_tiles % 4 == 0) ? 4 : 2;
    a.pair = (pair_ok && m_tiles % cs == 0 && a.ksplit <= 1 && grid >= cs) ? cs : 0;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (a.pair) {
        grid -= grid % cs;
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cs;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = kGemmSmem;
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = na;
    TDM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    TDM_CHECK_LAUNCH(name);
    return TDM_OK;
}

}  // namespace tdm
