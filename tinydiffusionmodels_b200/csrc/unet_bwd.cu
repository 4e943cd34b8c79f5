// unet_bwd.cu — backward pass and optimizer of the MNIST DDPM training step (src/mnist.py:153-159):
//   loss = mse(model(q_sample(x0, t, noise), t), noise); loss.backward(); AdamW.step()
//
// Gradients flow through the same plane layout as the activations (bf16, fp32 accumulation):
//   * data gradients   : conv3x3_tc_kernel (conv_tc.cuh) with transposed, tap-flipped weights; conv2's data
//                        gradient is multiplied by conv1's ReLU mask and reduced into the time-embedding / conv1-bias
//                        gradients in the same epilogue (EPI_PLAIN_MASK)
//   * weight gradients : wgrad_dup_kernel (3x3) / wgrad_tc_kernel (1x1 skips) below — tcgen05.mma with BOTH operands
//                        MN-major straight from the plane tiles (K = positions), all taps accumulated in TMEM across
//                        the CTA's tiles, flushed once with 16-byte vector reductions into a tap-major scratch that
//                        wgrad_unpermute_kernel turns into the flat gradient
//   * a block's output gradient, its ReLU mask and the skip/conv2-bias gradients: one pass each —
//     loss_grad_kernel (rb4, with the MSE gradient and the out conv), mask_reduce_kernel<MR_UPSAMPLE_T> (rb3, with
//     the transpose of the upsample), <MR_POOL_T> (rb1, with the concat slice and the transpose of the pool),
//     <MR_PLANES> (rb2)
//   * rb1.conv1 / rb1.skip (Cin = 1): rb1_wgrad_kernel (SIMT)
// All parameter gradients land in ONE flat fp32 buffer in the reference's state_dict order — the buffer the fused
// AdamW operates on; data parallel, that buffer is a slot of a peer-mapped allocation and adamw_kernel<true> sums
// the ranks' gradients while it applies the update (peer.cu).
#include "common.cuh"
#include "conv_tc.cuh"
#include "tc05.cuh"
#include "unet_layout.cuh"
#include "unet_ws.cuh"

namespace tdm {

// ---------------------------------------------------------------------------------------------
// weight gradient on tensor cores
//   dW[cg][cx][tap] += sum_pos G[pos][cg] * X[pos + off(tap)][cx]
// A = G^T (M = cg, K = pos), B = X^T (N = cx, K = pos): in the plane layout [c/8][pos][8] the
// channel index is the contiguous one, i.e. both operands are MN-major, un-swizzled:
//   8 channels contiguous (16 B), 8 positions at 16 B stride (one 128 B core matrix),
//   channel groups at SBO = plane stride, position groups at LBO = 128 B.
// M = 64: TMEM row m lives in lane (m%16) + 32*(m/16); a second accumulator set is interleaved at
// lane offset 16, so nine taps need only 5*CX columns.
// ---------------------------------------------------------------------------------------------
// 16 consecutive fp32 accumulated with four 16-byte vector reductions (RED.E.ADD.F32x4 on sm_90+): the
// per-CTA flush of a weight gradient was the cost of the small-batch wgrad kernels when it went out as
// 4-byte atomics 36 bytes apart (OIHW order: the 9 taps of one (co, ci) are adjacent, consecutive ci are not).
__device__ __forceinline__ void red_add_f32x16(float* dst, const uint32_t (&r)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
        atomicAdd(reinterpret_cast<float4*>(dst) + i,
                  make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                              __uint_as_float(r[4 * i + 3])));
}

struct WgradArgs {
    const uint8_t* g;   // gradient planes (row -HALO of plane 0), CG channels
    int64_t g_ps;
    const uint8_t* x;   // activation planes, CX channels
    int64_t x_ps;
    float* dw;          // fp32, accumulated with 16-byte vector atomics: [CG][CX] (1x1) or TAP-MAJOR [9][CG][CX] (3x3 scratch)
    int nt;
    // wgrad_dup_kernel<.., SKIPW = true> (CG = 32): a second gradient (the block's unmasked output gradient) rides in
    // the spare quarter of the M = 128 tile and yields the 1x1 skip's weight gradient dws[CG][CX] from the same X tile
    const uint8_t* g2;
    float* dws;
};

template <int W, int CG, int CX, int TAPS>
struct WgradCfg {
    using G = Geo<W>;
    static constexpr int GPL = CG / 8, XPL = CX / 8;
    static constexpr int G_BYTES = GPL * kTile * 16;
    static constexpr int X_BYTES = XPL * G::RT * 16;
    static constexpr int STAGE_BYTES = G_BYTES + X_BYTES;
    static_assert(CG == 64 || X_BYTES >= 8 * kTile * 16 - G_BYTES, "M=64 pad rows must stay inside the stage");
    static constexpr int AVAIL = 227 * 1024 - 1024;
    static constexpr int NSTAGE = (AVAIL / STAGE_BYTES) > 4 ? 4 : (AVAIL / STAGE_BYTES);
    static_assert(NSTAGE >= 2, "stages");
    static constexpr int NCOLS = (TAPS == 9 ? 5 : 1) * CX;
    static constexpr int TMEM_COLS = NCOLS <= 32 ? 32 : NCOLS <= 64 ? 64 : NCOLS <= 128 ? 128 : NCOLS <= 256 ? 256 : 512;
    static_assert(NCOLS <= 512, "TMEM");
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 512 > kSoloSmem ? NSTAGE * STAGE_BYTES + 512 : kSoloSmem;   // one CTA per SM (conv_tc.cuh)
    static constexpr int THREADS = 192;
};

template <int W, int CG, int CX, int TAPS>
__global__ void __launch_bounds__(192, 1) wgrad_tc_kernel(const WgradArgs a) {
    using C = WgradCfg<W, CG, CX, TAPS>;
    using G = Geo<W>;
    static_assert(CG == 32 || CG == 64, "CG");
    static_assert(CX % 16 == 0 && CX <= 96, "CX");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_in = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE_BYTES);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + C::NSTAGE;
    uint64_t* bar_done = bar_empty + C::NSTAGE;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NSTAGE; ++i) {
            mbar_init(bar_full + i, 1);
            mbar_init(bar_empty + i, 1);
        }
        mbar_init(bar_done, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<C::TMEM_COLS>(s_tmem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_wait();                 // PDL (common.cuh): nothing above touches global memory
    pdl_launch_dependents();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 0) {
        int it = 0;
        for (int tile = blockIdx.x; tile < a.nt; tile += gridDim.x, ++it) {
            const int s = it % C::NSTAGE;
            const uint32_t ph = (it / C::NSTAGE) & 1;
            if (lane == 0) {
                mbar_wait(bar_empty + s, ph ^ 1);
                mbar_arrive_expect_tx(bar_full + s, C::STAGE_BYTES);
            }
            __syncwarp();
            uint8_t* st = s_in + s * C::STAGE_BYTES;
            if (lane < C::GPL) {
                bulk_g2s(st + lane * (kTile * 16),
                         a.g + lane * a.g_ps + ((int64_t)tile * kTile + G::GUARD) * 16, kTile * 16,
                         bar_full + s);
            } else if (lane >= 16 && lane < 16 + C::XPL) {
                const int j = lane - 16;
                bulk_g2s(st + C::G_BYTES + j * (G::RT * 16), a.x + j * a.x_ps + ((int64_t)tile * kTile - G::HALO + G::GUARD) * 16,
                         G::RT * 16, bar_full + s);
            }
        }
    } else if (warp == 1) {
        // whole warp, one elected lane issues (tc05.cuh: warp-convergent issue)
        constexpr uint32_t idesc = make_idesc_bf16(64, CX, 1, 1);
        if (elect_one()) {   // one thread runs the whole issue loop (conv_tc.cuh: uniform-register descriptors, no per-tile re-convergence)
        int it = 0;
        for (int tile = blockIdx.x; tile < a.nt; tile += gridDim.x, ++it) {
            const int s = it % C::NSTAGE;
            const uint32_t ph = (it / C::NSTAGE) & 1;
            mbar_wait(bar_full + s, ph);
            tc_fence_after_sync();
            const uint32_t g_addr = smem_u32(s_in + s * C::STAGE_BYTES);
            const uint64_t g_base = make_smem_desc(g_addr, 128, kTile * 16);
            const uint64_t x_base = make_smem_desc(g_addr + C::G_BYTES, 128, G::RT * 16);
            const uint32_t acc_flag = it != 0;
#pragma unroll
            for (int tap = 0; tap < TAPS; ++tap) {
                const int off = (TAPS == 1) ? 0 : (tap / 3 - 1) * G::Wp + (tap % 3 - 1);
                const uint32_t d = tmem_base + (tap < 5 ? tap * CX : ((16u << 16) + (tap - 5) * CX));
#pragma unroll
                for (int ks = 0; ks < kTile / 16; ++ks)
                    umma_bf16(d, desc_add(g_base, ks * 256), desc_add(x_base, (G::HALO + off + ks * 16) * 16), idesc,
                                    ks != 0 ? 1u : acc_flag);
            }
            umma_commit(bar_empty + s);
        }
        if (it > 0) umma_commit(bar_done);
        }
        __syncwarp();
    } else if ((int)blockIdx.x < a.nt) {
        // ===== one final flush: TMEM -> atomics into the flat gradient =====
        mbar_wait(bar_done, 0);
        tc_fence_after_sync();
        const int q = warp & 3;
        const int m = q * 16 + (lane & 15);
        const int upper = lane >> 4;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr int NSET = (TAPS == 9) ? 5 : 1;
#pragma unroll
        for (int ts = 0; ts < NSET; ++ts) {
            const int tap = upper ? 5 + ts : ts;
            const bool live = m < CG && tap < TAPS;
#pragma unroll
            for (int c0 = 0; c0 < CX; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + ts * CX + c0, r);
                tmem_ld_wait();
                if (live) {
                    if constexpr (TAPS == 1) {
                        red_add_f32x16(a.dw + (int64_t)m * CX + c0, r);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            atomicAdd(a.dw + ((int64_t)(m * CX + c0 + i) * TAPS + tap), __uint_as_float(r[i]));
                    }
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------
// 3x3 weight gradient with row-shifted duplicates of the gradient tile.
// With A = G^T only CG (32 or 64) of an MMA's M rows carry data.  Staging copy d of the G tile shifted
// by d rows (copy d = rows [r0-d, r0-d+128)) makes M-row block d compute
//     sum_k G[k-d][co] X[k+off][ci]  =  the tap whose offset is off + d
// so with off = (ky-1)*Wp - 1 one MMA yields the kx = 0,1(,2) taps of a kernel row:
//   CG = 32: three copies -> M = 128 (96 live rows), 3 MMA groups per k-step instead of 9
//   CG = 64: two copies   -> M = 128 for (kx0, kx1) plus an M = 64 MMA for kx2 on copy 0
// Every copy's 128-row windows tile the position axis (offset by d), so each (pos, tap) product is
// counted exactly once; two extra rows past np are swept so copy 2 reaches the last positions.
// ---------------------------------------------------------------------------------------------
template <int W, int CG, int CX>
struct WgradDupCfg {
    using G = Geo<W>;
    static constexpr int NDUP = (CG == 32) ? 3 : 2;
    static constexpr int GPL = CG / 8, XPL = CX / 8;
    static constexpr int G_BYTES = 16 * kTile * 16;            // 16 M-groups of 8 channels (M = 128), 32 KB
    static constexpr int X_BYTES = XPL * G::RT * 16;
    static constexpr int STAGE_BYTES = G_BYTES + X_BYTES;
    static constexpr int AVAIL = 227 * 1024 - 1024;
    static constexpr int NSTAGE = (AVAIL / STAGE_BYTES) > 4 ? 4 : (AVAIL / STAGE_BYTES);
    static_assert(NSTAGE >= 2, "stages");
    // TMEM: per kernel row ky one [128 x CX] accumulator; CG = 64 adds one [64 x CX] (kx2) per ky,
    // two of which share columns through the lane-16 interleave
    static constexpr int NCOLS = 3 * CX + (CG == 64 ? 2 * CX : 0);
    static constexpr int TMEM_COLS = NCOLS <= 128 ? 128 : NCOLS <= 256 ? 256 : 512;
    static_assert(NCOLS <= 512, "TMEM");
    static constexpr int SMEM_BYTES = NSTAGE * STAGE_BYTES + 512 > kSoloSmem ? NSTAGE * STAGE_BYTES + 512 : kSoloSmem;   // one CTA per SM (conv_tc.cuh)
    static constexpr int THREADS = 192;
};

template <int W, int CG, int CX, bool SKIPW = false>
__global__ void __launch_bounds__(192, 1) wgrad_dup_kernel(const WgradArgs a) {
    using C = WgradDupCfg<W, CG, CX>;
    using G = Geo<W>;
    static_assert(CG == 32 || CG == 64, "CG");
    static_assert(!SKIPW || CG == 32, "the skip gradient uses the fourth quarter of the M = 128 tile (three copies at CG = 32)");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_in = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::NSTAGE * C::STAGE_BYTES);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + C::NSTAGE;
    uint64_t* bar_done = bar_empty + C::NSTAGE;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NSTAGE; ++i) {
            mbar_init(bar_full + i, 1);
            mbar_init(bar_empty + i, 1);
        }
        mbar_init(bar_done, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<C::TMEM_COLS>(s_tmem);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_wait();                 // PDL (common.cuh): nothing above touches global memory
    pdl_launch_dependents();
    const uint32_t tmem_base = *s_tmem;
    // 12 (CG=32) or 16 (CG=64) of the 16 M-groups hold data; SKIPW fills the last four with the second gradient,
    // staged like copy 1 (shifted one row) so that the ky = 1 MMA, whose X window starts one row early, pairs
    // g2[pos] with X[pos]: the centre tap, i.e. the 1x1 convolution
    constexpr int LIVE_PLANES = C::NDUP * C::GPL + (SKIPW ? C::GPL : 0);
    constexpr int LIVE_BYTES = LIVE_PLANES * kTile * 16 + C::X_BYTES;

    if (warp == 0) {
        int it = 0;
        for (int tile = blockIdx.x; tile < a.nt; tile += gridDim.x, ++it) {
            const int s = it % C::NSTAGE;
            const uint32_t ph = (it / C::NSTAGE) & 1;
            if (lane == 0) {
                mbar_wait(bar_empty + s, ph ^ 1);
                mbar_arrive_expect_tx(bar_full + s, LIVE_BYTES);
            }
            __syncwarp();
            uint8_t* st = s_in + s * C::STAGE_BYTES;
            if (lane < LIVE_PLANES) {
                const int d = lane / C::GPL, j = lane - d * C::GPL;   // copy d, channel plane j
                const uint8_t* gsrc = (SKIPW && d == C::NDUP) ? a.g2 : a.g;
                const int shift = (SKIPW && d == C::NDUP) ? 1 : d;
                bulk_g2s(st + lane * (kTile * 16), gsrc + j * a.g_ps + ((int64_t)tile * kTile - shift + G::GUARD) * 16,
                         kTile * 16, bar_full + s);
            } else if (lane >= 16 && lane < 16 + C::XPL) {
                const int j = lane - 16;
                bulk_g2s(st + C::G_BYTES + j * (G::RT * 16),
                         a.x + j * a.x_ps + ((int64_t)tile * kTile - G::HALO + G::GUARD) * 16, G::RT * 16, bar_full + s);
            }
        }
    } else if (warp == 1) {
        // whole warp, one elected lane issues (tc05.cuh: warp-convergent issue)
        constexpr uint32_t idesc128 = make_idesc_bf16(128, CX, 1, 1);
        constexpr uint32_t idesc64 = make_idesc_bf16(64, CX, 1, 1);
        if (elect_one()) {   // one thread runs the whole issue loop (conv_tc.cuh: uniform-register descriptors, no per-tile re-convergence)
        int it = 0;
        for (int tile = blockIdx.x; tile < a.nt; tile += gridDim.x, ++it) {
            const int s = it % C::NSTAGE;
            const uint32_t ph = (it / C::NSTAGE) & 1;
            mbar_wait(bar_full + s, ph);
            tc_fence_after_sync();
            const uint32_t g_addr = smem_u32(s_in + s * C::STAGE_BYTES);
            const uint64_t g_base = make_smem_desc(g_addr, 128, kTile * 16);
            const uint64_t x_base = make_smem_desc(g_addr + C::G_BYTES, 128, G::RT * 16);
            const uint32_t acc_flag = it != 0;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int off = (ky - 1) * G::Wp - 1;   // kx = 0; copy d adds d
#pragma unroll
                for (int ks = 0; ks < kTile / 16; ++ks) {
                    const uint64_t ad = desc_add(g_base, ks * 256);
                    umma_bf16(tmem_base + ky * CX, ad, desc_add(x_base, (G::HALO + off + ks * 16) * 16), idesc128,
                                    ks != 0 ? 1u : acc_flag);
                    if constexpr (CG == 64) {
                        // kx = 2 from copy 0 with the X window moved two rows on; ky = 0,1 share columns
                        // 3*CX.. through the lane-16 interleave of M = 64 accumulators, ky = 2 sits at 4*CX
                        const uint32_t d2 = tmem_base + (ky < 2 ? 3 * CX + ((uint32_t)(ky * 16) << 16) : 4 * CX);
                        umma_bf16(d2, ad, desc_add(x_base, (G::HALO + off + 2 + ks * 16) * 16), idesc64,
                                        ks != 0 ? 1u : acc_flag);
                    }
                }
            }
            umma_commit(bar_empty + s);
        }
        if (it > 0) umma_commit(bar_done);
        }
        __syncwarp();
    } else if ((int)blockIdx.x < a.nt) {
        mbar_wait(bar_done, 0);
        tc_fence_after_sync();
        const int q = warp & 3;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        // --- M = 128 accumulators: row m = d*CG + co lives in lane m ---
        const int m = q * 32 + lane;
        const int d = m / CG, co = m - d * CG;
        const bool live = d < C::NDUP;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
            for (int c0 = 0; c0 < CX; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + ky * CX + c0, r);
                tmem_ld_wait();
                if (live) red_add_f32x16(a.dw + ((int64_t)((ky * 3 + d) * CG + co) * CX + c0), r);
                if (SKIPW && d == C::NDUP && ky == 1) red_add_f32x16(a.dws + (int64_t)co * CX + c0, r);
            }
        }
        if constexpr (CG == 64) {
            // --- M = 64 accumulators (kx = 2): row m2 = q*16 + (lane & 15); lanes >= 16 hold ky = 1 ---
            const int m2 = q * 16 + (lane & 15);
            const int upper = lane >> 4;
#pragma unroll
            for (int blk = 0; blk < 2; ++blk) {   // blk 0: columns 3*CX (ky 0 / ky 1 interleaved), blk 1: 4*CX (ky 2)
                const int ky = blk == 0 ? upper : 2;
                const bool ok = blk == 0 || upper == 0;
#pragma unroll
                for (int c0 = 0; c0 < CX; c0 += 16) {
                    uint32_t r[16];
                    tmem_ld16(taddr + (3 + blk) * CX + c0, r);
                    tmem_ld_wait();
                    if (ok) red_add_f32x16(a.dw + ((int64_t)((ky * 3 + 2) * CG + m2) * CX + c0), r);
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int W, int CG, int CX, bool SKIPW = false>
static int launch_wgrad_dup(WgradArgs a, int64_t np, cudaStream_t st, const char* name) {
    using C = WgradDupCfg<W, CG, CX>;
    auto kern = wgrad_dup_kernel<W, CG, CX, SKIPW>;
    TDM_SET_MAX_DYN_SMEM(kern, C::SMEM_BYTES);
    a.nt = (int)((np + 2 + kTile - 1) / kTile);   // two rows past np so the shifted copies reach the last positions
    const int grid = a.nt < num_sms() ? a.nt : num_sms();
    launch_pdl(kern, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, a);
    TDM_CHECK_LAUNCH(name);
    return TDM_OK;
}

template <int W, int CG, int CX, int TAPS>
static int launch_wgrad(const WgradArgs& a, cudaStream_t st, const char* name) {
    using C = WgradCfg<W, CG, CX, TAPS>;
    auto kern = wgrad_tc_kernel<W, CG, CX, TAPS>;
    TDM_SET_MAX_DYN_SMEM(kern, C::SMEM_BYTES);
    const int grid = a.nt < num_sms() ? a.nt : num_sms();
    launch_pdl(kern, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, a);
    TDM_CHECK_LAUNCH(name);
    return TDM_OK;
}

// ---------------------------------------------------------------------------------------------
// elementwise pieces
// ---------------------------------------------------------------------------------------------
constexpr int kChunk = 512;   // positions per block in the reducing elementwise kernels (few same-address atomics)
constexpr int kMlp = 4;       // independent 16-byte loads in flight per thread

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// MSE gradient + 1x1 output convolution backward.  blockDim = (32, 4, kLgGroups): x -> position, y -> plane,
// z -> quarter of the block's kChunk positions (same shape as mask_reduce_kernel below: one round trip of loads
// per thread, shared-memory combine, one set of atomics per block).
//   g = 2*(eps - noise)/n ; go4[c] = out.weight[c]*g ; d out.weight[c] += g*h4[c] ; d out.bias += g
// and, in the same pass, what rb4 needs from its output gradient (src/mnist.py:65-66 backward):
//   gc4 = go4 (.) relu_mask2 ; d rb4.skip.bias[c] += sum go4[c] ; d rb4.conv2.bias[c] += sum gc4[c]
constexpr int kLgGroups = 4;
__global__ void __launch_bounds__(32 * 4 * kLgGroups, 2)   // <= 64 registers: two blocks per SM (86 registers left one, 24 % occupancy)
loss_grad_kernel(const float* __restrict__ eps, const float* __restrict__ noise,
                 const uint8_t* __restrict__ h4, int64_t ps, const float* __restrict__ wo,
                 uint8_t* __restrict__ go, float* __restrict__ d_wo, float* __restrict__ d_bo,
                 float* __restrict__ loss, int batch, int64_t npos, float inv_n,
                 const uint32_t* __restrict__ mask, uint8_t* __restrict__ gc, float* __restrict__ d_plain,
                 float* __restrict__ d_masked) {
    pdl_wait();   // PDL (common.cuh): first statement, nothing before it touches global memory
    pdl_launch_dependents();
    using G = Geo<28>;
    constexpr int kIter = kChunk / (32 * kLgGroups);
    __shared__ float s_red[kLgGroups][4][26];   // [group][plane][d_wo 0..7 | loss 8 | d_bo 9 | plain 10..17 | masked 18..25]
    const int lane = threadIdx.x, j = threadIdx.y, zg = threadIdx.z;
    float w[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        w[k] = __ldg(wo + j * 8 + k);
        acc[k] = 0.f;
    }
    float lsum = 0.f, gsum = 0.f;
    float sp[8], sm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sp[k] = sm[k] = 0.f;
    const int64_t base = (int64_t)blockIdx.x * kChunk + zg * (kIter * 32);
    float dv[kIter];
    uint4 hvs[kIter];
    uint32_t words[kIter];
    bool ok[kIter];
#pragma unroll
    for (int u = 0; u < kIter; ++u) {   // issue all loads before using any
        const int64_t pos = base + u * 32 + lane;
        const int b = (int)((uint32_t)pos / (uint32_t)G::S);   // positions fit 32 bits: division by a constant is a multiply-shift
        const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
        const int r = rem / G::Wp, c = rem - r * G::Wp;
        ok[u] = pos < npos && b < batch && r >= 1 && c < G::W;
        dv[u] = 0.f;
        hvs[u] = make_uint4(0, 0, 0, 0);
        words[u] = 0;
        if (ok[u]) {
            words[u] = mask[pos];   // 32 channels: one word per position
            const int64_t i = (int64_t)b * 784 + (r - 1) * 28 + c;
            dv[u] = __ldg(eps + i) - __ldg(noise + i);
            hvs[u] = *reinterpret_cast<const uint4*>(h4 + j * ps + (pos + G::GUARD) * 16);
        }
    }
#pragma unroll
    for (int u = 0; u < kIter; ++u) {
        const int64_t pos = base + u * 32 + lane;
        if (pos >= npos) continue;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (ok[u]) {
            const float d = dv[u];
            const float g = 2.0f * d * inv_n;
            const uint32_t* hw = &hvs[u].x;
            uint32_t* ow = &o.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 h = unpack_bf16x2(hw[k]);
                acc[2 * k] = fmaf(g, h.x, acc[2 * k]);
                acc[2 * k + 1] = fmaf(g, h.y, acc[2 * k + 1]);
                ow[k] = pack_bf16x2(w[2 * k] * g, w[2 * k + 1] * g);
            }
            if (j == 0) {
                lsum = fmaf(d * inv_n, d, lsum);
                gsum += g;
            }
        }
        *reinterpret_cast<uint4*>(go + j * ps + (pos + G::GUARD) * 16) = o;
        // the masked copy and the two bias reductions, on the bf16 values just stored
        const uint32_t bits = (words[u] >> (j * 8)) & 0xffu;
        const uint32_t* ow2 = &o.x;
        uint4 om;
        uint32_t* mwp = &om.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = unpack_bf16x2(ow2[k]);
            const float m0 = (bits >> (2 * k)) & 1u ? f.x : 0.f;
            const float m1 = (bits >> (2 * k + 1)) & 1u ? f.y : 0.f;
            sp[2 * k] += f.x;
            sp[2 * k + 1] += f.y;
            sm[2 * k] += m0;
            sm[2 * k + 1] += m1;
            mwp[k] = pack_bf16x2(m0, m1);
        }
        *reinterpret_cast<uint4*>(gc + j * ps + (pos + G::GUARD) * 16) = om;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float v = warp_sum(acc[k]), x = warp_sum(sp[k]), y = warp_sum(sm[k]);
        if (lane == 0) {
            s_red[zg][j][k] = v;
            s_red[zg][j][10 + k] = x;
            s_red[zg][j][18 + k] = y;
        }
    }
    lsum = warp_sum(lsum);
    gsum = warp_sum(gsum);
    if (lane == 0) {
        s_red[zg][j][8] = lsum;   // zero except on plane 0
        s_red[zg][j][9] = gsum;
    }
    __syncthreads();
    if (zg == 0 && lane < 26) {
        float v = 0.f;
#pragma unroll
        for (int z = 0; z < kLgGroups; ++z) v += s_red[z][j][lane];
        if (lane < 8) atomicAdd(d_wo + j * 8 + lane, v);
        else if (lane < 10) { if (j == 0) atomicAdd(lane == 8 ? loss : d_bo, v); }
        else if (lane < 18) atomicAdd(d_plain + j * 8 + (lane - 10), v);
        else atomicAdd(d_masked + j * 8 + (lane - 18), v);
    }
}

// out = g (.) relu_mask, with the per-channel reductions the block needs:
//   d_plain[c] += sum g[c]            (time_emb.bias or skip.bias gradient; nullable)
//   d_ts[c]    += sum g[c]*t/1000     (time_emb.weight gradient; nullable)
//   d_masked[c]+= sum g[c]*mask       (conv bias gradient; nullable)
// where g is a block's output gradient, produced on the fly from its sources (SRC) and also stored unmasked
// (`gout`, for the block's skip path) unless it already exists as a tensor:
//   MR_PLANES      g = in[pos]                                               (gout unused)
//   MR_UPSAMPLE_T  g = 2x2 sum of the 28x28 planes `in` (transpose of the nearest x2 upsample, src/mnist.py:83)
//   MR_POOL_T      g = in[pos] + 0.25 * in2[pos14(y/2, x/2)]   (concat slice + transpose of avg_pool2d, :80)
// blockDim = (32, CH/8, kMrGroups): x -> position, y -> channel plane, z -> quarter of the block's kChunk
// positions.  Every thread has its loads in flight at once (one round trip per block), the four position
// groups are combined in shared memory, and 24 lanes per plane issue the block's atomics in parallel
// (ncu, B=512: the one-warp-per-plane version ran at 27 % occupancy and 42 % issue, 27 us per launch).
// `out` may alias `in` for MR_PLANES.
enum : int { MR_PLANES = 0, MR_UPSAMPLE_T = 1, MR_POOL_T = 2 };
struct MaskReduceArgs {
    const uint8_t* in;      // source planes (MR_UPSAMPLE_T / MR_POOL_T: 28x28 geometry)
    int64_t in_ps;
    const uint8_t* in2;     // MR_POOL_T: 14x14 planes
    int64_t in2_ps;
    const uint32_t* mask;   // one uint32 per 32 channels per position, null = all ones
    int64_t mask_stride;
    uint8_t* out;           // g (.) mask, geometry W
    uint8_t* gout;          // g unmasked (nullable)
    int64_t out_ps;
    const int64_t* t;
    int batch;
    int64_t npos;
    float* d_plain;
    float* d_ts;
    float* d_masked;
};
constexpr int kMrGroups = 4;
static_assert(kChunk == kMrGroups * kMlp * 32, "one batch of loads per thread");
template <int SRC, int W>
__global__ void __launch_bounds__(32 * 8 * kMrGroups) mask_reduce_kernel(const MaskReduceArgs a) {
    pdl_wait();   // PDL (common.cuh): first statement, nothing before it touches global memory
    pdl_launch_dependents();
    using G = Geo<W>;
    __shared__ float s_red[kMrGroups][8][24];   // [position group][plane][plain 0..7 | ts 8..15 | masked 16..23]
    const int lane = threadIdx.x, j = threadIdx.y, zg = threadIdx.z;
    float sp[8], st[8], sm[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) sp[k] = st[k] = sm[k] = 0.f;
    const int64_t base = (int64_t)blockIdx.x * kChunk + zg * (kMlp * 32);
    uint4 gvs[kMlp];
    uint32_t words[kMlp];
#pragma unroll
    for (int u = 0; u < kMlp; ++u) {   // issue all loads before using any
        const int64_t pos = base + u * 32 + lane;
        gvs[u] = make_uint4(0, 0, 0, 0);
        words[u] = 0;
        if (pos >= a.npos) continue;
        words[u] = a.mask ? a.mask[(j >> 2) * a.mask_stride + pos] : 0xffffffffu;
        if constexpr (SRC == MR_PLANES) {
            gvs[u] = *reinterpret_cast<const uint4*>(a.in + j * a.in_ps + (pos + G::GUARD) * 16);
        } else {
            const int b = (int)((uint32_t)pos / (uint32_t)G::S);   // positions fit 32 bits: multiply-shift
            const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
            const int r = rem / G::Wp, c = rem - r * G::Wp;
            if (b < a.batch && r >= 1 && c < G::W) {
                uint32_t* ow = &gvs[u].x;
                if constexpr (SRC == MR_UPSAMPLE_T) {
                    static_assert(SRC != MR_UPSAMPLE_T || W == 14, "2x2 sum: 28x28 -> 14x14");
                    using GI = Geo<28>;
                    const int64_t p00 = (int64_t)b * GI::S + (2 * (r - 1) + 1) * GI::Wp + 2 * c;
                    const uint8_t* src = a.in + j * a.in_ps + (p00 + GI::GUARD) * 16;
                    const uint4 q0 = *reinterpret_cast<const uint4*>(src);
                    const uint4 q1 = *reinterpret_cast<const uint4*>(src + 16);
                    const uint4 q2 = *reinterpret_cast<const uint4*>(src + GI::Wp * 16);
                    const uint4 q3 = *reinterpret_cast<const uint4*>(src + GI::Wp * 16 + 16);
                    const uint32_t *a0 = &q0.x, *a1 = &q1.x, *a2 = &q2.x, *a3 = &q3.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f0 = unpack_bf16x2(a0[k]), f1 = unpack_bf16x2(a1[k]);
                        const float2 f2 = unpack_bf16x2(a2[k]), f3 = unpack_bf16x2(a3[k]);
                        ow[k] = pack_bf16x2(f0.x + f1.x + f2.x + f3.x, f0.y + f1.y + f2.y + f3.y);
                    }
                } else {
                    static_assert(SRC != MR_POOL_T || W == 28, "pool transpose: 14x14 -> 28x28");
                    using G14 = Geo<14>;
                    const int64_t p14 = (int64_t)b * G14::S + ((r - 1) / 2 + 1) * G14::Wp + c / 2;
                    const uint4 av = *reinterpret_cast<const uint4*>(a.in + j * a.in_ps + (pos + G::GUARD) * 16);
                    const uint4 pv = *reinterpret_cast<const uint4*>(a.in2 + j * a.in2_ps + (p14 + G14::GUARD) * 16);
                    const uint32_t *aw = &av.x, *pw = &pv.x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = unpack_bf16x2(aw[k]), h = unpack_bf16x2(pw[k]);
                        ow[k] = pack_bf16x2(fmaf(0.25f, h.x, f.x), fmaf(0.25f, h.y, f.y));
                    }
                }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kMlp; ++u) {
        const int64_t pos = base + u * 32 + lane;
        if (pos >= a.npos) continue;
        const uint4 gv = gvs[u];
        if (SRC != MR_PLANES && a.gout) *reinterpret_cast<uint4*>(a.gout + j * a.out_ps + (pos + G::GUARD) * 16) = gv;
        const uint32_t bits = (words[u] >> ((j & 3) * 8)) & 0xffu;
        float ts = 0.f;
        if (a.d_ts) {
            const int b = (int)((uint32_t)pos / (uint32_t)G::S);
            ts = b < a.batch ? (float)__ldg(a.t + b) / 1000.0f : 0.f;
        }
        const uint32_t* gw = &gv.x;
        uint4 o;
        uint32_t* ow = &o.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 f = unpack_bf16x2(gw[k]);
            const float m0 = (bits >> (2 * k)) & 1u ? f.x : 0.f;
            const float m1 = (bits >> (2 * k + 1)) & 1u ? f.y : 0.f;
            sp[2 * k] += f.x;
            sp[2 * k + 1] += f.y;
            st[2 * k] = fmaf(f.x, ts, st[2 * k]);
            st[2 * k + 1] = fmaf(f.y, ts, st[2 * k + 1]);
            sm[2 * k] += m0;
            sm[2 * k + 1] += m1;
            ow[k] = pack_bf16x2(m0, m1);
        }
        if (a.out) *reinterpret_cast<uint4*>(a.out + j * a.out_ps + (pos + G::GUARD) * 16) = o;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x = warp_sum(sp[k]), y = warp_sum(st[k]), z = warp_sum(sm[k]);
        if (lane == 0) {
            s_red[zg][j][k] = x;
            s_red[zg][j][8 + k] = y;
            s_red[zg][j][16 + k] = z;
        }
    }
    __syncthreads();
    if (zg == 0 && lane < 24) {
        float v = 0.f;
#pragma unroll
        for (int z = 0; z < kMrGroups; ++z) v += s_red[z][j][lane];
        float* dst = lane < 8 ? a.d_plain : lane < 16 ? a.d_ts : a.d_masked;
        if (dst) atomicAdd(dst + j * 8 + (lane & 7), v);
    }
}

// rb1.conv1 (Cin = 1) and rb1.skip (1x1, Cin = 1) weight gradients:
//   dW1[co][tap] += sum gc[pos][co] * x[pos + off(tap)]      dWs[co] += sum go[pos][co] * x[pos]
// block = 256 threads = 8 channel groups (4 channels) x 32 position slices (4 positions) of a 128-position
// tile; a thread keeps its 4 x 9 + 4 partial sums in registers over all the block's tiles (11 shared-memory
// loads per 40 FMAs; one (co, tap) pair per thread was 2 loads per FMA and shared-memory bound at 50 us for
// B = 512), then the slices are combined once: shuffles inside a warp, shared memory across the 8 warps.
constexpr int kR1Stride = 36;   // floats per staged position (32 channels, padded so rows stay 16-byte aligned)
__global__ void __launch_bounds__(256)
rb1_wgrad_kernel(const uint8_t* __restrict__ gc, const uint8_t* __restrict__ go, int64_t ps,
                 const float* __restrict__ x, float* __restrict__ d_w1, float* __restrict__ d_ws,
                 int batch, int nt) {
    pdl_wait();   // PDL (common.cuh): first statement, nothing before it touches global memory
    pdl_launch_dependents();
    using G = Geo<28>;
    constexpr int HL = G::Wp + 1;   // largest tap offset
    __shared__ __align__(16) float s_gc[128 * kR1Stride], s_go[128 * kR1Stride];
    __shared__ float s_x[128 + 2 * HL];   // x at positions p0-HL .. p0+127+HL in the padded geometry (0 at pads): every tap
                                          // is then a fixed offset into this window, no per-tap bounds checks
    const int tid = threadIdx.x;
    const int cg = tid & 7, sl = tid >> 3;
    float acc[4][9], accs[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        accs[c] = 0.f;
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[c][k] = 0.f;
    }
    for (int tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        __syncthreads();
        for (int i = tid; i < 128 * 4; i += 256) {
            const int p = i & 127, j = i >> 7;
            const int64_t pos = (int64_t)tile * 128 + p;
            const uint4 v = *reinterpret_cast<const uint4*>(gc + j * ps + (pos + G::GUARD) * 16);
            const uint4 u = *reinterpret_cast<const uint4*>(go + j * ps + (pos + G::GUARD) * 16);
            const uint32_t *vw = &v.x, *uw = &u.x;
            float fv[8], fu[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = unpack_bf16x2(vw[k]), h = unpack_bf16x2(uw[k]);
                fv[2 * k] = f.x; fv[2 * k + 1] = f.y;
                fu[2 * k] = h.x; fu[2 * k + 1] = h.y;
            }
            float4* dg = reinterpret_cast<float4*>(s_gc + p * kR1Stride + j * 8);
            float4* du = reinterpret_cast<float4*>(s_go + p * kR1Stride + j * 8);
            dg[0] = make_float4(fv[0], fv[1], fv[2], fv[3]);
            dg[1] = make_float4(fv[4], fv[5], fv[6], fv[7]);
            du[0] = make_float4(fu[0], fu[1], fu[2], fu[3]);
            du[1] = make_float4(fu[4], fu[5], fu[6], fu[7]);
        }
        if (tid < 128 + 2 * HL) {
            const int pos = tile * 128 - HL + tid;
            float v = 0.f;
            if (pos >= 0) {
                const int b = (int)((uint32_t)pos / (uint32_t)G::S);   // positions fit 32 bits: division by a constant is a multiply-shift
                const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
                const int r = rem / G::Wp, c = rem - r * G::Wp;
                if (b < batch && r >= 1 && c < G::W) v = __ldg(x + (int64_t)b * 784 + (r - 1) * 28 + c);
            }
            s_x[tid] = v;
        }
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            const int p = sl * 4 + pp;
            const float4 g4 = *reinterpret_cast<const float4*>(s_gc + p * kR1Stride + cg * 4);
            const float4 o4 = *reinterpret_cast<const float4*>(s_go + p * kR1Stride + cg * 4);
            const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, ov[4] = {o4.x, o4.y, o4.z, o4.w};
            // gradients at pad positions are stored as zeros, so whatever the window holds around them is harmless
            float xv[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) xv[k] = s_x[HL + p + (k / 3 - 1) * G::Wp + (k % 3 - 1)];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
#pragma unroll
                for (int k = 0; k < 9; ++k) acc[c][k] = fmaf(gv[c], xv[k], acc[c][k]);
                accs[c] = fmaf(ov[c], xv[4], accs[c]);
            }
        }
    }
    // combine the 32 position slices: lanes differing in bits 3,4 share cg; then the 8 warps through smem
    __syncthreads();
    float* s_red = s_gc;   // reuse: [8 warps][8 cg][40]
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            float v = k < 9 ? acc[c][k] : accs[c];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (lane < 8) s_red[(warp * 8 + lane) * 40 + c * 10 + k] = v;
        }
    }
    __syncthreads();
    for (int i = tid; i < 320; i += 256) {   // i = cg*40 + c*10 + k
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += s_red[w * 320 + i];
        const int g = i / 40, c = (i % 40) / 10, k = i % 10;
        const int co = g * 4 + c;
        if (k < 9) atomicAdd(d_w1 + co * 9 + k, v);
        else atomicAdd(d_ws + co, v);
    }
}

// ---------------------------------------------------------------------------------------------
// fused AdamW over the flat buffers — torch.optim.AdamW's update (decoupled weight decay):
//   p *= 1 - lr*wd; m = lerp(m, g, 1-b1); v = b2*v + (1-b2)*g*g;
//   p -= (lr / (1-b1^k)) * m / (sqrt(v)/sqrt(1-b2^k) + eps)
// `step` lives on the device (1-based k of THIS update) so a captured graph can be replayed.
// ---------------------------------------------------------------------------------------------
//
// PEER = true is the data-parallel form, "gradient all-reduce + AdamW" in one kernel (peer.cu has the buffer
// layout): every rank's gradient of step k sits in slot k & 1 of a buffer all ranks have mapped; the kernel
//   1. announces "my gradient k is complete" by storing k into flags[rank] on every peer (release, system scope),
//   2. waits until flags[r] >= k for every r in its OWN buffer (acquire, system scope; bounded spin, then trap),
//   3. sums the world gradients with loads over NVLink in rank order - the same order on every rank, so the
//      replicas stay bit-identical - and applies the update.
// No rank can be more than one step ahead of a peer that is still reading (it would need that peer's flag for
// the next step first), which is why two slots are enough.
constexpr int kMaxPeers = 8;
struct PeerBases { const uint8_t* base[kMaxPeers]; };

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

template <bool PEER>
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
             float grad_scale, const int64_t* __restrict__ step, PeerBases peers, int world, int rank,
             int64_t grad_off0, int64_t grad_slot_floats, long long timeout_cycles) {
    pdl_wait();   // PDL (common.cuh): first statement, nothing before it touches global memory
    pdl_launch_dependents();
    __shared__ float s_c[2];
    const int64_t kstep = *step;
    if (threadIdx.x == 0) {
        const double k = (double)kstep;
        s_c[0] = (float)((double)lr / (1.0 - pow((double)b1, k)));   // step size
        s_c[1] = (float)sqrt(1.0 - pow((double)b2, k));               // sqrt(bias_correction2)
    }
    if constexpr (PEER) {
        if ((int)threadIdx.x < world) {
            const int r = threadIdx.x;
            if (blockIdx.x == 0) {
                __threadfence_system();   // this rank's gradient (earlier kernels of the stream) before the flag
                st_release_sys(reinterpret_cast<uint64_t*>(const_cast<uint8_t*>(peers.base[r])) + rank, (uint64_t)kstep);
            }
            const uint64_t* mine = reinterpret_cast<const uint64_t*>(peers.base[rank]) + r;
            const long long t0 = clock64();
            while (ld_acquire_sys(mine) < (uint64_t)kstep) {
                if (clock64() - t0 > timeout_cycles) {   // a peer died; fail loudly instead of hanging (peer_timeout_cycles)
                    printf("tdm adamw_peer: rank %d timed out waiting for rank %d at step %lld\n", rank, r, (long long)kstep);
                    __trap();
                }
                __nanosleep(200);
            }
        }
    }
    __syncthreads();
    const float step_size = s_c[0], sbc2 = s_c[1];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gsum;
    if constexpr (PEER) {
        const int64_t off = grad_off0 + (kstep & 1) * grad_slot_floats * 4;
        gsum = 0.f;
        for (int r = 0; r < world; ++r) gsum += ld_relaxed_sys(reinterpret_cast<const float*>(peers.base[r] + off) + i);
    } else {
        gsum = g[i];
    }
    const float gi = gsum * grad_scale;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    const float denom = sqrtf(vi) / sbc2 + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
}

// ---------------------------------------------------------------------------------------------
// orchestration
// ---------------------------------------------------------------------------------------------
template <int SRC, int W>
static int mask_reduce(const MaskReduceArgs& a, int ch, cudaStream_t st) {
    const unsigned grid = (unsigned)((a.npos + kChunk - 1) / kChunk);
    launch_pdl(mask_reduce_kernel<SRC, W>, dim3(grid), dim3(32, ch / 8, kMrGroups), 0, st, a);
    TDM_CHECK_LAUNCH("mask_reduce");
    return TDM_OK;
}

// 3x3 weight gradients leave the wgrad kernels tap-major ([tap][Cout][Cin], contiguous along Cin so the flush is
// vector atomics); this puts them into the flat gradient in the parameter order (OIHW) and re-zeroes the scratch.
struct UnpermJob { int off, cout, cin; };
constexpr int kUnpermJobs = 7;
__constant__ UnpermJob c_unperm[kUnpermJobs] = {
    {P::rb1_c2w, 32, 32}, {P::rb2_c1w, 64, 32}, {P::rb2_c2w, 64, 64}, {P::rb3_c1w, 64, 64},
    {P::rb3_c2w, 64, 64}, {P::rb4_c1w, 32, 96}, {P::rb4_c2w, 32, 32}};
static_assert(P::rb1_c2w % 4 == 0 && P::rb2_c1w % 4 == 0 && P::rb2_c2w % 4 == 0 && P::rb3_c1w % 4 == 0 &&
              P::rb3_c2w % 4 == 0 && P::rb4_c1w % 4 == 0 && P::rb4_c2w % 4 == 0 && P::rb2_sw % 4 == 0 &&
              P::rb4_sw % 4 == 0, "16-byte vector atomics need 4-float aligned tensors");

__global__ void wgrad_unpermute_kernel(float* __restrict__ scr, float* __restrict__ dflat) {
    pdl_wait();   // PDL (common.cuh): first statement, nothing before it touches global memory
    pdl_launch_dependents();
    const UnpermJob j = c_unperm[blockIdx.y];
    const int n = j.cout * j.cin * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int tap = i % 9, rest = i / 9;           // rest = co * cin + ci
        float* s = scr + j.off + tap * (j.cout * j.cin) + rest;
        dflat[j.off + i] = *s;
        *s = 0.f;
    }
}

static int unet_backward_impl(const uint8_t* wp, const float* x, const int64_t* t, const float* noise,
                              const float* eps, float* dflat, float* loss, uint8_t* ws,
                              int64_t ws_bytes, int64_t batch, cudaStream_t st) {
    TDM_CHECK_ARG(wp && x && t && noise && eps && dflat && loss && ws, "unet_backward: null pointer");
    TDM_CHECK_ARG(batch > 0 && batch <= (1 << 20), "unet_backward: batch out of range");
    const UNetWs L = make_ws(batch, true);
    TDM_CHECK_ARG(ws_bytes >= L.total, "unet_backward: workspace too small (%lld < %lld)",
                  (long long)ws_bytes, (long long)L.total);
    const float* fp = reinterpret_cast<const float*>(wp + WP::flat);
    const int B = (int)batch;
    const int nt28 = (int)L.nt28, nt14 = (int)L.nt14;
    const int H28 = Geo<28>::GUARD, H14 = Geo<14>::GUARD, S28 = Geo<28>::S, S14 = Geo<14>::S;
    auto M = [&](int64_t off) { return reinterpret_cast<const uint32_t*>(ws + off); };
    int rc;
    ConvArgs c{};
    WgradArgs w{};
    float* gscr = reinterpret_cast<float*>(ws + L.gscr);

    TDM_CHECK_CUDA(cudaMemsetAsync(dflat, 0, sizeof(float) * P::count, st));
    TDM_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));

    // ---- loss and out conv ---------------------------------------------------------------
    launch_pdl(loss_grad_kernel, dim3((unsigned)((L.np28 + kChunk - 1) / kChunk)), dim3(32, 4, kLgGroups), 0, st, 
        eps, noise, ws + L.h4, L.ps28, fp + P::out_w, ws + L.go28, dflat + P::out_w, dflat + P::out_b,
        loss, B, L.np28, 1.0f / (float)(batch * 784), M(L.m2_4), ws + L.gc28, dflat + P::rb4_sb, dflat + P::rb4_c2b);
    TDM_CHECK_LAUNCH("loss_grad");

    // ---- rb4: x_in = cat (96), h = t4, g_out = go28 ----------------------------------------
    // (rb4's output gradient was masked and reduced by loss_grad_kernel above: gc28, skip.bias, conv2.bias)
    MaskReduceArgs mr{};
    w = WgradArgs{ws + L.gc28, L.ps28, ws + L.t4, L.ps28, gscr + P::rb4_c2w, nt28};
    if ((rc = launch_wgrad_dup<28, 32, 32>(w, L.np28, st, "wgrad_rb4_c2"))) return rc;
    c = ConvArgs{}; c.t = t; c.batch = B; c.np = (int)L.np28;
    c.in = ws + L.gc28; c.in_ps = L.ps28; c.w = wp + WP::d_rb4_c2; c.out = ws + L.gh28; c.out_ps = L.ps28;
    // conv2's data gradient leaves the kernel already multiplied by conv1's ReLU mask, with the time-embedding and
    // conv1-bias gradients reduced in the same epilogue (EPI_PLAIN_MASK): the buffer holds d loss / d conv1 pre-activation
    c.mask = const_cast<uint32_t*>(M(L.m1_4)); c.mask_stride = L.np28;
    c.red_plain = dflat + P::rb4_tb; c.red_ts = dflat + P::rb4_tw; c.red_masked = dflat + P::rb4_c1b;
    if ((rc = launch_conv<28, 32, 32, EPI_PLAIN_MASK, false, 9, KX::rb4c2>(c, st, "dgrad_rb4_c2"))) return rc;
    // conv1's and the 1x1 skip's weight gradients share the concat tile: the skip's comes out of the spare quarter
    w = WgradArgs{ws + L.gh28, L.ps28, ws + L.cat, L.ps28, gscr + P::rb4_c1w, nt28};
    w.g2 = ws + L.go28; w.dws = dflat + P::rb4_sw;
    if ((rc = launch_wgrad_dup<28, 32, 96, true>(w, L.np28, st, "wgrad_rb4_c1_skip"))) return rc;
    c = ConvArgs{}; c.t = t; c.batch = B; c.np = (int)L.np28;
    // gradient w.r.t. the concat input = conv1^T(gc1) + skip^T(g_out): one kernel, the 1x1 skip transpose accumulates
    // into the same tile from a second input (its packed weights sit right behind conv1's)
    static_assert(WP::d_rb4_sk == WP::d_rb4_c1 + WP::conv_bytes(32, 96), "skip image must follow the 3x3 image");
    c.in = ws + L.gh28; c.in_ps = L.ps28; c.in3 = ws + L.go28; c.in3_ps = L.ps28;
    c.w = wp + WP::d_rb4_c1; c.out = ws + L.gcat; c.out_ps = L.ps28;
    if ((rc = launch_conv<28, 32, 96, EPI_PLAIN, false, 9, 0, 0, false, 32>(c, st, "dgrad_rb4_c1_skip"))) return rc;

    // ---- through the concat: channels 0..63 -> up(h3)^T -> g_out of rb3 ---------------------
    // ---- rb3: x_in = h2, h = t3, identity skip, g_out = go14a -------------------------------
    // one pass: go14a = up^T(gcat[0:64]) (kept for the identity skip), gc14 = go14a (.) mask, conv2-bias gradient
    mr = MaskReduceArgs{};
    mr.in = ws + L.gcat; mr.in_ps = L.ps28; mr.mask = M(L.m2_3); mr.mask_stride = L.np14;
    mr.out = ws + L.gc14; mr.gout = ws + L.go14a; mr.out_ps = L.ps14; mr.t = t; mr.batch = B; mr.npos = L.np14;
    mr.d_masked = dflat + P::rb3_c2b;
    if ((rc = mask_reduce<MR_UPSAMPLE_T, 14>(mr, 64, st))) return rc;
    w = WgradArgs{ws + L.gc14, L.ps14, ws + L.t3, L.ps14, gscr + P::rb3_c2w, nt14};
    if ((rc = launch_wgrad_dup<14, 64, 64>(w, L.np14, st, "wgrad_rb3_c2"))) return rc;
    c = ConvArgs{}; c.t = t; c.batch = B; c.np = (int)L.np14;
    c.in = ws + L.gc14; c.in_ps = L.ps14; c.w = wp + WP::d_rb3_c2; c.out = ws + L.gh14; c.out_ps = L.ps14;
    // conv2's data gradient leaves the kernel already multiplied by conv1's ReLU mask, with the time-embedding and
    // conv1-bias gradients reduced in the same epilogue (EPI_PLAIN_MASK): the buffer holds d loss / d conv1 pre-activation
    c.mask = const_cast<uint32_t*>(M(L.m1_3)); c.mask_stride = L.np14;
    c.red_plain = dflat + P::rb3_tb; c.red_ts = dflat + P::rb3_tw; c.red_masked = dflat + P::rb3_c1b;
    if ((rc = launch_conv<14, 64, 64, EPI_PLAIN_MASK, false, 9, KX::rb3c2>(c, st, "dgrad_rb3_c2"))) return rc;
    w = WgradArgs{ws + L.gh14, L.ps14, ws + L.h2, L.ps14, gscr + P::rb3_c1w, nt14};
    if ((rc = launch_wgrad_dup<14, 64, 64>(w, L.np14, st, "wgrad_rb3_c1"))) return rc;
    c = ConvArgs{}; c.t = t; c.batch = B; c.np = (int)L.np14;
    c.in = ws + L.gh14; c.in_ps = L.ps14; c.w = wp + WP::d_rb3_c1; c.res = ws + L.go14a; c.res_ps = L.ps14;
    c.out = ws + L.go14b; c.out_ps = L.ps14;   // g_out of rb2
    if ((rc = launch_conv<14, 64, 64, EPI_PLAIN, false, 9, KX::rb3c1>(c, st, "dgrad_rb3_c1"))) return rc;

    // ---- rb2: x_in = p1 (32), h = t2, skip 32->64, g_out = go14b ----------------------------
    mr = MaskReduceArgs{};
    mr.in = ws + L.go14b; mr.in_ps = L.ps14; mr.mask = M(L.m2_2); mr.mask_stride = L.np14;
    mr.out = ws + L.gc14; mr.out_ps = L.ps14; mr.t = t; mr.batch = B; mr.npos = L.np14;
    mr.d_plain = dflat + P::rb2_sb; mr.d_masked = dflat + P::rb2_c2b;
    if ((rc = mask_reduce<MR_PLANES, 14>(mr, 64, st))) return rc;
    w = WgradArgs{ws + L.gc14, L.ps14, ws + L.t2, L.ps14, gscr + P::rb2_c2w, nt14};
    if ((rc = launch_wgrad_dup<14, 64, 64>(w, L.np14, st, "wgrad_rb2_c2"))) return rc;
    c = ConvArgs{}; c.t = t; c.batch = B; c.np = (int)L.np14;
    c.in = ws + L.gc14; c.in_ps = L.ps14; c.w = wp + WP::d_rb2_c2; c.out = ws + L.gh14; c.out_ps = L.ps14;
    // conv2's data gradient leaves the kernel already multiplied by conv1's ReLU mask, with the time-embedding and
    // conv1-bias gradients reduced in the same epilogue (EPI_PLAIN_MASK): the buffer holds d loss / d conv1 pre-activation
    c.mask = const_cast<uint32_t*>(M(L.m1_2)); c.mask_stride = L.np14;
    c.red_plain = dflat + P::rb2_tb; c.red_ts = dflat + P::rb2_tw; c.red_masked = dflat + P::rb2_c1b;
    if ((rc = launch_conv<14, 64, 64, EPI_PLAIN_MASK, false, 9, KX::rb2c2>(c, st, "dgrad_rb2_c2"))) return rc;
    w = WgradArgs{ws + L.gh14, L.ps14, ws + L.p1, L.ps14, gscr + P::rb2_c1w, nt14};
    if ((rc = launch_wgrad_dup<14, 64, 32>(w, L.np14, st, "wgrad_rb2_c1"))) return rc;
    w = WgradArgs{ws + L.go14b, L.ps14, ws + L.p1, L.ps14, dflat + P::rb2_sw, nt14};
    if ((rc = launch_wgrad<14, 64, 32, 1>(w, st, "wgrad_rb2_skip"))) return rc;
    c = ConvArgs{}; c.t = t; c.batch = B; c.np = (int)L.np14;
    static_assert(WP::d_rb2_sk == WP::d_rb2_c1 + WP::conv_bytes(64, 32) && KX::rb2c1 == 0, "skip image must follow the 3x3 image");
    c.in = ws + L.gh14; c.in_ps = L.ps14; c.in3 = ws + L.go14b; c.in3_ps = L.ps14;
    c.w = wp + WP::d_rb2_c1; c.out = ws + L.gp1; c.out_ps = L.ps14;
    if ((rc = launch_conv<14, 64, 32, EPI_PLAIN, false, 9, 0, 0, false, 64>(c, st, "dgrad_rb2_c1_skip"))) return rc;

    // ---- h1 receives: concat channels 64..95 + avg-pool transpose of g_p1 -------------------
    // (fused below: go28 = gcat[64:96] + pool^T(gp1) is produced by the pass that also masks it)

    // ---- rb1: x_in = x (1 channel), h = t1, skip 1->32, g_out = go28 ------------------------
    mr = MaskReduceArgs{};
    mr.in = ws + L.gcat + 8 * L.ps28; mr.in_ps = L.ps28; mr.in2 = ws + L.gp1; mr.in2_ps = L.ps14;
    mr.mask = M(L.m2_1); mr.mask_stride = L.np28;
    mr.out = ws + L.gc28; mr.gout = ws + L.go28; mr.out_ps = L.ps28; mr.t = t; mr.batch = B; mr.npos = L.np28;
    mr.d_plain = dflat + P::rb1_sb; mr.d_masked = dflat + P::rb1_c2b;
    if ((rc = mask_reduce<MR_POOL_T, 28>(mr, 32, st))) return rc;
    w = WgradArgs{ws + L.gc28, L.ps28, ws + L.t1, L.ps28, gscr + P::rb1_c2w, nt28};
    if ((rc = launch_wgrad_dup<28, 32, 32>(w, L.np28, st, "wgrad_rb1_c2"))) return rc;
    c = ConvArgs{}; c.t = t; c.batch = B; c.np = (int)L.np28;
    c.in = ws + L.gc28; c.in_ps = L.ps28; c.w = wp + WP::d_rb1_c2; c.out = ws + L.gh28; c.out_ps = L.ps28;
    // conv2's data gradient leaves the kernel already multiplied by conv1's ReLU mask, with the time-embedding and
    // conv1-bias gradients reduced in the same epilogue (EPI_PLAIN_MASK): the buffer holds d loss / d conv1 pre-activation
    c.mask = const_cast<uint32_t*>(M(L.m1_1)); c.mask_stride = L.np28;
    c.red_plain = dflat + P::rb1_tb; c.red_ts = dflat + P::rb1_tw; c.red_masked = dflat + P::rb1_c1b;
    if ((rc = launch_conv<28, 32, 32, EPI_PLAIN_MASK, false, 9, KX::rb1c2>(c, st, "dgrad_rb1_c2"))) return rc;
    {
        // 4 blocks per SM requested although 3 are resident (80 registers): sizing the grid to exactly one resident
        // wave (3 x SMs) is 1-2 % faster when every SM takes its 3 blocks at once, but under programmatic dependent
        // launch the step time turned bimodal (presumably some SMs start with fewer than 3 and the left-over blocks
        // then cost a whole extra block time:
        // 0.644 / 0.707 ms training steps at B = 512); shorter blocks in 1.33 waves are the robust choice.
        const int grid = nt28 < 4 * num_sms() ? nt28 : 4 * num_sms();
        launch_pdl(rb1_wgrad_kernel, dim3(grid), dim3(256), 0, st, ws + L.gh28, ws + L.go28, L.ps28, x, dflat + P::rb1_c1w,
                                               dflat + P::rb1_sw, B, nt28);
        TDM_CHECK_LAUNCH("rb1_wgrad");
    }
    launch_pdl(wgrad_unpermute_kernel, dim3(16, kUnpermJobs), dim3(256), 0, st, gscr, dflat);
    TDM_CHECK_LAUNCH("wgrad_unpermute");
    return TDM_OK;
}

}  // namespace tdm

using namespace tdm;

extern "C" int tdm_unet_forward_train(const void* wpack, const float* x, const int64_t* t,
                                      float* eps_out, void* workspace, int64_t workspace_bytes,
                                      int64_t batch, void* stream) {
    StepArgs sa;
    sa.train = 1;
    return unet_forward_impl(reinterpret_cast<const uint8_t*>(wpack), x, t, eps_out,
                             reinterpret_cast<uint8_t*>(workspace), workspace_bytes, batch, sa,
                             (cudaStream_t)stream);
}

extern "C" int tdm_unet_backward(const void* wpack, const float* x, const int64_t* t, const float* noise,
                                 const float* eps, float* flat_grad, float* loss_out, void* workspace,
                                 int64_t workspace_bytes, int64_t batch, void* stream) {
    return unet_backward_impl(reinterpret_cast<const uint8_t*>(wpack), x, t, noise, eps, flat_grad,
                              loss_out, reinterpret_cast<uint8_t*>(workspace), workspace_bytes, batch,
                              (cudaStream_t)stream);
}

extern "C" int tdm_adamw_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                              int64_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, float grad_scale, const int64_t* step_dev,
                              void* stream) {
    TDM_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step_dev && n >= 0, "tdm_adamw_flat: bad arguments");
    if (n == 0) return TDM_OK;
    launch_pdl(adamw_kernel<false>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
               params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_dev,
               PeerBases{}, 1, 0, (int64_t)0, (int64_t)0, 0LL);
    TDM_CHECK_LAUNCH("tdm_adamw_flat");
    return TDM_OK;
}

extern "C" int64_t tdm_peer_grad_offset(int64_t n, int parity);

// How long a rank spins for its peers' "gradient k complete" flags before it traps.  Ranks may legitimately be
// far apart at step 1 (dataset decode, graph capture), so the default is minutes, not seconds;
// TDM_PEER_TIMEOUT_S overrides it.  Counted in SM clocks at a nominal 2 GHz.
static long long peer_timeout_cycles() {
    static const long long cycles = [] {
        double secs = 300.0;
        if (const char* e = std::getenv("TDM_PEER_TIMEOUT_S")) {
            const double v = std::atof(e);
            if (v > 0.0) secs = v;
        }
        return (long long)(secs * 2.0e9);
    }();
    return cycles;
}

extern "C" int tdm_adamw_flat_peer(float* params, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                                   float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                                   const int64_t* step_dev, const void* const* host_peer_bases, int world,
                                   int rank, void* stream) {
    TDM_CHECK_ARG(params && exp_avg && exp_avg_sq && step_dev && host_peer_bases && n > 0,
                  "tdm_adamw_flat_peer: bad arguments");
    TDM_CHECK_ARG(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world,
                  "tdm_adamw_flat_peer: world %d / rank %d out of range (max %d ranks)", world, rank, kMaxPeers);
    PeerBases pb{};
    for (int r = 0; r < world; ++r) {
        TDM_CHECK_ARG(host_peer_bases[r], "tdm_adamw_flat_peer: null peer buffer %d", r);
        pb.base[r] = reinterpret_cast<const uint8_t*>(host_peer_bases[r]);
    }
    const int64_t off0 = tdm_peer_grad_offset(n, 0);
    const int64_t slot = (tdm_peer_grad_offset(n, 1) - off0) / 4;
    launch_pdl(adamw_kernel<true>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
               params, (const float*)nullptr, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, grad_scale,
               step_dev, pb, world, rank, off0, slot, peer_timeout_cycles());
    TDM_CHECK_LAUNCH("tdm_adamw_flat_peer");
    return TDM_OK;
}
