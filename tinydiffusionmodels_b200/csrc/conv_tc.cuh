// conv_tc.cuh — the persistent warp-specialised tcgen05 implicit-GEMM convolution shared by the
// UNet forward (unet_fwd.cu) and the data-gradient pass of the backward (unet_bwd.cu).
//
//   warp 0        producer: weights once, then one input tile (all channel planes, halo included)
//                 per iteration via 1-D bulk async copies into a ring of smem stages
//   warp 1        MMA issuer: the whole warp runs the issue loop convergently, one elected lane issues the
//                 tcgen05.mma of a tile (M = 128 positions, K = 16 channels); the A operand of every tap is
//                 the same smem tile at a shifted row offset
//   warps 2..17   kEpiGroups = 4 groups of 4 epilogue warps; group g owns TMEM accumulator stage g, so four
//                 tiles are in flight; TMEM lane quarter = warp % 4, one thread per output position
//   warps 18..    (PROD > 0 only) gather warps that BUILD input planes in the smem stage instead of copying them:
//                 the nearest-x2 upsample of the 14x14 rb3 output for rb4.conv1 (cp.async; the upsampled tensor
//                 never reaches HBM), or the im2col of the single-channel image for rb1.conv1
//
// Three MMA schedules (template KXC), chosen per layer from measured sweeps (unet_layout.cuh).  With both
// operands in shared memory one MMA costs max(N/2, (4096 + 32 N)/128) cycles (tools/micro/mma_rate.cu), i.e. a
// small-N tap is bound by re-reading the A tile, and sharing that read between taps pays until the epilogue
// that undoes the sharing costs more:
//   0  nine taps, N = COUT each:              D[p]     = sum_tap W_tap X[p + off(tap)]
//   1  kx-triple: the three kx taps of a kernel row share one MMA with N = 3*COUT (weights of kx = 0,1,2 side
//      by side), A shifted by (ky-1)*Wp only:  Y[q][kx] = sum_ky W[ky][kx] X[q + (ky-1)Wp]
//      and the epilogue finishes               out[p]   = Y[p-1][0] + Y[p][1] + Y[p+1][2]
//      with two warp shuffles per channel (neighbour rows are neighbour lanes; the rows across a warp boundary
//      travel through smem).  Tiles overlap by two rows: tile t outputs rows [126t, 126t+126).
//   2  kx-pair: kx = 0,1 share an N = 2*COUT MMA, kx = 2 is accumulated into the kx = 1 columns with A one row
//      further:                                out[p]   = Z0[p-1] + Z1[p]       (one shuffle; tiles of 127 rows)
//
// Epilogues (template EPI): bias/ReLU/time-embedding, residual or 1x1-skip variants, the 2x2 upsample scatter
// (training), the final 1x1 out conv fused with the DDPM reverse step and in-kernel Philox noise, and for the
// backward the plain data gradient or the data gradient already masked and reduced for the ReLU it enters.
#pragma once
#include <cstring>
#include "common.cuh"
#include "diffusion_math.cuh"
#include "tc05.cuh"
#include "unet_layout.cuh"

namespace tdm {

// EPI_PLAIN: out = acc (+ residual if given) — no bias, no ReLU: the data-gradient convolutions.
// EPI_PLAIN_MASK: the data gradient of a block's conv2, finished for conv1 in the same pass:
//     out = acc (.) relu_mask1      red_plain[c] += sum acc[c]   (time_emb.bias gradient)
//     red_ts[c] += sum acc[c]*t/1000 (time_emb.weight gradient)  red_masked[c] += sum out[c] (conv1.bias gradient)
// (src/mnist.py:74-76 backward).  The unmasked gradient never reaches HBM.
enum : int { EPI_CONV1 = 0, EPI_RES = 1, EPI_RES_X = 2, EPI_RES_UP = 3, EPI_FINAL = 4, EPI_PLAIN = 5, EPI_PLAIN_MASK = 6 };

// Per-channel epilogue parameters BY VALUE (CPAR = true).  Kernel arguments live in the constant bank, so a
// compile-time-indexed a.cp.bias[ch] is a c[0x0][..] operand of the FADD itself: no load instruction and,
// what matters, no shared-memory wavefront.  ncu (capture F of round 1, a scratch capture - DESIGN.md section 6): the
// broadcast LDS.128 of these vectors cost ~8 wavefronts each (4 "ideal" + bank conflicts with the MMA's own
// operand reads) and kept l1tex__data_pipe_lsu_wavefronts at 80-90 % in rb1.conv2 / rb2.conv1 / rb4.conv1.
// The values must be known on the host at launch: the sampling engines register a host mirror of the flat
// parameters (tdm_unet_pack_weights_host); training updates parameters on the device and keeps CPAR = false.
struct ChanPar {
    float bias[64];
    float tw[64];
    float tb[64];
    float sbias[64];
    float aux[72];   // as s_aux: [0,32) aux_w, [32,64) aux_b (EPI_RES_X) / [32] out bias (EPI_FINAL)
};

struct ConvArgs {
    const uint8_t* in;     // input planes: row -GUARD of plane 0
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
    const uint8_t* res;    // residual planes (EPI_RES / EPI_RES_UP / EPI_FINAL / EPI_PLAIN)
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
    int np;                // positions covered by the buffers (multiple of 128)
    int batch;
    // training forward: ReLU masks, one uint32 per 32 channels per position: mask[chunk*mask_stride+pos]
    // (written by the forward epilogues, read by EPI_PLAIN_MASK)
    uint32_t* mask;
    int64_t mask_stride;
    float* red_plain;      // EPI_PLAIN_MASK: per-channel reductions accumulated with atomics
    float* red_ts;
    float* red_masked;
    // upsample gather (rb4.conv1, sampling): half-resolution source of input planes 0..7 (14x14 geometry, row -GUARD of plane 0)
    const uint8_t* in2;
    int64_t in2_ps;
    // CIN2 > 0: a second input (same geometry, CIN2 channels) whose 1x1 convolution accumulates into the same output:
    // the skip path's data gradient rides in the kernel of conv1's (weights: right after the 3x3 image)
    const uint8_t* in3;
    int64_t in3_ps;
    ChanPar cp;            // CPAR = true only
};

constexpr int kEpiGroups = 4;  // epilogue warp groups == TMEM accumulator stages (tiles in flight)
constexpr int kHalves = 1;     // warps per TMEM lane quarter inside a group (2: alternate 16-channel chunks)

// When the MMA thread takes the waits for tile i+1: 0 = at the top of tile i+1, 1 = blocking in the middle of tile
// i's MMAs, 2 = probed (test_wait) in the middle of tile i, blocking at the top of tile i+1 only if the probe failed.
#ifndef TDM_PREWAIT
#define TDM_PREWAIT 2
#endif
#ifndef TDM_MAX_STAGES
#define TDM_MAX_STAGES 4
#endif
constexpr int kBarBytes = 512;               // smem reserved for the mbarriers + the TMEM base slot
constexpr int kMaxStages = TDM_MAX_STAGES;   // input ring depth cap (bytes in flight per SM = stages x tile bytes)
constexpr int kGatherPlanes = 8;   // planes built by the gather producers when PROD > 0 (the 64 h3 channels)
constexpr int kSoloSmem = 115 * 1024;   // > (228 KB - 2 x 1 KB reserved) / 2
#ifndef TDM_GATHER_WARPS
#define TDM_GATHER_WARPS 2
#endif
#ifndef TDM_IM2COL_UNROLL
#define TDM_IM2COL_UNROLL 4
#endif
#ifndef TDM_IM2COL_WARPS
#define TDM_IM2COL_WARPS 8
#endif
constexpr int kIm2colWarps = TDM_IM2COL_WARPS;   // PROD used by rb1.conv1
constexpr int kIm2colUnroll = TDM_IM2COL_UNROLL; // tile rows per lane whose 9 loads are in flight together
#ifndef TDM_GATHER_MODE
#define TDM_GATHER_MODE 2   // 0 = LDG.128 -> registers -> STS.128; 1 = cp.async.cg; 2 = cp.async.ca
#endif
constexpr int kGatherWarps = TDM_GATHER_WARPS;   // PROD used by rb4.conv1 on the sampling path

template <int W, int CIN, int COUT, bool SKIPG, int TAPS, int KXC, int PROD = 0, int CIN2 = 0, int NGRP = kEpiGroups>
struct ConvCfg {
    using G = Geo<W>;
    static constexpr int NPL = CIN / 8;
    static constexpr int NPL_ALL = (CIN + CIN2) / 8;   // planes of a stage: the 3x3 input, then the extra 1x1 input
    static constexpr int STAGE_BYTES = NPL_ALL * G::RT * 16;
    static constexpr int WCONV_BYTES = TAPS * CIN * COUT * 2;
    static constexpr int W_BYTES = WCONV_BYTES + (SKIPG ? CIN * COUT * 2 : 0) + CIN2 * COUT * 2;
    static constexpr int PARAM_BYTES = 5 * 96 * 4;
    static constexpr int XCH_BYTES = (KXC == 1 ? NGRP * kHalves * 4 * 2 * COUT * 4 : KXC == 2 ? NGRP * kHalves * 4 * COUT * 4 : 0) + NGRP * 2 * 128 * 4;
    static constexpr int MAX_SMEM = 227 * 1024;
    static constexpr int AVAIL = MAX_SMEM - W_BYTES - PARAM_BYTES - XCH_BYTES - kBarBytes;
    // a gather warp may run at most one ring phase ahead of the MMA warp (mbarrier parity), so the ring is
    // at least as deep as there are gather warps
    static constexpr int STAGE_CAP = PROD > kMaxStages ? PROD : kMaxStages;
    static constexpr int NSTAGE = (AVAIL / STAGE_BYTES) > STAGE_CAP ? STAGE_CAP : (AVAIL / STAGE_BYTES);
    static_assert(PROD <= NSTAGE, "more gather warps than input stages");
    static_assert(NSTAGE >= 2, "need at least two input stages");
    static constexpr int NACC = NGRP;   // epilogue groups == TMEM accumulator stages == tiles in flight
    static constexpr int NMAIN = KXC == 1 ? 3 * COUT : KXC == 2 ? 2 * COUT : COUT;   // columns of the conv accumulator
    static constexpr int ACC_COLS = NMAIN + (SKIPG ? COUT : 0);
    static constexpr int TMEM_COLS = (NACC * ACC_COLS <= 32) ? 32 : (NACC * ACC_COLS <= 64) ? 64
                                   : (NACC * ACC_COLS <= 128) ? 128 : (NACC * ACC_COLS <= 256) ? 256 : 512;
    static_assert(NACC * ACC_COLS <= 512, "accumulators exceed TMEM");
    static_assert(NMAIN <= 256 && NMAIN % 16 == 0, "UMMA N");
    static constexpr int SMEM_USED = W_BYTES + NSTAGE * STAGE_BYTES + PARAM_BYTES + XCH_BYTES + kBarBytes;
    // Requested size: more than half an SM's shared memory, so that two of these persistent CTAs can never share an
    // SM.  Under programmatic dependent launch the next kernel's CTAs are placed as SMs free up one by one; a small
    // kernel (rb1.conv2: 74 KB, 55 registers) would land twice on the first free SMs and leave others empty, and
    // with static tile striding the doubled-up SMs then set the kernel's time (measured: +9 % per reverse step).
    static constexpr int SMEM_BYTES = SMEM_USED > kSoloSmem ? SMEM_USED : kSoloSmem;
    static_assert((1 + 2 * NSTAGE + 2 * NGRP) * 8 + 8 <= kBarBytes && NGRP >= 1 && 64 + 128 * kHalves * NGRP + 32 * (PROD > 0 ? PROD : 0) <= 1024, "mbarrier block overflows its smem slot");
    // warp 0 producer, warp 1 MMA issuer, NACC groups of 4 epilogue warps, then the gather warps
    static constexpr int THREADS = 64 + 128 * kHalves * NACC + 32 * PROD;
    // gather kind: 0 none, 1 nearest-x2 upsample of the 64 h3 channels (rb4.conv1), 2 im2col of the
    // single-channel image (rb1.conv1 as a 1x1 convolution over 32 "channels" = 27 hi/lo tap terms)
    static constexpr int GK = PROD == 0 ? 0 : (CIN == 96 ? 1 : 2);
    static constexpr int BULK_PLANES = GK == 0 ? NPL_ALL : GK == 1 ? NPL - kGatherPlanes : 0;
    static_assert(CIN2 == 0 || (GK == 0 && KXC == 0 && !SKIPG && CIN2 % 16 == 0 && NPL_ALL <= 32), "extra 1x1 input: plain nine-tap kernels only");
    // arrivals on a stage's full barrier: the bulk issuer, plus the gather warp owning the tile
    // (one per lane when it copies with cp.async)
    static constexpr int FULL_ARRIVALS = (BULK_PLANES > 0 ? 1 : 0) + (GK == 0 ? 0 : (GK == 1 && TDM_GATHER_MODE != 0) ? 32 : 1);
    // tile t: accumulator rows [t*TSTRIDE - ROW0, +128)
    static constexpr int TSTRIDE = KXC == 1 ? 126 : KXC == 2 ? 127 : 128;
    static constexpr int ROW0 = KXC ? 1 : 0;
    static constexpr int ROW1 = KXC == 1 ? 127 : 128;   // output rows are tile rows [ROW0, ROW1)
    static constexpr int CW = 16;              // channels per epilogue chunk
};

// Development aid (TDM_NVCC_DEFS=-DTDM_TIMELINE=<EPI>): CTA 0 of the kernels whose EPI matches records clock64() at
// the pipeline events of its first 96 tiles into g_timeline[tile][event]; tools/timeline_probe.py prints the deltas.
#ifdef TDM_TIMELINE
__device__ long long g_timeline[96 * 16];
#define TDM_TL(EPI_, it_, ev_)                                                                  \
    do {                                                                                        \
        if ((EPI_) == TDM_TIMELINE && blockIdx.x == 0 && (it_) < 96 && (threadIdx.x & 31) == 0) \
            g_timeline[(it_) * 16 + (ev_)] = clock64();                                         \
    } while (0)
#else
#define TDM_TL(EPI_, it_, ev_) do {} while (0)
#endif

template <int N>
__device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t (&r)[N]) {
    if constexpr (N == 16) tmem_ld16(taddr, r);
    else tmem_ld32(taddr, r);
}

// Sum x[0..15] over the 32 lanes of a warp so that lane l (and l ^ 16) ends up with the total of x[l & 15]:
// recursive halving (8 + 4 + 2 + 1 exchanges) plus one exchange across the half-warps - 16 shuffles where sixteen
// independent butterfly reductions would take 80.
__device__ __forceinline__ float lane_transpose_sum16(const float (&x)[16], int lane) {
    float a8[8], a4[4], a2[2];
    const bool b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = b3 ? x[8 + i] : x[i], send = b3 ? x[i] : x[8 + i];
        a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b2 ? a8[4 + i] : a8[i], send = b2 ? a8[i] : a8[4 + i];
        a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b1 ? a4[2 + i] : a4[i], send = b1 ? a4[i] : a4[2 + i];
        a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    const float keep = b0 ? a2[1] : a2[0], send = b0 ? a2[0] : a2[1];
    float v = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;   // channel index = 8*b3 + 4*b2 + 2*b1 + b0 = lane & 15
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

template <int W, int CIN, int COUT, int EPI, bool SKIPG, int TAPS = 9, int KXC = 0, int PROD = 0, bool CPAR = false, int CIN2 = 0, int NGRP = kEpiGroups>
__global__ void __launch_bounds__(64 + 128 * kHalves * NGRP + 32 * PROD, 1) conv3x3_tc_kernel(const __grid_constant__ ConvArgs a) {
    constexpr bool kPlain = (EPI == EPI_PLAIN || EPI == EPI_PLAIN_MASK);
    static_assert(!CPAR || (COUT <= 64 && !kPlain), "by-value channel parameters: forward epilogues, <= 64 channels");
    using C = ConvCfg<W, CIN, COUT, SKIPG, TAPS, KXC, PROD, CIN2, NGRP>;
    static_assert(C::GK != 1 || (W == 28 && CIN == 96), "upsample gather serves rb4.conv1's concat input");
    static_assert(C::GK != 2 || (W == 28 && CIN == 32 && COUT == 32 && TAPS == 1 && KXC == 0), "im2col gather serves rb1.conv1");
    using G = Geo<W>;
    static_assert(COUT == 32 || COUT == 64 || COUT == 96, "COUT");
    static_assert(TAPS == 9 || TAPS == 1, "3x3 or 1x1");
    static_assert(KXC == 0 || TAPS == 9, "kx-combining is a 3x3 schedule");
    static_assert(KXC >= 0 && KXC <= 2, "KXC: 0 nine taps, 1 kx-triple, 2 kx-pair");
    static_assert(!SKIPG || COUT <= 64, "skip GEMM variant is forward-only");
    static_assert(EPI != EPI_FINAL || COUT == 32, "final epilogue expects 32 channels");
    constexpr int CW = C::CW;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w = smem;
    uint8_t* s_in = smem + C::W_BYTES;
    float* s_par = reinterpret_cast<float*>(s_in + C::NSTAGE * C::STAGE_BYTES);
    float* s_bias = s_par;
    float* s_tw = s_par + 96;
    float* s_tb = s_par + 192;
    float* s_sbias = s_par + 288;
    float* s_aux = s_par + 384;
    float* s_dot = s_par + 480;                       // [grp][tile parity][128] partial out-conv dots
    float* s_xch = s_dot + C::NACC * 2 * 128;         // kx-combine boundary rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_dot) + C::XCH_BYTES);
    uint64_t* bar_w = bars;
    uint64_t* bar_full = bars + 1;
    uint64_t* bar_empty = bar_full + C::NSTAGE;
    uint64_t* bar_accf = bar_empty + C::NSTAGE;
    uint64_t* bar_acce = bar_accf + C::NACC;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acce + C::NACC);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nt = (a.np + C::TSTRIDE - 1) / C::TSTRIDE;

    // ---- setup (everything before pdl_wait touches no global memory: it may overlap the previous kernel) ----
    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < C::NSTAGE; ++i) {
            mbar_init(bar_full + i, C::FULL_ARRIVALS);
            mbar_init(bar_empty + i, 1);
        }
        for (int i = 0; i < C::NACC; ++i) {
            mbar_init(bar_accf + i, 1);
            mbar_init(bar_acce + i, 4 * kHalves);  // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<C::TMEM_COLS>(s_tmem);
    pdl_wait();                 // the previous kernel's results are visible from here on
    pdl_launch_dependents();    // the next kernel may start its own setup as our CTAs retire
    if (!CPAR && threadIdx.x < COUT) {
        const int c = threadIdx.x;
        s_bias[c] = kPlain ? 0.f : a.bias[c];
        s_tw[c] = (EPI == EPI_CONV1) ? a.tw[c] : 0.f;
        s_tb[c] = (EPI == EPI_CONV1) ? a.tb[c] : 0.f;
        s_sbias[c] = SKIPG ? a.sbias[c] : 0.f;
        s_aux[c] = (EPI == EPI_RES_X || EPI == EPI_FINAL) ? a.aux_w[c] : 0.f;
        if (EPI == EPI_RES_X) s_aux[32 + c] = a.aux_b[c];
        if (EPI == EPI_FINAL && c == 0) s_aux[32] = a.aux_b[0];
    }
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
        for (int tile = blockIdx.x; C::BULK_PLANES > 0 && tile < nt; tile += gridDim.x, ++it) {
            const int s = it % C::NSTAGE;
            const uint32_t ph = (it / C::NSTAGE) & 1;
            if (lane == 0) {
                mbar_wait(bar_empty + s, ph ^ 1);
                TDM_TL(EPI, it, 0);
                mbar_arrive_expect_tx(bar_full + s, C::BULK_PLANES * G::RT * 16);
            }
            __syncwarp();
            if (lane < C::BULK_PLANES) {
                // smem row 0 = global row tile*TSTRIDE - ROW0 - HALO; the buffers start at row -GUARD.
                // With gather producers the bulk planes (a.in = their first plane) sit after the gathered ones;
                // with an extra 1x1 input its planes (a.in3) follow the 3x3 input's.
                const int64_t row = (int64_t)tile * C::TSTRIDE - C::ROW0 - G::HALO + G::GUARD;
                const uint8_t* src = (CIN2 > 0 && lane >= C::NPL) ? a.in3 + (lane - C::NPL) * a.in3_ps : a.in + lane * a.in_ps;
                bulk_g2s(s_in + s * C::STAGE_BYTES + (lane + (C::NPL_ALL - C::BULK_PLANES)) * (G::RT * 16), src + row * 16,
                         G::RT * 16, bar_full + s);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected lane of this warp runs the persistent issue loop =====
        constexpr uint32_t idesc = make_idesc_bf16(128, C::NMAIN);
        constexpr uint32_t idesc_skip = make_idesc_bf16(128, COUT);
        mbar_wait(bar_w, 0);
        constexpr uint32_t B_LBO = (KXC ? 3 * COUT : COUT) * 16;
        // The MMAs of a tile are issued inside `if (elect_one())`: ptxas then keeps the descriptors in UNIFORM registers
        // and every "base + constant" is one UIADD3.64 next to its UTCHMMA (2 SASS instructions per MMA).  Issued with
        // the election inside each asm statement (or under `lane == 0`) every UTCHMMA dragged ~10 ELECT / R2UR / VOTEU
        // instructions along to move its operands out of vector registers, and the issuing warp - not the tensor pipe -
        // set the pace of the N = 32 layers (58 cycles per 40-cycle MMA; ncu source page, tools/micro).
        const uint32_t w_addr = smem_u32(s_w);
        const uint32_t in_addr0 = smem_u32(s_in);
        const uint32_t tmem_u = tmem_base;
        // ONE elected thread runs the whole persistent issue loop.  The barriers of tile i+1 (accumulator free, input
        // full) are probed in the middle of tile i's MMAs, while the pipe is busy with what is already queued (a
        // thread can run ~7 MMAs ahead); when the probe succeeds nothing but the two commits sits between the last MMA
        // of a tile and the first MMA of the next, when it fails the thread finishes tile i first and blocks at the
        // top of tile i+1 (TDM_PREWAIT).  Measured with the in-kernel timeline (tools/timeline_probe.py): with a per-tile
        // "all lanes wait, elect, issue, __syncwarp" loop 475 of rb1.conv2's 1,160 cycles per tile passed between
        // the last commit and the next first MMA - two already-satisfied mbarrier waits cost ~95 cycles each - and
        // the pipe drained every tile.
        if (elect_one()) {
            auto wait_tile = [&](int itw) {
                mbar_wait(bar_acce + itw % C::NACC, ((itw / C::NACC) & 1) ^ 1);
                mbar_wait(bar_full + itw % C::NSTAGE, (itw / C::NSTAGE) & 1);
            };
            auto prewait = [&](int itw, bool more) -> bool {
#if TDM_PREWAIT == 1
                if (more) wait_tile(itw);
                return true;
#elif TDM_PREWAIT == 2
                return more && mbar_test(bar_acce + itw % C::NACC, ((itw / C::NACC) & 1) ^ 1) &&
                       mbar_test(bar_full + itw % C::NSTAGE, (itw / C::NSTAGE) & 1);
#else
                return false;
#endif
            };
            int it = 0;
            bool ready = false;
            for (int tile = blockIdx.x; tile < nt; tile += gridDim.x, ++it) {
                const int s = it % C::NSTAGE;
                const int acc = it % C::NACC;
                const bool more = tile + (int)gridDim.x < nt;
                if (!ready) wait_tile(it);
                TDM_TL(EPI, it, 3);
                if constexpr (C::GK == 1 && TDM_GATHER_MODE != 0) fence_proxy_async_smem();   // cp.async (generic proxy) data -> async-proxy MMA reads
                tc_fence_after_sync();
                const uint64_t in_base = make_smem_desc(in_addr0 + (uint32_t)s * C::STAGE_BYTES, G::RT * 16, 128);
                // rebuilt per tile on purpose: kept across the loop the 18+ weight descriptors live in vector registers
                // and cost two R2UR each per tile; rebuilt from the uniform address they are one UIADD3.64 apiece
                uint32_t w_addr_t = w_addr;
                asm volatile("" : "+r"(w_addr_t));   // opaque: keeps the compiler from hoisting the descriptors out of the loop
                const uint64_t w_base = make_smem_desc(w_addr_t, B_LBO, 128);
                const uint64_t ws_base = make_smem_desc(w_addr_t + C::WCONV_BYTES, COUT * 16, 128);
                const uint32_t d = tmem_u + acc * C::ACC_COLS;
                if constexpr (KXC == 1) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        if (ky == 2) ready = prewait(it + 1, more);
#pragma unroll
                        for (int ks = 0; ks < CIN / 16; ++ks) {
                            umma_bf16(d, desc_add(in_base, (2 * ks) * (G::RT * 16) + (G::HALO + (ky - 1) * G::Wp) * 16),
                                      desc_add(w_base, ((ky * C::NPL + 2 * ks) * 3 * COUT) * 16), idesc, (ky | ks) != 0);
                        }
                    }
                } else if constexpr (KXC == 2) {
                    // kx = 0,1 share one N = 2*COUT MMA (columns [Z0 | Z1], A at the centre column); kx = 2 is a plain
                    // tap accumulated into Z1 with A one row further.  Same weight image as the triple schedule.
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        if (ky == 2) ready = prewait(it + 1, more);
#pragma unroll
                        for (int ks = 0; ks < CIN / 16; ++ks) {
                            umma_bf16(d, desc_add(in_base, (2 * ks) * (G::RT * 16) + (G::HALO + (ky - 1) * G::Wp) * 16),
                                      desc_add(w_base, ((ky * C::NPL + 2 * ks) * 3 * COUT) * 16), idesc, (ky | ks) != 0);
                            umma_bf16(d + COUT, desc_add(in_base, (2 * ks) * (G::RT * 16) + (G::HALO + (ky - 1) * G::Wp + 1) * 16),
                                      desc_add(w_base, ((ky * C::NPL + 2 * ks) * 3 * COUT + 2 * COUT) * 16), idesc_skip, 1u);
                        }
                    }
                } else {
#pragma unroll
                    for (int tap = 0; tap < TAPS; ++tap) {
#ifdef TDM_DBG_TAPS
                        if (tap >= TDM_DBG_TAPS) break;   // timing experiment only (wrong results)
#endif
                        if (TAPS == 9 && tap == 6) ready = prewait(it + 1, more);
                        const int off = (TAPS == 1) ? 0 : (tap / 3 - 1) * G::Wp + (tap % 3 - 1);
#pragma unroll
                        for (int ks = 0; ks < CIN / 16; ++ks) {
                            umma_bf16(d, desc_add(in_base, (2 * ks) * (G::RT * 16) + (G::HALO + off) * 16),
                                      desc_add(w_base, ((tap * C::NPL + 2 * ks) * COUT) * 16), idesc, (tap | ks) != 0);
                        }
                        if (tap == 0) TDM_TL(EPI, it, 9);
                        if (tap == 2) TDM_TL(EPI, it, 10);
                        if (tap == 5) TDM_TL(EPI, it, 11);
                        if (tap == 8) TDM_TL(EPI, it, 12);
                    }
                    if constexpr (CIN2 > 0) {
                        // the extra input's 1x1 convolution lands in the same accumulator (centre tap, its own weights)
#pragma unroll
                        for (int ks = 0; ks < CIN2 / 16; ++ks) {
                            umma_bf16(d, desc_add(in_base, (C::NPL + 2 * ks) * (G::RT * 16) + G::HALO * 16),
                                      desc_add(ws_base, ((2 * ks) * COUT) * 16), idesc, 1u);
                        }
                    }
                }
                if constexpr (SKIPG) {
#pragma unroll
                    for (int ks = 0; ks < CIN / 16; ++ks) {
                        umma_bf16(d + C::NMAIN, desc_add(in_base, (2 * ks) * (G::RT * 16) + G::HALO * 16),
                                  desc_add(ws_base, ((2 * ks) * COUT) * 16), idesc_skip, ks != 0);
                    }
                }
                umma_commit(bar_empty + s);   // smem stage reusable once these MMAs retire
                umma_commit(bar_accf + acc);  // accumulator complete
                TDM_TL(EPI, it, 13);
                if (TAPS == 1) ready = false;   // two MMAs per tile: nothing to hide the waits behind
            }
        }
        __syncwarp();
    } else if (PROD != 0 && warp >= 2 + 4 * kHalves * C::NACC) {
        // ===== gather producers (PROD warps): planes 0..7 of the tile = nearest-x2 upsample of the 14x14 source.
        //       Warp w owns tiles it = w (mod PROD) entirely; lane = smem row.  Measured on B200 @16384 (kernel us):
        //         TDM_GATHER_MODE 2  cp.async.ca (through L1), 2 warps   936   <- default
        //                         1  cp.async.cg (L1 bypass),  2 warps  1036   (ncu: 34 smem wavefronts per LDGSTS,
        //                                                                       one per lane; the x2/y2 duplicates
        //                                                                       also go back to L2)
        //                         0  LDG.128 -> regs -> STS.128, 2 warps 1263   (latency-bound: 16 loads in flight)
        //         gather skipped (upper bound of what is left to win)    761
        //       Warp count: registers are per SM sub-partition, so <= 20 warps (5 per SMSP) keep the 96-register
        //       cap the epilogue warps need; 21+ warps drop it to 80 and cost ~90 us.  cp.async needs no staging
        //       registers and no waiting: completion reaches the full barrier through
        //       cp.async.mbarrier.arrive.noinc (one arrival per lane), so a warp runs ahead as far as the ring
        //       has free stages. =====
        const int pw = warp - (2 + 4 * kHalves * C::NACC);
        if constexpr (C::GK == 2) {
            // ===== im2col of the single-channel image (src/mnist.py:74 conv1 of rb1): tile row p gets the 3x3
            //       window of x around p as 32 "channels", so the 1 -> 32 convolution is one K = 32 GEMM:
            //         k  0.. 8  hi(x_tap)   x  hi(w_tap)        hi(v) = bf16(v), lo(v) = bf16(v - hi(v))
            //         k  9..17  lo(x_tap)   x  hi(w_tap)
            //         k 18..26  hi(x_tap)   x  lo(w_tap)        (k 27..31 zero)
            //       i.e. the product keeps ~16 mantissa bits of both factors with fp32 accumulation - the
            //       CUDA-core fp32 version of this layer was issue-bound at 0.41 ms (B = 16384).
            //       Only tile rows [0,128) are built: a 1x1 convolution reads no halo. =====
            for (int it = pw, tile = blockIdx.x + pw * gridDim.x; tile < nt; tile += PROD * gridDim.x, it += PROD) {
                const int s = it % C::NSTAGE;
                const uint32_t ph = (it / C::NSTAGE) & 1;
                mbar_wait(bar_empty + s, ph ^ 1);
                uint8_t* st = s_in + s * C::STAGE_BYTES + G::HALO * 16;   // tile row 0 of plane 0
#pragma unroll kIm2colUnroll
                for (int r = lane; r < 128; r += 32) {
                    const int pos = tile * 128 + r;
                    const int b = (int)((uint32_t)pos / (uint32_t)G::S);
                    const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
                    const int rw = rem / G::Wp, c = rem - rw * G::Wp;
                    const bool ok = b < a.batch && rw >= 1 && c < G::W;
                    const float* xb = a.x + (int64_t)b * 784 + (rw - 1) * 28 + c;   // centre pixel
                    float hi[9], lo[9];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx) {
                            const bool in = ok && (unsigned)(rw - 1 + ky - 1) < 28u && (unsigned)(c + kx - 1) < 28u;
                            const float v = in ? __ldg(xb + (ky - 1) * 28 + (kx - 1)) : 0.f;
                            const float h = __bfloat162float(__float2bfloat16_rn(v));
                            hi[ky * 3 + kx] = h;
                            lo[ky * 3 + kx] = v - h;
                        }
                    float e[32];
#pragma unroll
                    for (int k = 0; k < 9; ++k) { e[k] = hi[k]; e[9 + k] = lo[k]; e[18 + k] = hi[k]; }
#pragma unroll
                    for (int k = 27; k < 32; ++k) e[k] = 0.f;
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) {
                        uint4 o;
                        o.x = pack_bf16x2(e[8 * pl + 0], e[8 * pl + 1]);
                        o.y = pack_bf16x2(e[8 * pl + 2], e[8 * pl + 3]);
                        o.z = pack_bf16x2(e[8 * pl + 4], e[8 * pl + 5]);
                        o.w = pack_bf16x2(e[8 * pl + 6], e[8 * pl + 7]);
                        *reinterpret_cast<uint4*>(st + pl * (G::RT * 16) + r * 16) = o;
                    }
                }
                fence_proxy_async_smem();   // this lane's generic-proxy stores -> visible to the MMA's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_full + s);
            }
        } else {
        using GS = Geo<14>;
        static_assert(PROD == 0 || G::RT % 64 == 0, "two rows per lane per iteration");
        for (int it = pw, tile = blockIdx.x + pw * gridDim.x; tile < nt; tile += PROD * gridDim.x, it += PROD) {
            const int s = it % C::NSTAGE;
            const uint32_t ph = (it / C::NSTAGE) & 1;
            mbar_wait(bar_empty + s, ph ^ 1);
            uint8_t* st = s_in + s * C::STAGE_BYTES;
            const int pos0 = tile * C::TSTRIDE - C::ROW0 - G::HALO;   // may be negative for tile 0
#if TDM_GATHER_MODE != 0
#pragma unroll 1
            for (int r = lane; r < G::RT; r += 32) {   // smem row
                const int pos = pos0 + r;
                const uint8_t* src = a.in2;            // any valid address when the row is zero-filled
                uint32_t nbytes = 0;
                if (pos >= 0) {
                    const int b = (int)((uint32_t)pos / (uint32_t)G::S);
                    const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
                    const int rw = rem / G::Wp, c = rem - rw * G::Wp;
                    if (b < a.batch && rw >= 1 && c < G::W) {
                        const int64_t p14 = (int64_t)b * GS::S + ((rw - 1) / 2 + 1) * GS::Wp + c / 2;
                        src = a.in2 + (p14 + GS::GUARD) * 16;
                        nbytes = 16;
                    }
                }
#pragma unroll
                for (int j = 0; j < kGatherPlanes; ++j)
                    cp_async16<TDM_GATHER_MODE == 2>(st + j * (G::RT * 16) + r * 16, nbytes ? src + j * a.in2_ps : src, nbytes);
            }
            cp_async_arrive_noinc(bar_full + s);
#else
#pragma unroll 1
            for (int r = lane; r < G::RT; r += 64) {   // smem rows r and r + 32
                const uint8_t* src[2];
                bool ok[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int pos = pos0 + r + 32 * h;
                    src[h] = a.in2;
                    ok[h] = false;
                    if (pos >= 0) {
                        const int b = (int)((uint32_t)pos / (uint32_t)G::S);   // positions fit 32 bits: multiply-shift
                        const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
                        const int rw = rem / G::Wp, c = rem - rw * G::Wp;
                        if (b < a.batch && rw >= 1 && c < G::W) {
                            const int64_t p14 = (int64_t)b * GS::S + ((rw - 1) / 2 + 1) * GS::Wp + c / 2;
                            src[h] = a.in2 + (p14 + GS::GUARD) * 16;
                            ok[h] = true;
                        }
                    }
                }
                uint4 v[2][kGatherPlanes];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < kGatherPlanes; ++j)
                        v[h][j] = ok[h] ? __ldg(reinterpret_cast<const uint4*>(src[h] + j * a.in2_ps))
                                        : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int j = 0; j < kGatherPlanes; ++j)
                        *reinterpret_cast<uint4*>(st + j * (G::RT * 16) + (r + 32 * h) * 16) = v[h][j];
            }
            fence_proxy_async_smem();   // this lane's generic-proxy stores -> visible to the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + s);
#endif
        }
        }
    } else {
        // ===== epilogue: group g = (warp-2)/4 owns accumulator stage g (tiles it = g, g+NACC, ..),
        //       so the epilogues of consecutive tiles overlap; TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        const int grp = (warp - 2) / (4 * kHalves);
        const int half = ((warp - 2) >> 2) % kHalves;   // which alternate 16-channel chunks this warp owns
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + grp * C::ACC_COLS;
        // EPI_PLAIN_MASK: lane l (and l ^ 16) carries the running sums of channel 16*chunk + (l & 15)
        constexpr int kRed = (EPI == EPI_PLAIN_MASK) ? COUT / (kHalves * CW) : 1;
        float red_p[kRed], red_t[kRed], red_m[kRed];
#pragma unroll
        for (int i = 0; i < kRed; ++i) red_p[i] = red_t[i] = red_m[i] = 0.f;
        int n = 0;
        for (int tile = blockIdx.x + grp * gridDim.x; tile < nt; tile += C::NACC * gridDim.x, ++n) {
            const uint32_t aph = n & 1;
            // ---- phase A: everything that does not need the accumulator (overlaps the MMAs) ----
            const int trow = q * 32 + lane;                                   // row inside the tile
            const int64_t pos = (int64_t)tile * C::TSTRIDE - C::ROW0 + trow;  // global position
            const bool owned = trow >= C::ROW0 && trow < C::ROW1 && pos < a.np;  // this tile outputs pos
            const int b = (int)((uint32_t)pos / (uint32_t)G::S);   // positions fit 32 bits: division by a constant is a multiply-shift
            const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
            const int r = rem / G::Wp, c = rem - r * G::Wp;
            const bool valid = owned && pos >= 0 && b < a.batch && r >= 1 && c < G::W;
            const int y = r - 1;

            float ts = 0.f;
            if ((EPI == EPI_CONV1 || EPI == EPI_PLAIN_MASK) && valid) ts = (float)(int)__ldg(a.t + b) / 1000.0f;   // via int32 (I2F.S64 is a slow-path conversion)
            uint32_t mw[EPI == EPI_PLAIN_MASK ? (COUT + 31) / 32 : 1];
            if constexpr (EPI == EPI_PLAIN_MASK) {
#pragma unroll
                for (int i = 0; i < (COUT + 31) / 32; ++i) mw[i] = (valid && owned) ? __ldg(a.mask + i * a.mask_stride + pos) : 0u;
            }
            float xin = 0.f;
            if ((EPI == EPI_RES_X || (EPI == EPI_FINAL && half == 0)) && valid && a.x)
                xin = __ldg(a.x + (int64_t)b * 784 + y * 28 + c);
            constexpr bool kHasRes = (EPI == EPI_RES || EPI == EPI_RES_UP || EPI == EPI_FINAL || EPI == EPI_PLAIN);   // EPI_PLAIN_MASK: none
            // residual planes of this warp's chunks only: local index i -> plane (i/2)*4 + half*2 + i%2
            uint4 rv[kHasRes ? COUT / (8 * kHalves) : 1];
            if constexpr (kHasRes) {
#pragma unroll
                for (int i = 0; i < COUT / (8 * kHalves); ++i) {
                    const int pl = (kHalves == 2) ? (i >> 1) * 4 + half * 2 + (i & 1) : i;
                    rv[i] = make_uint4(0, 0, 0, 0);
                    if (valid && (EPI != EPI_PLAIN || a.res))  // residual planes share this geometry
                        rv[i] = *reinterpret_cast<const uint4*>(a.res + pl * a.res_ps + (pos + G::GUARD) * 16);
                }
            }
            StepCoef sc{};
            float zz = 0.f;
            bool add_noise = false;
            if constexpr (EPI == EPI_FINAL) {
                if (a.fuse_step && valid && half == 0) {   // half 0 finishes the pixel
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

            int up_p00[2] = {0, 0};
            bool up_ok[2] = {false, false};
            if constexpr (EPI == EPI_RES_UP) {
                const int my_p00 = b * Geo<28>::S + (2 * y + 1) * Geo<28>::Wp + 2 * c;   // top-left of the 2x2 block
#pragma unroll
                for (int hs = 0; hs < 2; ++hs) {
                    const int src = (lane >> 1) + 16 * hs;
                    up_p00[hs] = __shfl_sync(0xffffffffu, my_p00, src);
                    up_ok[hs] = __shfl_sync(0xffffffffu, (int)valid, src) != 0;
                }
            }

            // ---- phase B: drain the accumulator ----
            mbar_wait(bar_accf + grp, aph);
            if (q == 0) TDM_TL(EPI, n * C::NACC + grp, 5);
            tc_fence_after_sync();
            float dot = 0.f;
            uint32_t mbits = 0;

#pragma unroll
            for (int ci = 0; ci < COUT / (kHalves * CW); ++ci) {
                const int c0 = (kHalves * ci + half) * CW;
                const bool last_chunk = ci == COUT / (kHalves * CW) - 1;
                float acc[CW];
                uint32_t r2[CW];
                if constexpr (KXC == 2) {
                    // out[p] = Z0[p-1] + Z1[p]: one shuffle per channel; the row above lane 0 comes through smem
                    uint32_t d0[CW], d1[CW];
                    tmem_ld_n<CW>(taddr + c0, d0);
                    tmem_ld_n<CW>(taddr + COUT + c0, d1);
                    if constexpr (SKIPG) tmem_ld_n<CW>(taddr + C::NMAIN + c0, r2);
                    tmem_ld_wait();
                    if (last_chunk) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_acce + grp);
                        if (q == 0) TDM_TL(EPI, n * C::NACC + grp, 7);
                    }
                    float4* xs = reinterpret_cast<float4*>(s_xch + (grp * 4 + q) * COUT + c0);   // per (group, quarter)
                    if (lane == 31) {
#pragma unroll
                        for (int k = 0; k < CW / 4; ++k)
                            xs[k] = make_float4(__uint_as_float(d0[4 * k]), __uint_as_float(d0[4 * k + 1]),
                                                __uint_as_float(d0[4 * k + 2]), __uint_as_float(d0[4 * k + 3]));
                    }
                    named_bar_sync(1 + grp * kHalves + half, 128);
                    // q == 0: tile row 0 is never an output row, any finite value will do
                    const float4* xprev = reinterpret_cast<const float4*>(s_xch + (grp * 4 + (q > 0 ? q - 1 : 0)) * COUT + c0);
                    const bool first = lane == 0;
#pragma unroll
                    for (int k4 = 0; k4 < CW / 4; ++k4) {
                        const float4 pu = xprev[k4];   // same address for all lanes: broadcast
                        const float pus[4] = {pu.x, pu.y, pu.z, pu.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = 4 * k4 + j;
                            const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[k]), 1);
                            acc[k] = (first ? pus[j] : up) + __uint_as_float(d1[k]);
                        }
                    }
                } else if constexpr (KXC == 1) {
                    uint32_t d0[CW], d1[CW], d2[CW];
                    tmem_ld_n<CW>(taddr + c0, d0);
                    tmem_ld_n<CW>(taddr + COUT + c0, d1);
                    tmem_ld_n<CW>(taddr + 2 * COUT + c0, d2);
                    if constexpr (SKIPG) tmem_ld_n<CW>(taddr + 3 * COUT + c0, r2);
                    tmem_ld_wait();
                    if (last_chunk) {
                        // all TMEM reads of this accumulator are done: hand it back to the MMA warp
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_acce + grp);
                        if (q == 0) TDM_TL(EPI, n * C::NACC + grp, 7);
                    }
                    // rows p-1 / p+1 are lanes -1 / +1; across the warp boundary they come through smem.
                    // Everything below is branch-free per channel (selects, broadcast loads): per-channel
                    // `if (lane == 0)` patches compile to divergent branches and dominated the kernel.
                    float4* xs = reinterpret_cast<float4*>(s_xch + ((grp * 4 + q) * 2) * COUT + c0);   // per (group, quarter)
                    if (lane == 31) {
#pragma unroll
                        for (int k = 0; k < CW / 4; ++k)
                            xs[k] = make_float4(__uint_as_float(d0[4 * k]), __uint_as_float(d0[4 * k + 1]),
                                                __uint_as_float(d0[4 * k + 2]), __uint_as_float(d0[4 * k + 3]));
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < CW / 4; ++k)
                            xs[COUT / 4 + k] = make_float4(__uint_as_float(d2[4 * k]), __uint_as_float(d2[4 * k + 1]),
                                                           __uint_as_float(d2[4 * k + 2]), __uint_as_float(d2[4 * k + 3]));
                    }
                    named_bar_sync(1 + grp * kHalves + half, 128);   // the 4 warps (quarters) sharing this chunk
                    // q == 0 / q == 3: tile rows 0 / 127 are never output rows, any finite value will do
                    const float4* xprev = reinterpret_cast<const float4*>(s_xch + ((grp * 4 + (q > 0 ? q - 1 : 0)) * 2) * COUT + c0);
                    const float4* xnext = reinterpret_cast<const float4*>(s_xch + ((grp * 4 + (q < 3 ? q + 1 : 3)) * 2 + 1) * COUT + c0);
                    const bool first = lane == 0, last = lane == 31;
#pragma unroll
                    for (int k4 = 0; k4 < CW / 4; ++k4) {
                        const float4 pu = xprev[k4], pd = xnext[k4];   // same address for all lanes: broadcast
                        const float pus[4] = {pu.x, pu.y, pu.z, pu.w}, pds[4] = {pd.x, pd.y, pd.z, pd.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = 4 * k4 + j;
                            const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[k]), 1);
                            const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[k]), 1);
                            acc[k] = (first ? pus[j] : up) + __uint_as_float(d1[k]) + (last ? pds[j] : dn);
                        }
                    }
                } else {
                    uint32_t r1[CW];
                    tmem_ld_n<CW>(taddr + c0, r1);
                    if constexpr (SKIPG) tmem_ld_n<CW>(taddr + COUT + c0, r2);
                    tmem_ld_wait();
                    if (last_chunk) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_acce + grp);
                        if (q == 0) TDM_TL(EPI, n * C::NACC + grp, 7);
                    }
#pragma unroll
                    for (int k = 0; k < CW; ++k) acc[k] = __uint_as_float(r1[k]);
                }
                if constexpr (EPI == EPI_PLAIN_MASK) {
                    // rows this tile does not own (overlapping kx-combined tiles) and pad rows contribute nothing
                    const bool mine = valid && owned;
                    const uint32_t bits = mw[c0 / 32] >> (c0 & 31);
                    float gp[CW], gt[CW];
#pragma unroll
                    for (int k = 0; k < CW; ++k) {
                        const float g = mine ? acc[k] : 0.f;
                        gp[k] = g;
                        gt[k] = g * ts;
                        acc[k] = (bits >> k) & 1u ? g : 0.f;   // what is stored: the gradient w.r.t. conv1's pre-activation
                    }
                    red_p[ci] += lane_transpose_sum16(gp, lane);
                    red_t[ci] += lane_transpose_sum16(gt, lane);
                    red_m[ci] += lane_transpose_sum16(acc, lane);
                }
#pragma unroll
                for (int pj = 0; pj < CW / 8; ++pj) {
                    const int plane = c0 / 8 + pj;
                    float v[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int ch = c0 + pj * 8 + k;
                        if constexpr (kPlain) {
                            v[k] = acc[pj * 8 + k];
                        } else {
                            v[k] = fmaxf(acc[pj * 8 + k] + (CPAR ? a.cp.bias[ch] : s_bias[ch]), 0.f);
                        }
                    }
                    if (!kPlain && a.mask) {   // training only (uniform branch)
#pragma unroll
                        for (int k = 0; k < 8; ++k) mbits |= (v[k] > 0.f ? 1u : 0u) << ((c0 + pj * 8 + k) & 31);
                    }
                    if constexpr (EPI == EPI_CONV1) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int ch = c0 + pj * 8 + k;
                            v[k] += CPAR ? fmaf(a.cp.tw[ch], ts, a.cp.tb[ch]) : fmaf(s_tw[ch], ts, s_tb[ch]);
                        }
                    } else if constexpr (EPI == EPI_RES_X) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int ch = c0 + pj * 8 + k;
                            v[k] += CPAR ? fmaf(a.cp.aux[ch], xin, a.cp.aux[32 + ch]) : fmaf(s_aux[ch], xin, s_aux[32 + ch]);
                        }
                    } else if constexpr (kHasRes) {
                        const uint32_t* rw = &rv[ci * 2 + pj].x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 f = unpack_bf16x2(rw[k]);
                            v[2 * k] += f.x;
                            v[2 * k + 1] += f.y;
                        }
                    }
                    if constexpr (EPI == EPI_FINAL) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) dot = fmaf(CPAR ? a.cp.aux[c0 + pj * 8 + k] : s_aux[c0 + pj * 8 + k], v[k], dot);
                    }
                    if (EPI != EPI_FINAL || a.out) {
                        uint4 o;
                        o.x = valid ? pack_bf16x2(v[0], v[1]) : 0u;
                        o.y = valid ? pack_bf16x2(v[2], v[3]) : 0u;
                        o.z = valid ? pack_bf16x2(v[4], v[5]) : 0u;
                        o.w = valid ? pack_bf16x2(v[6], v[7]) : 0u;
                        if constexpr (EPI == EPI_RES_UP) {
                            // nearest x2 upsample (src/mnist.py:83): one 14x14 pixel -> 2x2 block of the
                            // 28x28 geometry; pad positions there keep their initial zeros.  Lanes hold
                            // consecutive source pixels whose destinations are 32 B apart, so lane pairs
                            // (2k, 2k+1) take pixel k's value and write the two adjacent 16 B halves:
                            // every store instruction covers contiguous sectors instead of half-sectors.
                            uint8_t* plane_base = a.out + plane * a.out_ps + (int64_t)Geo<28>::GUARD * 16;
#pragma unroll
                            for (int hs = 0; hs < 2; ++hs) {
                                const int src = (lane >> 1) + 16 * hs;
                                uint4 v;
                                v.x = __shfl_sync(0xffffffffu, o.x, src);
                                v.y = __shfl_sync(0xffffffffu, o.y, src);
                                v.z = __shfl_sync(0xffffffffu, o.z, src);
                                v.w = __shfl_sync(0xffffffffu, o.w, src);
                                if (up_ok[hs]) {
                                    uint8_t* dst = plane_base + ((int64_t)up_p00[hs] + (lane & 1)) * 16;
                                    *reinterpret_cast<uint4*>(dst) = v;
                                    *reinterpret_cast<uint4*>(dst + Geo<28>::Wp * 16) = v;
                                }
                            }
                        } else {
                            if (owned) *reinterpret_cast<uint4*>(a.out + plane * a.out_ps + (pos + G::GUARD) * 16) = o;
                        }
                    }
                    if constexpr (SKIPG) {
                        uint4 o2;
                        uint32_t* ow = &o2.x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int ch = c0 + pj * 8 + 2 * k;
                            const float s0 = __uint_as_float(r2[pj * 8 + 2 * k]) + (CPAR ? a.cp.sbias[ch] : s_sbias[ch]);
                            const float s1 = __uint_as_float(r2[pj * 8 + 2 * k + 1]) + (CPAR ? a.cp.sbias[ch + 1] : s_sbias[ch + 1]);
                            ow[k] = valid ? pack_bf16x2(s0, s1) : 0u;
                        }
                        if (owned) *reinterpret_cast<uint4*>(a.out2 + plane * a.out2_ps + (pos + G::GUARD) * 16) = o2;
                    }
                }
                if (!kPlain && a.mask && owned) {
                    // one uint32 per 32 channels per position; this warp owns 16 of its bits
                    uint16_t* m16 = reinterpret_cast<uint16_t*>(a.mask + (c0 / 32) * a.mask_stride + pos);
                    m16[(c0 >> 4) & 1] = valid ? (uint16_t)(mbits >> (c0 & 16)) : (uint16_t)0;
                }
                mbits = 0;
            }
            if (q == 0) TDM_TL(EPI, n * C::NACC + grp, 6);
            if constexpr (EPI == EPI_FINAL) {
                // the two halves each hold a partial dot of the 1x1 out conv: half 1 hands its part over
                if constexpr (kHalves == 2) {
                    float* sd = s_dot + (grp * 2 + (n & 1)) * 128 + trow;
                    if (half == 1) *sd = dot;
                    named_bar_sync(9 + grp, 256);
                    if (half == 0) dot += *sd;
                }
                if (half == 0 && valid) {
                    const float eps = dot + (CPAR ? a.cp.aux[32] : s_aux[32]);  // out conv bias (src/mnist.py:87)
                    const int64_t oi = (int64_t)b * 784 + y * 28 + c;
                    a.fout[oi] = a.fuse_step ? rstep1(sc, xin, eps, zz, add_noise) : eps;
                }
            }
        }
        if constexpr (EPI == EPI_PLAIN_MASK) {
            if (lane < 16) {
#pragma unroll
                for (int ci = 0; ci < kRed; ++ci) {
                    const int ch = (kHalves * ci + half) * CW + lane;
                    if (a.red_plain) atomicAdd(a.red_plain + ch, red_p[ci]);
                    if (a.red_ts) atomicAdd(a.red_ts + ch, red_t[ci]);
                    if (a.red_masked) atomicAdd(a.red_masked + ch, red_m[ci]);
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

template <int W, int CIN, int COUT, int EPI, bool SKIPG, int TAPS = 9, int KXC = 0, int PROD = 0, bool CPAR = false, int CIN2 = 0, int NGRP = kEpiGroups>
static int launch_conv(const ConvArgs& a, cudaStream_t st, const char* name) {
    using C = ConvCfg<W, CIN, COUT, SKIPG, TAPS, KXC, PROD, CIN2, NGRP>;
    auto kern = conv3x3_tc_kernel<W, CIN, COUT, EPI, SKIPG, TAPS, KXC, PROD, CPAR, CIN2, NGRP>;
    TDM_SET_MAX_DYN_SMEM(kern, C::SMEM_BYTES);
    const int nt = (a.np + C::TSTRIDE - 1) / C::TSTRIDE;
    const int grid = nt < num_sms() ? nt : num_sms();
    launch_pdl(kern, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, a);
    TDM_CHECK_LAUNCH(name);
    return TDM_OK;
}

// Forward-pass launch: with a host mirror `hfp` of the flat fp32 parameters (the device copy is `fp`), the
// per-channel vectors travel by value (CPAR); their host addresses follow from the device pointers in `a`.
template <int W, int CIN, int COUT, int EPI, bool SKIPG, int TAPS = 9, int KXC = 0, int PROD = 0, int NGRP = kEpiGroups>
static int launch_conv_fwd(ConvArgs& a, const float* fp, const float* hfp, cudaStream_t st, const char* name) {
    if (!hfp) return launch_conv<W, CIN, COUT, EPI, SKIPG, TAPS, KXC, PROD, false, 0, NGRP>(a, st, name);
    auto mirror = [&](float* dst, const float* dev, int n) {
        if (dev) std::memcpy(dst, hfp + (dev - fp), (size_t)n * sizeof(float));
    };
    mirror(a.cp.bias, a.bias, COUT);
    if (EPI == EPI_CONV1) { mirror(a.cp.tw, a.tw, COUT); mirror(a.cp.tb, a.tb, COUT); }
    if (SKIPG) mirror(a.cp.sbias, a.sbias, COUT);
    if (EPI == EPI_RES_X) { mirror(a.cp.aux, a.aux_w, 32); mirror(a.cp.aux + 32, a.aux_b, 32); }
    if (EPI == EPI_FINAL) { mirror(a.cp.aux, a.aux_w, 32); mirror(a.cp.aux + 32, a.aux_b, 1); }
    return launch_conv<W, CIN, COUT, EPI, SKIPG, TAPS, KXC, PROD, true, 0, NGRP>(a, st, name);
}

}  // namespace tdm
