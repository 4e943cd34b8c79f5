// resblock_tc.cuh — a whole ResidualBlock of the 28x28 level (src/mnist.py:45-61) as ONE persistent kernel:
//
//     t = relu(conv1(in) + b1) + time_emb(t)        conv1 on tcgen05, accumulator in TMEM
//     out = relu(conv2(t) + b2) + skip(in)          conv2 on tcgen05, its A operand read from SHARED MEMORY
//
// The intermediate t (and the 1x1 skip of rb4) never reach HBM: conv1's epilogue writes t as bf16 planes into a
// ring of shared-memory tiles, which is exactly the operand layout conv2's MMAs read at their nine row offsets.
// Layer by layer the reverse step moved 13.4 GB at 16,384 images (profiles/r01_traffic_step.json), 130x the
// algorithmic bytes, because t1/t4/s4 round-tripped HBM; this removes 324 KB of the 818 KB per image-step.
//
// One CTA owns a contiguous BAND of tiles of the flattened position axis (unet_layout.cuh) and streams through it.
// Both convolutions use tiles of stride kTS = 126 positions (tile T = positions [126T-1, 126T+127), rows 1..126
// are its output rows) so that conv1's kx-triple schedule and conv2's nine-tap schedule see the same tiles:
//
//   warp 0       producer: weights once; per tile the bulk-copied input (rb4: the 32 skip channels h1 as planes;
//                rb1: the fp32 window of x the tile's 3x3 neighbourhoods fall into)
//   warp 1       MMA issuer, one elected thread.  Step s:  conv1(s)  then  conv2(s-3)
//                  conv2(j) reads t of tiles j, j+1, j+2 (local indices; the band's first and last conv1 tiles
//                  are halo tiles), so it runs three steps behind: the tensor pipe works on conv1(s) while the
//                  epilogue of conv1(s-1) is still converting
//   warps 2..9   epilogue 1 (8 warps = 4 TMEM lane quarters x 2 channel halves), every tile: TMEM -> kx shift-add,
//                bias, ReLU, time embedding -> bf16 planes into the t ring (+ the skip rows into the stash ring)
//   warps 10..17 epilogue 2 (2 groups x 4 warps, alternating tiles): TMEM -> bias, ReLU, + skip, 1x1 out conv,
//                reverse step with in-kernel Philox noise (rb4) / + 1x1 skip of x -> h1 planes (rb1)
//   warps 18..   gather warps: the nearest-x2 upsample of the 14x14 rb3 output into the input stage (rb4) /
//                the im2col rows of the single-channel image, hi/lo bf16 terms, built from the staged window (rb1)
//
// Ring safety needs no barriers of its own - the tensor pipe executes in issue order and the MMA thread is the
// sequencer:  the t ring has 4 tile slots; tile i's last reader is conv2(i), issued at step i+3 BEFORE conv1(i+4),
// and epilogue 1 of tile i+4 (the next writer of the slot) only starts when conv1(i+4) has completed.  The stash
// slot of tile i is read by epilogue 2 of conv2(i-1) before it releases its accumulator; the MMA thread takes that
// release (it needs it for conv2(i+1) anyway) BEFORE it issues conv1(i+4).
#pragma once
#include "conv_tc.cuh"

namespace tdm {

enum : int { RB_KIND_RB1 = 1, RB_KIND_RB4 = 4 };

struct RbChanPar {
    float bias1[32];
    float tw[32];
    float tb[32];
    float sbias[32];   // 1x1 skip bias
    float bias2[32];
    float aux[40];     // rb4: [0,32) out.weight, [32] out.bias;  rb1: [0,32) skip.weight (the 1x1 skip of x)
};

struct RbArgs {
    const uint8_t* in;     // bulk input planes (rb4: h1 = the 32 skip channels), row -GUARD of plane 0
    int64_t in_ps;
    const uint8_t* in2;    // rb4: h3 planes at 14x14 (row -GUARD of plane 0), upsampled on the fly
    int64_t in2_ps;
    const uint8_t* w1;     // conv1 image, kx-triple: [ky][CIN/8][3*32][8]
    const uint8_t* wsk;    // 1x1 skip image [CIN/8][32][8]
    const uint8_t* w2;     // conv2 image, nine taps: [tap][4][32][8]
    const int64_t* t;      // [B]
    const float* x;        // [B,784] fp32: rb4: x_t (reverse step) or null; rb1: the block input (conv1 and 1x1 skip)
    uint8_t* out;          // rb1: h1 planes (row -GUARD of plane 0)
    int64_t out_ps;
    float* fout;           // rb4: [B,784] eps or x_{t-1}
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
    RbChanPar cp;
};

constexpr int kTS = 126;            // tile stride of the fused block (both convolutions)
constexpr int kRingSlots = 4;        // rb4 (shared memory is full); rb1 uses Rb1Cfg::RING_SLOTS
constexpr int kRingMargin = 40;     // mirrored rows in front of / behind the ring (>= 1 + Wp + 1 = 31)
constexpr int kRingRows = kRingMargin + kRingSlots * kTS + kRingMargin;   // 584

struct Rb4Cfg {
    using G = Geo<28>;
    static constexpr int CIN = 96, C = 32;
    static constexpr int NPL = CIN / 8;                              // 12 planes per input stage
    static constexpr int GATHER_PLANES = 8, BULK_PLANES = 4;
    static constexpr int STAGE_BYTES = NPL * G::RT * 16;             // 36,864
    static constexpr int NSTAGE = 2;
    static constexpr int W1_SIDE = NPL * 96 * 16;                    // ky = 0 / ky = 2 block: [12][96][8] bf16
    static constexpr int W1_MID = NPL * 128 * 16;                    // ky = 1 block with the skip rows: [12][128][8]
    static constexpr int W1_BYTES = 2 * W1_SIDE + W1_MID;            // 61,440
    static constexpr int W2_BYTES = 9 * 32 * 32 * 2;                 // 18,432
    static constexpr int RING_SLOTS = kRingSlots, LAG = 3;           // conv2(j) is issued at step j + LAG; slots >= LAG + 1
    static constexpr int RING_ROWS = kRingRows;
    static constexpr int RING_BYTES = 4 * kRingRows * 16;            // t: 4 planes x 584 rows
    static constexpr int STASH_BYTES = 4 * kRingSlots * kTS * 16;    // skip: 4 planes x 504 rows
    static constexpr int XCH_BYTES = 2 * 2 * 4 * 2 * 16 * 4;         // [tile parity][half][quarter][up|down][16 fp32]
    static constexpr int PROD = 2;                                   // gather warps
    static constexpr int EPI1_WARPS = 8, EPI2_WARPS = 8;
    static constexpr int THREADS = 32 * (2 + EPI1_WARPS + EPI2_WARPS + PROD);
    static constexpr int ACC1_COLS = 128, ACC2_COLS = 32;
    static constexpr int TMEM_COLS = 512;                            // 2 x 128 + 2 x 32 = 320 -> next power of two
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
    static constexpr int OFF_IN = OFF_W2 + W2_BYTES;
    static constexpr int OFF_RING = OFF_IN + NSTAGE * STAGE_BYTES;
    static constexpr int OFF_STASH = OFF_RING + RING_BYTES;
    static constexpr int OFF_XCH = OFF_STASH + STASH_BYTES;
    static constexpr int OFF_BAR = OFF_XCH + XCH_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static_assert(SMEM_BYTES <= 227 * 1024, "fused rb4 exceeds shared memory");
    static constexpr int FULL_ARRIVALS = 1 + 32;                     // bulk issuer + one cp.async arrival per gather lane
};

struct Rb1Cfg {
    using G = Geo<28>;
    static constexpr int CIN = 32, C = 32;                           // conv1 as a K = 32 GEMM over the im2col of x
    static constexpr int NPL = 4;
    static constexpr int STAGE_BYTES = NPL * kTile * 16;             // 8,192: im2col rows of the 128 tile rows (no halo)
    static constexpr int NSTAGE = 4;
    static constexpr int XWIN_FLOATS = 256;                          // fp32 window of x per tile (<= 189 pixels + alignment)
    static constexpr int W1_BYTES = 32 * 32 * 2;                     // [4][32][8] hi/hi/lo tap terms (unet_fwd.cu pack)
    static constexpr int W2_BYTES = 9 * 32 * 32 * 2;
    // one more ring slot than rb4 and conv2 one step further behind: the t tile conv2 waits for was finished a whole
    // step earlier, so the epilogue-1 latency (~800 cycles, of a ~1,000-cycle step) is off the MMA thread's path
    static constexpr int RING_SLOTS = 5, LAG = 4;
    static constexpr int RING_ROWS = kRingMargin + RING_SLOTS * kTS + kRingMargin;   // 710
    static constexpr int RING_BYTES = 4 * RING_ROWS * 16;
    static constexpr int STASH_BYTES = 0;
    static constexpr int XCH_BYTES = 0;
    static constexpr int PROD = 4;                                   // im2col warps
    static constexpr int EPI1_WARPS = 8, EPI2_WARPS = 8;
    static constexpr int THREADS = 32 * (2 + EPI1_WARPS + EPI2_WARPS + PROD);
    static constexpr int ACC1_COLS = 32, ACC2_COLS = 32;
    static constexpr int TMEM_COLS = 128;
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = OFF_W1 + W1_BYTES;
    static constexpr int OFF_IN = OFF_W2 + W2_BYTES;
    static constexpr int OFF_XWIN = OFF_IN + NSTAGE * STAGE_BYTES;
    static constexpr int OFF_RING = OFF_XWIN + NSTAGE * XWIN_FLOATS * 4;
    static constexpr int OFF_STASH = OFF_RING + RING_BYTES;
    static constexpr int OFF_XCH = OFF_STASH;
    static constexpr int OFF_BAR = OFF_XCH;
    static constexpr int SMEM_USED = OFF_BAR + 256;
    static constexpr int SMEM_BYTES = SMEM_USED > kSoloSmem ? SMEM_USED : kSoloSmem;   // one CTA per SM (conv_tc.cuh)
    static constexpr int FULL_ARRIVALS = 1;                          // the im2col warp that built the stage
};

template <int KIND> struct RbCfgOf { using type = Rb4Cfg; };
template <> struct RbCfgOf<RB_KIND_RB1> { using type = Rb1Cfg; };

// Index of the first real pixel at or after position `pos` in the [B][784] fp32 image array (monotone in pos):
// pad rows / pad columns map to the next pixel.  Used to bound the window of x a tile's 3x3 neighbourhoods touch.
__device__ __forceinline__ int64_t rb_pixel_lower_bound(int64_t pos, int batch) {
    using G = Geo<28>;
    if (pos <= 0) return 0;
    const int64_t b = pos / G::S;
    if (b >= batch) return (int64_t)batch * 784;
    const int rem = (int)(pos - b * G::S);
    const int rw = rem / G::Wp, c = rem - rw * G::Wp;
    if (rw == 0) return b * 784;                                  // pad row in front of the image
    if (c >= G::W) return b * 784 + (int64_t)rw * 28;             // pad column: first pixel of the next image row
    return b * 784 + (int64_t)(rw - 1) * 28 + c;
}

// Number of 126-position tiles that cover np positions, and CTA c's band of them.
__host__ __device__ inline int rb_num_tiles(int np) { return (np + kTS - 1) / kTS; }

template <int KIND>
__global__ void __launch_bounds__(RbCfgOf<KIND>::type::THREADS, 1) resblock_tc_kernel(const __grid_constant__ RbArgs a) {
    static_assert(KIND == RB_KIND_RB4 || KIND == RB_KIND_RB1, "fused block kinds: rb1, rb4");
    constexpr bool kRb4 = KIND == RB_KIND_RB4;
    using C = typename RbCfgOf<KIND>::type;
    using G = Geo<28>;
    using GS = Geo<14>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* s_w1 = smem + C::OFF_W1;
    uint8_t* s_w2 = smem + C::OFF_W2;
    uint8_t* s_in = smem + C::OFF_IN;
    uint8_t* s_ring = smem + C::OFF_RING;
    uint8_t* s_stash = smem + C::OFF_STASH;
    float* s_xch = reinterpret_cast<float*>(smem + C::OFF_XCH);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* bar_xw_full = bars + 24;            // rb1: NSTAGE (x window landed)
    uint64_t* bar_xw_empty = bars + 28;           // rb1: NSTAGE (x window consumed by the im2col warp)
    uint64_t* bar_w = bars;                       // 1
    uint64_t* bar_full = bars + 1;                // NSTAGE
    uint64_t* bar_empty = bar_full + C::NSTAGE;   // NSTAGE
    uint64_t* bar_acc1f = bar_empty + C::NSTAGE;  // 2
    uint64_t* bar_acc1e = bar_acc1f + 2;          // 2
    uint64_t* bar_tfull = bar_acc1e + 2;          // 2
    uint64_t* bar_acc2f = bar_tfull + 2;          // 2
    uint64_t* bar_acc2e = bar_acc2f + 2;          // 2
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc2e + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // ---- this CTA's band: conv2 / output tiles [j0, j1), conv1 tiles [j0-1, j1+1) ----
    const int ntiles = rb_num_tiles(a.np);
    const int j0 = (int)((int64_t)blockIdx.x * ntiles / gridDim.x);
    const int j1 = (int)((int64_t)(blockIdx.x + 1) * ntiles / gridDim.x);
    const int n2 = j1 - j0;          // conv2 tiles
    const int n1 = n2 + 2;           // conv1 tiles (local i = 0 .. n1-1  <->  global tile j0 - 1 + i)
    const int Tb = j0 - 1;

    // ---- setup (nothing before pdl_wait touches global memory) ----
    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        for (int i = 0; i < C::NSTAGE; ++i) {
            mbar_init(bar_full + i, C::FULL_ARRIVALS);
            mbar_init(bar_empty + i, 1);
            if (!kRb4) {
                mbar_init(bar_xw_full + i, 1);
                mbar_init(bar_xw_empty + i, 1);
            }
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_acc1f + i, 1);
            mbar_init(bar_acc1e + i, C::EPI1_WARPS);   // (unused: t-full doubles as "accumulator 1 free")
            mbar_init(bar_tfull + i, C::EPI1_WARPS);
            mbar_init(bar_acc2f + i, 1);
            mbar_init(bar_acc2e + i, 4);
        }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<C::TMEM_COLS>(s_tmem);
    pdl_wait();
    pdl_launch_dependents();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;

    if (n2 > 0) {
    if (warp == 0) {
        // ===== producer: weights once, then the bulk-copied input of every conv1 tile =====
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_w, C::W1_BYTES + C::W2_BYTES);
            if constexpr (kRb4) {
                // conv1: ky = 0 and ky = 2 blocks as they are ([12][96][8]); the ky = 1 block gets the 32 skip rows
                // behind its 96 (one N = 128 MMA per K step yields the three kx partials AND the 1x1 skip)
                bulk_g2s(s_w1, a.w1, Rb4Cfg::W1_SIDE, bar_w);
                bulk_g2s(s_w1 + Rb4Cfg::W1_SIDE + Rb4Cfg::W1_MID, a.w1 + 2 * Rb4Cfg::W1_SIDE, Rb4Cfg::W1_SIDE, bar_w);
                for (int k = 0; k < Rb4Cfg::NPL; ++k) {
                    bulk_g2s(s_w1 + Rb4Cfg::W1_SIDE + k * 2048, a.w1 + Rb4Cfg::W1_SIDE + k * 1536, 1536, bar_w);
                    bulk_g2s(s_w1 + Rb4Cfg::W1_SIDE + k * 2048 + 1536, a.wsk + k * 512, 512, bar_w);
                }
            } else {
                bulk_g2s(s_w1, a.w1, C::W1_BYTES, bar_w);
            }
            bulk_g2s(s_w2, a.w2, 16384, bar_w);
            bulk_g2s(s_w2 + 16384, a.w2 + 16384, C::W2_BYTES - 16384, bar_w);
        }
        if constexpr (kRb4) {
            for (int i = 0; i < n1; ++i) {
                const int s = i % C::NSTAGE;
                const uint32_t ph = (i / C::NSTAGE) & 1;
                if (lane == 0) {
                    mbar_wait(bar_empty + s, ph ^ 1);
                    mbar_arrive_expect_tx(bar_full + s, Rb4Cfg::BULK_PLANES * G::RT * 16);
                }
                __syncwarp();
                if (lane < Rb4Cfg::BULK_PLANES) {
                    // smem row 0 = global position 126*T - 1 - HALO; the buffers start at row -GUARD
                    const int64_t row = (int64_t)(Tb + i) * kTS - 1 - G::HALO + G::GUARD;
                    bulk_g2s(s_in + s * C::STAGE_BYTES + (Rb4Cfg::GATHER_PLANES + lane) * (G::RT * 16),
                             a.in + lane * a.in_ps + row * 16, G::RT * 16, bar_full + s);
                }
            }
        } else {
            // rb1: the window of x (fp32, contiguous in [B][784]) that the 3x3 neighbourhoods of tile rows
            // [126T-1, 126T+127) fall into: pixels lower_bound(pos0 - 30) .. lower_bound(pos0 + 128 + 30)
            if (lane == 0) {
                for (int i = 0; i < n1; ++i) {
                    const int s = i % C::NSTAGE;
                    const uint32_t ph = (i / C::NSTAGE) & 1;
                    mbar_wait(bar_xw_empty + s, ph ^ 1);
                    const int64_t pos0 = (int64_t)(Tb + i) * kTS - 1;
                    const int64_t lo = rb_pixel_lower_bound(pos0 - 30, a.batch) & ~(int64_t)3;          // 16-byte aligned
                    int64_t hi = (rb_pixel_lower_bound(pos0 + 128 + 30, a.batch) + 3) & ~(int64_t)3;
                    const int64_t total = (int64_t)a.batch * 784;                                          // multiple of 4
                    if (hi > total) hi = total;
                    const uint32_t bytes = hi > lo ? (uint32_t)(hi - lo) * 4 : 0;
                    if (bytes) {
                        mbar_arrive_expect_tx(bar_xw_full + s, bytes);
                        bulk_g2s(smem + Rb1Cfg::OFF_XWIN + s * Rb1Cfg::XWIN_FLOATS * 4, a.x + lo, bytes, bar_xw_full + s);
                    } else {
                        mbar_arrive(bar_xw_full + s);   // tile entirely outside the images: nothing to read
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (elect_one()) {
            constexpr uint32_t idesc_mid = make_idesc_bf16(128, 128);
            constexpr uint32_t idesc_side = make_idesc_bf16(128, 96);
            constexpr uint32_t idesc_c2 = make_idesc_bf16(128, 32);
            mbar_wait(bar_w, 0);
            const uint32_t w1_addr = smem_u32(s_w1), w2_addr = smem_u32(s_w2);
            const uint32_t in_addr = smem_u32(s_in), ring_addr = smem_u32(s_ring);
            // The barriers of step s+1 are PROBED (test_wait) between the MMAs of step s, while the pipe works through
            // what is already queued; a step only blocks on a barrier whose probe failed.  With blocking waits at
            // the top of every step the four already-satisfied waits cost ~600 cycles per step during which the
            // pipe drained (in-kernel timeline, tools/fused_timeline.py: rb1 1,580 cycles per step for 800 of MMAs).
            bool ok_full = false, ok_a2 = false, ok_tf = false;
            int tf_taken = -1;   // t-full phases are taken strictly in tile order, each exactly once
            auto take_tfull = [&](int upto) {
                for (; tf_taken < upto; ++tf_taken) {
                    const int i = tf_taken + 1;
                    if (!(ok_tf && i == upto)) mbar_wait(bar_tfull + (i & 1), (i >> 1) & 1);
                }
                ok_tf = false;
            };
            for (int s = 0; s < n1 + C::LAG; ++s) {
                const int j = s - C::LAG;
                TDM_TL(100 + KIND, s, 0);
                // the accumulator conv2(j) will write: taking its release HERE (before conv1(s) is issued) is what
                // orders epilogue 2 of conv2(j-2) - the reader of stash slot (s-4)%4 - before that slot's next
                // writer, the epilogue of conv1(s)
                if (j >= 0 && j < n2 && !ok_a2) mbar_wait(bar_acc2e + (j & 1), ((j >> 1) & 1) ^ 1);
                TDM_TL(100 + KIND, s, 1);
                if (s < n1) {
                    const int st = s % C::NSTAGE;
                    // "accumulator 1 free" = t-full of tile s-2: epilogue 1 arrives on it only after its TMEM reads.
                    // Taking it here also keeps each t-full barrier at most one phase ahead of this thread.
                    if (s >= 2) take_tfull(s - 2);
                    if (!ok_full) mbar_wait(bar_full + st, (s / C::NSTAGE) & 1);
                    TDM_TL(100 + KIND, s, 3);
                    if constexpr (kRb4) fence_proxy_async_smem();   // cp.async (generic proxy) rows -> async-proxy MMA reads
                    tc_fence_after_sync();
                    uint32_t w_t = w1_addr;
                    asm volatile("" : "+r"(w_t));   // rebuild the descriptors per tile from a uniform address (conv_tc.cuh)
                    const uint32_t d = tmem_base + (s & 1) * C::ACC1_COLS;
                    if constexpr (kRb4) {
                        const uint64_t in_base = make_smem_desc(in_addr + (uint32_t)st * C::STAGE_BYTES, G::RT * 16, 128);
                        const uint64_t wa = make_smem_desc(w_t, 96 * 16, 128);
                        const uint64_t wb = make_smem_desc(w_t + Rb4Cfg::W1_SIDE, 128 * 16, 128);
                        const uint64_t wc = make_smem_desc(w_t + Rb4Cfg::W1_SIDE + Rb4Cfg::W1_MID, 96 * 16, 128);
                        // ky = 1 first: its N = 128 MMAs initialise all four column groups [kx0 | kx1 | kx2 | skip]
#pragma unroll
                        for (int ks = 0; ks < C::CIN / 16; ++ks)
                            umma_bf16(d, desc_add(in_base, (2 * ks) * (G::RT * 16) + G::HALO * 16),
                                      desc_add(wb, (2 * ks) * 2048), idesc_mid, ks != 0);
#pragma unroll
                        for (int ks = 0; ks < C::CIN / 16; ++ks)
                            umma_bf16(d, desc_add(in_base, (2 * ks) * (G::RT * 16) + (G::HALO - G::Wp) * 16),
                                      desc_add(wa, (2 * ks) * 1536), idesc_side, 1u);
                        if (s >= 1 && tf_taken == s - 2) ok_tf = mbar_test(bar_tfull + ((s - 1) & 1), ((s - 1) >> 1) & 1);
#pragma unroll
                        for (int ks = 0; ks < C::CIN / 16; ++ks)
                            umma_bf16(d, desc_add(in_base, (2 * ks) * (G::RT * 16) + (G::HALO + G::Wp) * 16),
                                      desc_add(wc, (2 * ks) * 1536), idesc_side, 1u);
                    } else {
                        // rb1.conv1 as a K = 32 GEMM over the im2col rows (hi/lo bf16 tap terms, conv_tc.cuh)
                        const uint64_t in_base = make_smem_desc(in_addr + (uint32_t)st * C::STAGE_BYTES, kTile * 16, 128);
                        const uint64_t wa = make_smem_desc(w_t, 32 * 16, 128);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_bf16(d, desc_add(in_base, (2 * ks) * (kTile * 16)), desc_add(wa, (2 * ks) * 512), idesc_c2, ks != 0);
                    }
                    umma_commit(bar_empty + st);
                    umma_commit(bar_acc1f + (s & 1));
                    TDM_TL(100 + KIND, s, 4);
                }
                if (j >= 0 && j < n2) {
                    // conv2(j) reads t of tiles j .. j+2 (epilogue-1 threads fenced their generic-proxy stores)
                    take_tfull(j + 2);
                    TDM_TL(100 + KIND, s, 5);
                }
                ok_full = ok_a2 = ok_tf = false;
                if (j >= 0 && j < n2) {
                    tc_fence_after_sync();
                    uint32_t w_t = w2_addr;
                    asm volatile("" : "+r"(w_t));
                    const uint64_t w_base = make_smem_desc(w_t, 32 * 16, 128);
                    // centre tile = local conv1 tile j+1 in ring slot (j+1)%4; tile row 0 = ring row slot*126 - 1
                    const uint32_t row0 = kRingMargin + ((j + 1) % C::RING_SLOTS) * kTS - 1;
                    const uint64_t t_base = make_smem_desc(ring_addr + row0 * 16, C::RING_ROWS * 16, 128);
                    const uint32_t d = tmem_base + 2 * C::ACC1_COLS + (j & 1) * C::ACC2_COLS;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int off = (tap / 3 - 1) * G::Wp + (tap % 3 - 1);
                        // next step's barriers, one probe at a time with a few MMAs queued behind each
                        if (tap == 3 && s + 1 < n1) ok_full = mbar_test(bar_full + (s + 1) % C::NSTAGE, ((s + 1) / C::NSTAGE) & 1);
                        if (tap == 6 && j + 1 < n2) ok_a2 = mbar_test(bar_acc2e + ((j + 1) & 1), (((j + 1) >> 1) & 1) ^ 1);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            // off may be negative: add it as a signed row count to the 14-bit address field (never borrows:
                            // the ring sits far above shared-memory address 0)
                            umma_bf16(d, t_base + (uint64_t)(int64_t)(off + (2 * ks) * C::RING_ROWS),
                                      desc_add(w_base, ((tap * 4 + 2 * ks) * 32) * 16), idesc_c2, (tap | ks) != 0);
                        }
                    }
                    umma_commit(bar_acc2f + (j & 1));
                    TDM_TL(100 + KIND, s, 6);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 2 + C::EPI1_WARPS + C::EPI2_WARPS) {
        const int pw = warp - (2 + C::EPI1_WARPS + C::EPI2_WARPS);
        if constexpr (kRb4) {
        // ===== gather warps: planes 0..7 of the input stage = nearest-x2 upsample of h3 (src/mnist.py:83) =====
        for (int i = pw; i < n1; i += C::PROD) {
            const int s = i % C::NSTAGE;
            const uint32_t ph = (i / C::NSTAGE) & 1;
            mbar_wait(bar_empty + s, ph ^ 1);
            TDM_TL(100 + KIND, i, 12);
            uint8_t* st = s_in + s * C::STAGE_BYTES;
            const int pos0 = (Tb + i) * kTS - 1 - G::HALO;   // may be negative
            // lane's first row decoded once; every further row is 32 positions on: (row, column) += (1, 3) with carries
            int pos = pos0 + lane;
            int b = 0, rw = 0, c = 0;
            {
                const int pp = pos < 0 ? pos + G::S : pos;   // pos0 >= -160-S never happens: tiles start at T >= -1
                b = (int)((uint32_t)pp / (uint32_t)G::S);
                const int rem = (int)((uint32_t)pp - (uint32_t)b * (uint32_t)G::S);
                rw = rem / G::Wp;
                c = rem - rw * G::Wp;
                if (pos < 0) b -= 1;
            }
#pragma unroll 1
            for (int r = lane; r < G::RT; r += 32) {
                const uint8_t* src = a.in2;   // any valid address when the row is zero-filled
                uint32_t nbytes = 0;
                if (b >= 0 && b < a.batch && rw >= 1 && c < G::W) {
                    const int64_t p14 = (int64_t)b * GS::S + ((rw - 1) / 2 + 1) * GS::Wp + c / 2;
                    src = a.in2 + (p14 + GS::GUARD) * 16;
                    nbytes = 16;
                }
#pragma unroll
                for (int pl = 0; pl < Rb4Cfg::GATHER_PLANES; ++pl)
                    cp_async16<true>(st + pl * (G::RT * 16) + r * 16, nbytes ? src + pl * a.in2_ps : src, nbytes);
                c += 3; rw += 1;
                if (c >= G::Wp) { c -= G::Wp; rw += 1; }
                if (rw >= G::Wp) { rw -= G::Wp; b += 1; }
            }
            cp_async_arrive_noinc(bar_full + s);
            TDM_TL(100 + KIND, i, 13);
        }
        } else {
        // ===== im2col warps (rb1): tile row p gets the 3x3 window of x around p as 32 "channels" (conv_tc.cuh:
        //       k 0..8 hi(x) [x hi(w)], 9..17 lo(x) [x hi(w)], 18..26 hi(x) [x lo(w)], 27..31 zero), read from the
        //       fp32 window the producer staged in shared memory =====
        for (int i = pw; i < n1; i += C::PROD) {
            const int s = i % C::NSTAGE;
            const uint32_t ph = (i / C::NSTAGE) & 1;
            mbar_wait(bar_empty + s, ph ^ 1);      // the MMAs that read this stage last have retired
            mbar_wait(bar_xw_full + s, ph);        // this tile's window of x has landed
            TDM_TL(100 + KIND, i, 12);
            uint8_t* st = s_in + s * C::STAGE_BYTES;
            const float* xw = reinterpret_cast<const float*>(smem + Rb1Cfg::OFF_XWIN) + s * Rb1Cfg::XWIN_FLOATS;
            const int64_t pos0 = (int64_t)(Tb + i) * kTS - 1;
            const int64_t wlo = rb_pixel_lower_bound(pos0 - 30, a.batch) & ~(int64_t)3;
#pragma unroll 2
            for (int r = lane; r < kTile; r += 32) {
                const int64_t pos = pos0 + r;
                int b = 0, rw = 0, c = 0;
                if (pos >= 0) {
                    b = (int)((uint32_t)pos / (uint32_t)G::S);
                    const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
                    rw = rem / G::Wp;
                    c = rem - rw * G::Wp;
                }
                const bool ok = pos >= 0 && b < a.batch && rw >= 1 && c < G::W;
                const int e0 = ok ? (int)((int64_t)b * 784 + (rw - 1) * 28 + c - wlo) : 0;   // centre pixel, window-relative
                float hi[9], lo[9];
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const bool in = ok && (unsigned)(rw - 1 + ky - 1) < 28u && (unsigned)(c + kx - 1) < 28u;
                        const float v = in ? xw[e0 + (ky - 1) * 28 + (kx - 1)] : 0.f;
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
                    *reinterpret_cast<uint4*>(st + pl * (kTile * 16) + r * 16) = o;
                }
            }
            fence_proxy_async_smem();   // this lane's generic-proxy stores -> visible to the MMA's async-proxy reads
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_xw_empty + s);   // window consumed (all lanes' reads happened before the __syncwarp)
                mbar_arrive(bar_full + s);
            }
            TDM_TL(100 + KIND, i, 13);
        }
        }
    } else if (warp < 2 + C::EPI1_WARPS) {
        // ===== epilogue 1: every tile; warp = (TMEM lane quarter q, channel half) =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        const int c0 = half * 16;
        for (int i = 0; i < n1; ++i) {
            const int acc = i & 1;
            const int trow = q * 32 + lane;
            const int pos = (Tb + i) * kTS - 1 + trow;
            const bool owned = trow >= 1 && trow <= kTS;
            int b = 0, rr = 0, cc = 0;
            if (pos >= 0) {
                b = (int)((uint32_t)pos / (uint32_t)G::S);
                const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
                rr = rem / G::Wp;
                cc = rem - rr * G::Wp;
            }
            const bool valid = owned && pos >= 0 && b < a.batch && rr >= 1 && cc < G::W;
            float ts = 0.f;
            if (valid) ts = (float)(int)__ldg(a.t + b) / 1000.0f;

            mbar_wait(bar_acc1f + acc, (i >> 1) & 1);
            if (warp == 2) TDM_TL(100 + KIND, i, 7);
            tc_fence_after_sync();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * C::ACC1_COLS;
            float v[16];
            uint32_t sk[16];
            if constexpr (kRb4) {
                uint32_t d0[16], d1[16], d2[16];
                tmem_ld16(taddr + c0, d0);
                tmem_ld16(taddr + 32 + c0, d1);
                tmem_ld16(taddr + 64 + c0, d2);
                tmem_ld16(taddr + 96 + c0, sk);
                tmem_ld_wait();
                if (warp == 2) TDM_TL(100 + KIND, i, 8);
                // out[p] = Y0[p-1] + Y1[p] + Y2[p+1]: neighbour rows are neighbour lanes; across a warp boundary they
                // travel through shared memory (double-buffered by tile parity: one named barrier per tile)
                float* xbuf = s_xch + (((i & 1) * 2 + half) * 4) * 32;          // [quarter][up|down][16]
                float4* xs = reinterpret_cast<float4*>(xbuf + q * 32);
                if (lane == 31) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        xs[k] = make_float4(__uint_as_float(d0[4 * k]), __uint_as_float(d0[4 * k + 1]),
                                            __uint_as_float(d0[4 * k + 2]), __uint_as_float(d0[4 * k + 3]));
                }
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        xs[4 + k] = make_float4(__uint_as_float(d2[4 * k]), __uint_as_float(d2[4 * k + 1]),
                                                __uint_as_float(d2[4 * k + 2]), __uint_as_float(d2[4 * k + 3]));
                }
                named_bar_sync(1 + half, 128);   // the four quarter warps of this channel half
                if (warp == 2) TDM_TL(100 + KIND, i, 14);
                // q == 0 / q == 3: tile rows 0 / 127 are never output rows, any finite value will do
                const float4* xprev = reinterpret_cast<const float4*>(xbuf + (q > 0 ? q - 1 : 0) * 32);
                const float4* xnext = reinterpret_cast<const float4*>(xbuf + (q < 3 ? q + 1 : 3) * 32 + 16);
                const bool first = lane == 0, last = lane == 31;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const float4 pu = xprev[k4], pd = xnext[k4];
                    const float pus[4] = {pu.x, pu.y, pu.z, pu.w}, pds[4] = {pd.x, pd.y, pd.z, pd.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int k = 4 * k4 + jj;
                        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[k]), 1);
                        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[k]), 1);
                        const float accv = (first ? pus[jj] : up) + __uint_as_float(d1[k]) + (last ? pds[jj] : dn);
                        // src/mnist.py:57-59: relu(conv1 + b) + time_emb(t)
                        v[k] = fmaxf(accv + a.cp.bias1[c0 + k], 0.f) + fmaf(a.cp.tw[c0 + k], ts, a.cp.tb[c0 + k]);
                    }
                }
            } else {
                uint32_t d1[16];
                tmem_ld16(taddr + c0, d1);
                tmem_ld_wait();
                if (warp == 2) TDM_TL(100 + KIND, i, 8);
#pragma unroll
                for (int k = 0; k < 16; ++k)
                    v[k] = fmaxf(__uint_as_float(d1[k]) + a.cp.bias1[c0 + k], 0.f) + fmaf(a.cp.tw[c0 + k], ts, a.cp.tb[c0 + k]);
            }
            if (owned) {
                const int slot = i % C::RING_SLOTS;
                const int rrow = slot * kTS + (trow - 1);              // ring row (without the margin)
                uint8_t* tdst = s_ring + (size_t)(kRingMargin + rrow) * 16;
                // mirror: the first rows of slot 0 again behind the ring, the last rows of slot 3 again in front of it
                const int mirror = (rrow < kRingMargin) ? C::RING_SLOTS * kTS : (rrow >= C::RING_SLOTS * kTS - kRingMargin) ? -C::RING_SLOTS * kTS : 0;
#pragma unroll
                for (int pj = 0; pj < 2; ++pj) {
                    uint4 o;
                    o.x = valid ? pack_bf16x2(v[pj * 8 + 0], v[pj * 8 + 1]) : 0u;
                    o.y = valid ? pack_bf16x2(v[pj * 8 + 2], v[pj * 8 + 3]) : 0u;
                    o.z = valid ? pack_bf16x2(v[pj * 8 + 4], v[pj * 8 + 5]) : 0u;
                    o.w = valid ? pack_bf16x2(v[pj * 8 + 6], v[pj * 8 + 7]) : 0u;
                    const int plane = half * 2 + pj;
                    *reinterpret_cast<uint4*>(tdst + (size_t)plane * (C::RING_ROWS * 16)) = o;
                    if (mirror) *reinterpret_cast<uint4*>(tdst + (size_t)plane * (C::RING_ROWS * 16) + mirror * 16) = o;
                    if constexpr (kRb4) {
                        // 1x1 skip of the block input (src/mnist.py:61), kept as bf16 like the layer-by-layer path's s4
                        uint4 o2;
                        uint32_t* ow = &o2.x;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int ch = c0 + pj * 8 + 2 * k;
                            ow[k] = pack_bf16x2(__uint_as_float(sk[pj * 8 + 2 * k]) + a.cp.sbias[ch],
                                                __uint_as_float(sk[pj * 8 + 2 * k + 1]) + a.cp.sbias[ch + 1]);
                        }
                        *reinterpret_cast<uint4*>(s_stash + (size_t)rrow * 16 + (size_t)plane * (kRingSlots * kTS * 16)) = o2;
                    }
                }
            }
            if (warp == 2) TDM_TL(100 + KIND, i, 15);
            fence_proxy_async_smem();   // this thread's generic-proxy stores -> visible to conv2's async-proxy reads
            tc_fence_before_sync();     // ... and its TMEM reads ordered before the arrival that also frees the accumulator
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfull + acc);
            if (warp == 2) TDM_TL(100 + KIND, i, 9);
        }
    } else {
        // ===== epilogue 2: group g = tiles j = g (mod 2); warp = TMEM lane quarter =====
        const int q = warp & 3;
        const int grp = (warp - (2 + C::EPI1_WARPS)) >> 2;
        for (int j = grp; j < n2; j += 2) {
            const int trow = q * 32 + lane;
            const int pos = (j0 + j) * kTS - 1 + trow;
            const bool owned = trow >= 1 && trow <= kTS && pos < a.np;
            int b = 0, rr = 0, cc = 0;
            if (pos >= 0) {
                b = (int)((uint32_t)pos / (uint32_t)G::S);
                const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)G::S);
                rr = rem / G::Wp;
                cc = rem - rr * G::Wp;
            }
            const bool valid = owned && pos >= 0 && b < a.batch && rr >= 1 && cc < G::W;
            const int y = rr - 1;
            float xin = 0.f;
            StepCoef sc{};
            float zz = 0.f;
            bool add_noise = false;
            if constexpr (kRb4) {
                if (valid && a.fuse_step) {
                    xin = __ldg(a.x + (int64_t)b * 784 + y * 28 + cc);
                    add_noise = __ldg(a.t) != 0;  // src/mnist.py:176
                    const int64_t tb = __ldg(a.t + b);
                    sc = step_coef(tb, a.betas, a.alphas, a.sqrt_om);
                    if (add_noise) {
                        const int e = y * 28 + cc;
                        if (a.z) {
                            zz = __ldg(a.z + (int64_t)b * 784 + e);
                        } else {
                            const float4 n4 = philox_normal4(a.seed, a.sample_offset + (uint64_t)b, (uint32_t)(e >> 2),
                                                             a.step_id + (uint32_t)tb, kDomainReverse);
                            const int k = e & 3;
                            zz = k == 0 ? n4.x : k == 1 ? n4.y : k == 2 ? n4.z : n4.w;
                        }
                    }
                }
            } else {
                if (valid) xin = __ldg(a.x + (int64_t)b * 784 + y * 28 + cc);
            }
            mbar_wait(bar_acc2f + grp, (j >> 1) & 1);
            if (q == 2) TDM_TL(100 + KIND, j + C::LAG, 10);
            tc_fence_after_sync();
            uint4 rv[4];
            if constexpr (kRb4) {
                // the skip rows of this tile: written by epilogue 1 of local conv1 tile j+1 (slot (j+1)%4) long before
                // conv2(j) could be issued.  Read BEFORE the accumulator is released (see the MMA warp).
                const int srow = ((j + 1) & 3) * kTS + (trow >= 1 && trow <= kTS ? trow - 1 : 0);
#pragma unroll
                for (int pl = 0; pl < 4; ++pl)
                    rv[pl] = *reinterpret_cast<const uint4*>(s_stash + (size_t)pl * (kRingSlots * kTS * 16) + (size_t)srow * 16);
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + 2 * C::ACC1_COLS + grp * C::ACC2_COLS;
            uint32_t r1[32];
            tmem_ld32(taddr, r1);
            tmem_ld_wait();
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc2e + grp);
            if constexpr (kRb4) {
                float dot = 0.f;
#pragma unroll
                for (int pl = 0; pl < 4; ++pl) {
                    const uint32_t* rw = &rv[pl].x;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 f = unpack_bf16x2(rw[k]);
                        const int ch = pl * 8 + 2 * k;
                        // src/mnist.py:60-61 then the 1x1 out conv (:87)
                        const float v0 = fmaxf(__uint_as_float(r1[ch]) + a.cp.bias2[ch], 0.f) + f.x;
                        const float v1 = fmaxf(__uint_as_float(r1[ch + 1]) + a.cp.bias2[ch + 1], 0.f) + f.y;
                        dot = fmaf(a.cp.aux[ch], v0, dot);
                        dot = fmaf(a.cp.aux[ch + 1], v1, dot);
                    }
                }
                if (valid) {
                    const float eps = dot + a.cp.aux[32];
                    a.fout[(int64_t)b * 784 + y * 28 + cc] = a.fuse_step ? rstep1(sc, xin, eps, zz, add_noise) : eps;
                }
            } else {
                // rb1: h1 = relu(conv2 + b2) + (skip.weight * x + skip.bias)  (src/mnist.py:60-61, 1x1 skip of the image)
                if (owned) {
#pragma unroll
                    for (int pl = 0; pl < 4; ++pl) {
                        float hv[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const int ch = pl * 8 + k;
                            hv[k] = fmaxf(__uint_as_float(r1[ch]) + a.cp.bias2[ch], 0.f) + fmaf(a.cp.aux[ch], xin, a.cp.sbias[ch]);
                        }
                        uint4 o;
                        o.x = valid ? pack_bf16x2(hv[0], hv[1]) : 0u;
                        o.y = valid ? pack_bf16x2(hv[2], hv[3]) : 0u;
                        o.z = valid ? pack_bf16x2(hv[4], hv[5]) : 0u;
                        o.w = valid ? pack_bf16x2(hv[6], hv[7]) : 0u;
                        *reinterpret_cast<uint4*>(a.out + pl * a.out_ps + ((int64_t)pos + G::GUARD) * 16) = o;
                    }
                }
            }
            if (q == 2) TDM_TL(100 + KIND, j + C::LAG, 11);
        }
    }
    }

    // ---- teardown ----
    __syncwarp();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 2) tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int KIND>
static int launch_resblock(const RbArgs& a, cudaStream_t st, const char* name) {
    using C = typename RbCfgOf<KIND>::type;
    auto kern = resblock_tc_kernel<KIND>;
    TDM_SET_MAX_DYN_SMEM(kern, C::SMEM_BYTES);
    const int nt = rb_num_tiles(a.np);
    const int grid = nt < num_sms() ? nt : num_sms();
    launch_pdl(kern, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, a);
    TDM_CHECK_LAUNCH(name);
    return TDM_OK;
}

}  // namespace tdm
