// unet_layout.cuh — activation / weight layouts of the MNIST UNet kernels.
//
// "Plane" activation layout.  All images of a batch at one resolution are flattened into one
// position axis with a shared zero pad column per row and a shared zero pad row per image:
//     pos(b, y, x) = b*S + (y+1)*Wp + x,   Wp = W+1,   S = (H+1)*Wp
// so that every 3x3 tap is a constant position offset (ky-1)*Wp + (kx-1) and zero padding
// (src/mnist.py:48-49, padding=1) falls out of the pad positions being stored as zeros.
// Channels are stored in groups of 8 bf16 (16 bytes): tensor[plane = c/8][pos][c%8].  A plane
// is a dense array of 16-byte rows, which is exactly the un-swizzled K-major core-matrix layout
// tcgen05.mma reads, with the tap offset applied as a start-address shift.
// Each plane has GUARD zero rows in front and TAIL zero rows behind (never written) so a tile's
// halo read never leaves the allocation.
#pragma once
#include <cstdint>

namespace tdm {

constexpr int kTile = 128;  // positions per MMA tile (UMMA M)

template <int W_>
struct Geo {
    static constexpr int W = W_, H = W_;
    static constexpr int Wp = W_ + 1;
    static constexpr int S = (W_ + 1) * (W_ + 1);
    static constexpr int HALO = (W_ == 28) ? 32 : 16;  // >= Wp + 1
    static constexpr int RT = kTile + 2 * HALO;        // smem rows per plane per tile
    // zero rows in front of position 0 / behind the last tile in HBM (never written).  The fused 28x28 blocks
    // (resblock_tc.cuh) stream 126-row tiles and compute one halo tile on either side of a CTA's band, so their
    // input reads reach 126 + 1 + HALO rows before position 0 and up to 126 + 158 rows past the last position.
    static constexpr int GUARD = (W_ == 28) ? 168 : HALO + 8;
    static constexpr int TAIL = (W_ == 28) ? 296 : HALO + kTile;   // (kx-combined tiles overlap and may start up to 127 rows late)
    static_assert(HALO >= Wp + 1, "halo must cover the largest tap offset");
};

__host__ __device__ inline int64_t num_tiles(int64_t batch, int S) {
    return (batch * S + kTile - 1) / kTile;
}
// rows allocated per plane: GUARD | nt*128 positions | TAIL
__host__ __device__ inline int64_t plane_rows(int64_t batch, int S, int guard, int tail) {
    return num_tiles(batch, S) * kTile + guard + tail;
}

// ---- flat fp32 parameter vector (reference state_dict order) -----------------
namespace P {
constexpr int rb1_c1w = 0;                       // [32,1,3,3]
constexpr int rb1_c1b = rb1_c1w + 288;           // [32]
constexpr int rb1_c2w = rb1_c1b + 32;            // [32,32,3,3]
constexpr int rb1_c2b = rb1_c2w + 9216;
constexpr int rb1_tw = rb1_c2b + 32;             // [32,1]
constexpr int rb1_tb = rb1_tw + 32;
constexpr int rb1_sw = rb1_tb + 32;              // [32,1,1,1]
constexpr int rb1_sb = rb1_sw + 32;
constexpr int rb2_c1w = rb1_sb + 32;             // [64,32,3,3]
constexpr int rb2_c1b = rb2_c1w + 18432;
constexpr int rb2_c2w = rb2_c1b + 64;            // [64,64,3,3]
constexpr int rb2_c2b = rb2_c2w + 36864;
constexpr int rb2_tw = rb2_c2b + 64;
constexpr int rb2_tb = rb2_tw + 64;
constexpr int rb2_sw = rb2_tb + 64;              // [64,32,1,1]
constexpr int rb2_sb = rb2_sw + 2048;
constexpr int rb3_c1w = rb2_sb + 64;             // [64,64,3,3]
constexpr int rb3_c1b = rb3_c1w + 36864;
constexpr int rb3_c2w = rb3_c1b + 64;
constexpr int rb3_c2b = rb3_c2w + 36864;
constexpr int rb3_tw = rb3_c2b + 64;
constexpr int rb3_tb = rb3_tw + 64;
constexpr int rb4_c1w = rb3_tb + 64;             // [32,96,3,3]
constexpr int rb4_c1b = rb4_c1w + 27648;
constexpr int rb4_c2w = rb4_c1b + 32;            // [32,32,3,3]
constexpr int rb4_c2b = rb4_c2w + 9216;
constexpr int rb4_tw = rb4_c2b + 32;
constexpr int rb4_tb = rb4_tw + 32;
constexpr int rb4_sw = rb4_tb + 32;              // [32,96,1,1]
constexpr int rb4_sb = rb4_sw + 3072;
constexpr int out_w = rb4_sb + 32;               // [1,32,1,1]
constexpr int out_b = out_w + 32;                // [1]
constexpr int count = out_b + 1;
static_assert(count == 181473, "SimpleUNet has 181,473 parameters");
}  // namespace P

// ---- packed weight image: bf16 [tap][Cin/8][Cout][8] per tensor-core conv, then fp32 copy -------
namespace WP {
constexpr int64_t conv_bytes(int cin, int cout) { return 9LL * cin * cout * 2; }
constexpr int64_t skip_bytes(int cin, int cout) { return 1LL * cin * cout * 2; }
constexpr int64_t rb1_c2 = 0;
constexpr int64_t rb2_c1 = rb1_c2 + conv_bytes(32, 32);
constexpr int64_t rb2_sk = rb2_c1 + conv_bytes(32, 64);   // must directly follow rb2_c1
constexpr int64_t rb2_c2 = rb2_sk + skip_bytes(32, 64);
constexpr int64_t rb3_c1 = rb2_c2 + conv_bytes(64, 64);
constexpr int64_t rb3_c2 = rb3_c1 + conv_bytes(64, 64);
constexpr int64_t rb4_c1 = rb3_c2 + conv_bytes(64, 64);
constexpr int64_t rb4_sk = rb4_c1 + conv_bytes(96, 32);   // must directly follow rb4_c1
constexpr int64_t rb4_c2 = rb4_sk + skip_bytes(96, 32);
constexpr int64_t fwd_end = rb4_c2 + conv_bytes(32, 32);
// data-gradient convolutions: weights transposed (Cin<->Cout) and tap-flipped, same plane format
// [tap][Cin'/8][Cout'][8] with Cin' = forward Cout, Cout' = forward Cin
constexpr int64_t d_rb1_c2 = fwd_end;                               // 32 -> 32
constexpr int64_t d_rb2_c2 = d_rb1_c2 + conv_bytes(32, 32);         // 64 -> 64
constexpr int64_t d_rb2_c1 = d_rb2_c2 + conv_bytes(64, 64);         // 64 -> 32
constexpr int64_t d_rb2_sk = d_rb2_c1 + conv_bytes(64, 32);         // 64 -> 32 (1x1)
constexpr int64_t d_rb3_c2 = d_rb2_sk + skip_bytes(64, 32);         // 64 -> 64
constexpr int64_t d_rb3_c1 = d_rb3_c2 + conv_bytes(64, 64);         // 64 -> 64
constexpr int64_t d_rb4_c2 = d_rb3_c1 + conv_bytes(64, 64);         // 32 -> 32
constexpr int64_t d_rb4_c1 = d_rb4_c2 + conv_bytes(32, 32);         // 32 -> 96
constexpr int64_t d_rb4_sk = d_rb4_c1 + conv_bytes(32, 96);         // 32 -> 96 (1x1)
// rb1.conv1 as a K = 32 GEMM over the im2col of x: [4][32][8] bf16, k = hi/hi/lo tap terms (conv_tc.cuh)
constexpr int64_t rb1_c1 = d_rb4_sk + skip_bytes(32, 96);
constexpr int64_t bf16_end = rb1_c1 + 32 * 32 * 2;
constexpr int64_t flat = (bf16_end + 255) / 256 * 256;     // fp32 copy of the flat parameters
constexpr int64_t total = (flat + 4LL * P::count + 255) / 256 * 256;
}  // namespace WP

// ---- MMA schedule per 3x3 convolution (conv_tc.cuh) -------------------------------------------------
// bit i of TDM_KXC_MASK: kx-triple (one N = 3*Cout MMA per ky); of TDM_KX2_MASK: kx-pair (N = 2*Cout + N = Cout).
// A tcgen05.mma with both operands in shared memory costs max(N/2, (4096 + 32 N)/128) cycles
// (tools/micro/mma_rate.cu), so small-N taps are bound by re-reading the A tile; sharing it between taps
// helps until the shift-add epilogue / TMEM footprint eats the gain.  Chosen from B200 sweeps
// (profiles/r01_kxc_sweep_B4096.txt, profiles/r01_kx2_sweep_B16384.txt): triple where the MMA count
// dominates (rb4.conv1, K = 9*96), pair on the three 64->64 convolutions (4 x 128 TMEM columns still fit),
// nine taps on the 32->32 ones.
#ifndef TDM_KXC_MASK
#define TDM_KXC_MASK 0x20
#endif
#ifndef TDM_KX2_MASK
#define TDM_KX2_MASK 0x1C
#endif
namespace KX {
// 0 = nine taps, 1 = kx-triple (N = 3*Cout), 2 = kx-pair (N = 2*Cout + N = Cout); conv_tc.cuh
constexpr int sel(int bit) { return ((TDM_KXC_MASK >> bit) & 1) ? 1 : ((TDM_KX2_MASK >> bit) & 1) ? 2 : 0; }
constexpr int rb1c2 = sel(0);
constexpr int rb2c1 = sel(1);
constexpr int rb2c2 = sel(2);
constexpr int rb3c1 = sel(3);
constexpr int rb3c2 = sel(4);
constexpr int rb4c1 = sel(5);
constexpr int rb4c2 = sel(6);
}  // namespace KX

}  // namespace tdm
