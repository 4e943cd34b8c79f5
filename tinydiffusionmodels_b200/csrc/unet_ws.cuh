// unet_ws.cuh — activation / gradient workspace layout shared by unet_fwd.cu and unet_bwd.cu.
#pragma once
#include <cuda_runtime.h>

#include "unet_layout.cuh"

namespace tdm {

struct UNetWs {
    int64_t nt28, nt14, ps28, ps14;  // tiles, plane stride in bytes
    int64_t np28, np14;              // positions covered by tiles (nt*128): mask stride
    // forward activations (bf16 planes)
    int64_t t1, cat, p1, t2, s2, h2, t3, t4, s4, h3;
    int64_t fwd_total;
    // training extras: h4, ReLU masks (uint32 per 32 channels per position), gradient scratch
    int64_t h4;
    int64_t m1_1, m2_1, m1_2, m2_2, m1_3, m2_3, m1_4, m2_4;
    int64_t go28, gc28, gh28, gcat;       // 28-level gradients: 32, 32, 32, 96 channels
    int64_t go14a, go14b, gc14, gh14, gp1;  // 14-level gradients: 64, 64, 64, 64, 32 channels
    int64_t gscr;                           // fp32 [P::count]: 3x3 weight gradients as [tap][Cout][Cin] at the tensor's flat offset
    int64_t total;
};

static inline UNetWs make_ws(int64_t batch, bool for_backward) {
    UNetWs w{};
    w.nt28 = num_tiles(batch, Geo<28>::S);
    w.nt14 = num_tiles(batch, Geo<14>::S);
    w.np28 = w.nt28 * kTile;
    w.np14 = w.nt14 * kTile;
    w.ps28 = plane_rows(batch, Geo<28>::S, Geo<28>::GUARD, Geo<28>::TAIL) * 16;
    w.ps14 = plane_rows(batch, Geo<14>::S, Geo<14>::GUARD, Geo<14>::TAIL) * 16;
    int64_t o = 0;
    auto take = [&](int64_t bytes) {
        int64_t at = o;
        o += (bytes + 255) / 256 * 256;
        return at;
    };
    w.t1 = take(4 * w.ps28);
    w.cat = take(12 * w.ps28);
    w.p1 = take(4 * w.ps14);
    w.t2 = take(8 * w.ps14);
    w.s2 = take(8 * w.ps14);
    w.h2 = take(8 * w.ps14);
    w.t3 = take(8 * w.ps14);
    w.t4 = take(4 * w.ps28);
    w.s4 = take(4 * w.ps28);
    w.h3 = take(8 * w.ps14);   // rb3 output at 14x14 (sampling path: upsampled on the fly by rb4.conv1)
    w.fwd_total = o;
    if (for_backward) {
        w.h4 = take(4 * w.ps28);
        w.m1_1 = take(1 * w.np28 * 4);
        w.m2_1 = take(1 * w.np28 * 4);
        w.m1_2 = take(2 * w.np14 * 4);
        w.m2_2 = take(2 * w.np14 * 4);
        w.m1_3 = take(2 * w.np14 * 4);
        w.m2_3 = take(2 * w.np14 * 4);
        w.m1_4 = take(1 * w.np28 * 4);
        w.m2_4 = take(1 * w.np28 * 4);
        w.go28 = take(4 * w.ps28);
        w.gc28 = take(4 * w.ps28);
        w.gh28 = take(4 * w.ps28);
        w.gcat = take(12 * w.ps28);
        w.go14a = take(8 * w.ps14);
        w.go14b = take(8 * w.ps14);
        w.gc14 = take(8 * w.ps14);
        w.gh14 = take(8 * w.ps14);
        w.gp1 = take(4 * w.ps14);
        w.gscr = take(4LL * P::count);   // tap-major weight-gradient scratch, kept all-zero between steps (unet_bwd.cu)
    }
    w.total = o;
    return w;
}

struct StepArgs {
    int train = 0;  // keep ReLU masks + h4 for the backward pass (workspace laid out for_backward)
    int fuse_step = 0;
    const float* z = nullptr;
    const float* betas = nullptr;
    const float* alphas = nullptr;
    const float* sqrt_om = nullptr;
    uint64_t seed = 0, sample_offset = 0;
    uint32_t step_id = 0;
};

// SimpleUNet forward (nine launches); defined in unet_fwd.cu, also used by the training step.
int unet_forward_impl(const uint8_t* wp, const float* x, const int64_t* t, float* fout, uint8_t* ws,
                      int64_t ws_bytes, int64_t batch, const StepArgs& sa, cudaStream_t st);

}  // namespace tdm
