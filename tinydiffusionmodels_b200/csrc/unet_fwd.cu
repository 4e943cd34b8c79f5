// unet_fwd.cu — SimpleUNet.forward (src/mnist.py:76-87) and the fused p_sample
// (src/mnist.py:167-180) as nine launches:
//
//   k1  rb1.conv1   1->32  @28   tcgen05: 8 gather warps build the im2col of x (hi/lo bf16 terms, K = 32) -> t1
//   k2  rb1.conv2   32->32 @28   tcgen05, epilogue relu + 1x1 skip of x           -> h1 (cat[8:12])
//   k3  avg_pool2                                                                  -> p1
//   k4  rb2.conv1   32->64 @14   tcgen05 (+ 1x1 skip GEMM into a 2nd accumulator)  -> t2, s2
//   k5  rb2.conv2   64->64 @14   tcgen05 (kx-pair), epilogue relu + s2             -> h2
//   k6  rb3.conv1   64->64 @14   tcgen05 (kx-pair)                                 -> t3
//   k7  rb3.conv2   64->64 @14   tcgen05 (kx-pair), epilogue relu + h2             -> h3 @14 (sampling)
//                                training: nearest x2 scatter                      -> cat[0:8]
//   k8  rb4.conv1   96->32 @28   tcgen05 (kx-triple, + 1x1 skip GEMM); sampling: 2 gather warps upsample
//                                h3 into the smem tile, h1 arrives by bulk copy    -> t4, s4
//   k9  rb4.conv2   32->32 @28   tcgen05, epilogue relu + s4, 1x1 out conv, and (p_sample) the
//                                reverse-step update with in-kernel Philox noise   -> eps | x_{t-1}
//
// Every convolution is the same persistent warp-specialised kernel (conv_tc.cuh): warp 0 streams input
// tiles (one bulk async copy per 8-channel plane, halo included) into a ring of smem stages,
// warp 1 issues the tcgen05.mma of a tile (M=128 positions, K=16) whose A operand is the *same* smem tile
// addressed at different row offsets per tap, four groups of four warps drain the four TMEM accumulator
// stages through the fused epilogue.  Weights stay resident in smem.
#include <memory>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "diffusion_math.cuh"
#include "tc05.cuh"
#include "unet_layout.cuh"
#include "conv_tc.cuh"
#include "unet_ws.cuh"
#include "resblock_tc.cuh"

// epilogue groups (= tiles in flight) of the two nine-tap 32->32 convolutions, whose 32-column accumulators leave
// TMEM room for more than the default four
#ifndef TDM_G_RB1C2
#define TDM_G_RB1C2 4
#endif
#ifndef TDM_G_RB4C2
#define TDM_G_RB4C2 4
#endif

namespace tdm {

// ---------------------------------------------------------------------------------------------
// weight packing: flat fp32 (state_dict order, OIHW) -> bf16 [tap][Cin/8][Cout][8]
// ---------------------------------------------------------------------------------------------
struct PackJob {
    int src;       // offset of the OIHW tensor in the flat params
    int64_t dst;   // byte offset into wpack
    int cin, cout, taps;  // of the FORWARD convolution
    int transpose;        // 1: emit the data-gradient image (channels swapped, taps flipped)
    int kxc;              // 1: kx-combined image [ky][K/8][3*N][8] (n' = kx*N + n), else [tap][K/8][N][8]
};
constexpr int kPackJobs = 18;
__constant__ PackJob c_jobs[kPackJobs] = {
    {P::rb1_c2w, WP::rb1_c2, 32, 32, 9, 0, KX::rb1c2}, {P::rb2_c1w, WP::rb2_c1, 32, 64, 9, 0, KX::rb2c1},
    {P::rb2_sw, WP::rb2_sk, 32, 64, 1, 0, 0},          {P::rb2_c2w, WP::rb2_c2, 64, 64, 9, 0, KX::rb2c2},
    {P::rb3_c1w, WP::rb3_c1, 64, 64, 9, 0, KX::rb3c1}, {P::rb3_c2w, WP::rb3_c2, 64, 64, 9, 0, KX::rb3c2},
    {P::rb4_c1w, WP::rb4_c1, 96, 32, 9, 0, KX::rb4c1}, {P::rb4_sw, WP::rb4_sk, 96, 32, 1, 0, 0},
    {P::rb4_c2w, WP::rb4_c2, 32, 32, 9, 0, KX::rb4c2},
    // data-gradient images follow the schedule of the forward conv they transpose
    {P::rb1_c2w, WP::d_rb1_c2, 32, 32, 9, 1, KX::rb1c2}, {P::rb2_c2w, WP::d_rb2_c2, 64, 64, 9, 1, KX::rb2c2},
    {P::rb2_c1w, WP::d_rb2_c1, 32, 64, 9, 1, KX::rb2c1}, {P::rb2_sw, WP::d_rb2_sk, 32, 64, 1, 1, 0},
    {P::rb3_c2w, WP::d_rb3_c2, 64, 64, 9, 1, KX::rb3c2}, {P::rb3_c1w, WP::d_rb3_c1, 64, 64, 9, 1, KX::rb3c1},
    {P::rb4_c2w, WP::d_rb4_c2, 32, 32, 9, 1, KX::rb4c2}, {P::rb4_c1w, WP::d_rb4_c1, 96, 32, 9, 1, 0},
    {P::rb4_sw, WP::d_rb4_sk, 96, 32, 1, 1, 0}};

__global__ void pack_weights_kernel(const float* __restrict__ flat, uint8_t* __restrict__ wpack) {
    pdl_wait();   // PDL (common.cuh): first statement, nothing before it touches global memory
    pdl_launch_dependents();
    const PackJob j = c_jobs[blockIdx.y];
    const int n = j.taps * j.cin * j.cout;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(wpack + j.dst);
    // K = contraction channels, N = produced channels of the convolution this image feeds
    const int kch = j.transpose ? j.cout : j.cin;
    const int nch = j.transpose ? j.cin : j.cout;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i & 7;
        int r = i >> 3;
        int nn, cp, tap;
        if (j.kxc) {
            const int n3 = r % (3 * nch);
            r /= 3 * nch;
            cp = r % (kch / 8);
            const int ky = r / (kch / 8);
            const int kx = n3 / nch;
            nn = n3 - kx * nch;
            tap = ky * 3 + kx;
        } else {
            nn = r % nch;
            r /= nch;
            cp = r % (kch / 8);
            tap = r / (kch / 8);
        }
        const int kk = cp * 8 + k;
        const int co = j.transpose ? kk : nn;
        const int ci = j.transpose ? nn : kk;
        const int src_tap = j.transpose ? (j.taps - 1 - tap) : tap;   // 180-degree flip
        dst[i] = __float2bfloat16_rn(flat[j.src + (co * j.cin + ci) * j.taps + src_tap]);
    }
    if (blockIdx.y == 1) {
        // rb1.conv1 [32,1,3,3] as the B operand of the im2col GEMM: [K/8 = 4][32][8], k = tap terms
        // hi(w) (k 0..8, pairs with hi(x)), hi(w) again (k 9..17, pairs with lo(x)), lo(w) (k 18..26, pairs with hi(x))
        __nv_bfloat16* d1 = reinterpret_cast<__nv_bfloat16*>(wpack + WP::rb1_c1);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 32 * 32; i += gridDim.x * blockDim.x) {
            const int k = (i >> 8) * 8 + (i & 7), co = (i >> 3) & 31;
            float v = 0.f;
            if (k < 27) {
                const float w = flat[P::rb1_c1w + co * 9 + k % 9];
                const float h = __bfloat162float(__float2bfloat16_rn(w));
                v = k < 18 ? h : w - h;
            }
            d1[i] = __float2bfloat16_rn(v);
        }
    }
    if (blockIdx.y == 0) {
        float* f = reinterpret_cast<float*>(wpack + WP::flat);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P::count; i += gridDim.x * blockDim.x)
            f[i] = flat[i];
    }
}

// ---------------------------------------------------------------------------------------------
// k3: 2x2 average pool, 28-geometry planes -> 14-geometry planes (src/mnist.py:80)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
avgpool_kernel(const uint8_t* __restrict__ in, int64_t in_ps, uint8_t* __restrict__ out,
               int64_t out_ps, int batch) {
    using GI = Geo<28>;
    using GO = Geo<14>;
    pdl_wait();
    pdl_launch_dependents();
    const int64_t pos = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int plane = blockIdx.y;
    const int b = (int)((uint32_t)pos / (uint32_t)GO::S);   // positions fit 32 bits: division by a constant is a multiply-shift
    const int rem = (int)((uint32_t)pos - (uint32_t)b * (uint32_t)GO::S);
    const int r = rem / GO::Wp, c = rem - r * GO::Wp;
    const bool valid = b < batch && r >= 1 && c < GO::W;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (valid) {
        const int y = r - 1;
        const int64_t p00 = (int64_t)b * GI::S + (2 * y + 1) * GI::Wp + 2 * c;
        const uint8_t* src = in + plane * in_ps + (p00 + GI::GUARD) * 16;
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
    *reinterpret_cast<uint4*>(out + plane * out_ps + (pos + GO::GUARD) * 16) = o;
}

// ---------------------------------------------------------------------------------------------
// host mirrors of the flat parameters, keyed by the packed-weight buffer they describe
// (tdm_unet_pack_weights_host registers, tdm_unet_pack_weights / tdm_unet_forget_host_params drop)
// ---------------------------------------------------------------------------------------------
struct HostMirror {
    const void* wpack;
    std::unique_ptr<float[]> v;   // heap block: its address survives vector growth
};
static std::mutex g_hm_mu;
static std::vector<HostMirror> g_hm;

static const float* find_host_params(const void* wpack) {
    std::lock_guard<std::mutex> lk(g_hm_mu);
    for (const auto& m : g_hm)
        if (m.wpack == wpack) return m.v.get();
    return nullptr;
}
static void drop_host_params(const void* wpack) {
    std::lock_guard<std::mutex> lk(g_hm_mu);
    for (size_t i = 0; i < g_hm.size(); ++i)
        if (g_hm[i].wpack == wpack) { g_hm.erase(g_hm.begin() + (long)i); return; }
}

// ---------------------------------------------------------------------------------------------
// the nine-launch forward
// ---------------------------------------------------------------------------------------------
// optional per-kernel timing (bench.py's roofline): events recorded between the nine launches
static thread_local cudaEvent_t* g_prof = nullptr;

// The 28x28 residual blocks as single kernels (resblock_tc.cuh) unless TDM_UNFUSED=1 / tdm_unet_set_fused(0).
static std::atomic<int> g_fused{-1};
static bool fused_blocks_enabled() {
    int v = g_fused.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* e = std::getenv("TDM_UNFUSED");
        v = (e && e[0] == '1') ? 0 : 1;
        g_fused.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}
#define TDM_PROF(i)                                   \
    do {                                              \
        if (g_prof) cudaEventRecord(g_prof[i], st);   \
    } while (0)

int unet_forward_impl(const uint8_t* wp, const float* x, const int64_t* t, float* fout,
                             uint8_t* ws, int64_t ws_bytes, int64_t batch, const StepArgs& sa,
                             cudaStream_t st) {
    TDM_CHECK_ARG(wp && x && t && fout && ws, "unet_forward: null pointer");
    TDM_CHECK_ARG(batch > 0 && batch <= (1 << 20), "unet_forward: batch %lld out of range", (long long)batch);
    TDM_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 255) == 0 && (reinterpret_cast<uintptr_t>(wp) & 255) == 0,
                  "unet_forward: workspace and wpack must be 256-byte aligned");
    const UNetWs L = make_ws(batch, sa.train != 0);
    TDM_CHECK_ARG(ws_bytes >= L.total, "unet_forward: workspace too small (%lld < %lld)",
                  (long long)ws_bytes, (long long)L.total);
    const float* fp = reinterpret_cast<const float*>(wp + WP::flat);
    // sampling with a registered host mirror: per-channel epilogue parameters go by value (conv_tc.cuh: ChanPar)
    const float* hfp = sa.train ? nullptr : find_host_params(wp);
    const int B = (int)batch;
    const int nt14 = (int)L.nt14;
    int rc;

    const bool fused = !sa.train && hfp && fused_blocks_enabled();
    auto mk = [&](int64_t off) { return sa.train ? reinterpret_cast<uint32_t*>(ws + off) : nullptr; };
    ConvArgs a{};
    if (fused) {
        // k1 + k2 fused: the whole rb1 block in one kernel (resblock_tc.cuh), t1 stays in shared memory -> h1
        TDM_PROF(0);
        RbArgs r{};
        r.x = x; r.t = t; r.w1 = wp + WP::rb1_c1; r.w2 = wp + WP::rb1_c2;
        r.out = ws + L.cat + 8 * L.ps28; r.out_ps = L.ps28;
        r.np = (int)L.np28; r.batch = B;
        std::memcpy(r.cp.bias1, hfp + P::rb1_c1b, 32 * sizeof(float));
        std::memcpy(r.cp.tw, hfp + P::rb1_tw, 32 * sizeof(float));
        std::memcpy(r.cp.tb, hfp + P::rb1_tb, 32 * sizeof(float));
        std::memcpy(r.cp.sbias, hfp + P::rb1_sb, 32 * sizeof(float));
        std::memcpy(r.cp.bias2, hfp + P::rb1_c2b, 32 * sizeof(float));
        std::memcpy(r.cp.aux, hfp + P::rb1_sw, 32 * sizeof(float));
        if ((rc = launch_resblock<RB_KIND_RB1>(r, st, "rb1_fused"))) return rc;
        TDM_PROF(1);
    } else {
    // k1: rb1.conv1 (1 -> 32) + ReLU + time bias -> t1, on the tensor pipe: gather warps build the im2col of x
    TDM_PROF(0);
    a.t = t; a.batch = B; a.np = (int)L.np28;
    a.x = x; a.w = wp + WP::rb1_c1; a.bias = fp + P::rb1_c1b; a.tw = fp + P::rb1_tw; a.tb = fp + P::rb1_tb;
    a.out = ws + L.t1; a.out_ps = L.ps28;
    a.mask = mk(L.m1_1); a.mask_stride = L.np28;
    if ((rc = launch_conv_fwd<28, 32, 32, EPI_CONV1, false, 1, 0, kIm2colWarps>(a, fp, hfp, st, "rb1_conv1"))) return rc;

    TDM_PROF(1);
    a = ConvArgs{};
    a.t = t;
    a.batch = B;
    // k2: rb1.conv2 -> h1 = cat planes 8..11
    a.in = ws + L.t1; a.in_ps = L.ps28; a.w = wp + WP::rb1_c2; a.bias = fp + P::rb1_c2b;
    a.out = ws + L.cat + 8 * L.ps28; a.out_ps = L.ps28;
    a.x = x; a.aux_w = fp + P::rb1_sw; a.aux_b = fp + P::rb1_sb; a.np = (int)L.np28;
    a.mask = mk(L.m2_1); a.mask_stride = L.np28;
    if ((rc = launch_conv_fwd<28, 32, 32, EPI_RES_X, false, 9, KX::rb1c2, 0, TDM_G_RB1C2>(a, fp, hfp, st, "rb1_conv2"))) return rc;

    }

    // k3: pool h1 -> p1
    TDM_PROF(2);
    launch_pdl(avgpool_kernel, dim3(nt14, 4), dim3(128), 0, st, ws + L.cat + 8 * L.ps28, L.ps28, ws + L.p1, L.ps14, B);
    TDM_CHECK_LAUNCH("avgpool");

    // k4: rb2.conv1 (+skip) -> t2, s2
    TDM_PROF(3);
    a = ConvArgs{}; a.t = t; a.batch = B; a.np = (int)L.np14;
    a.in = ws + L.p1; a.in_ps = L.ps14; a.w = wp + WP::rb2_c1; a.bias = fp + P::rb2_c1b;
    a.tw = fp + P::rb2_tw; a.tb = fp + P::rb2_tb; a.sbias = fp + P::rb2_sb;
    a.out = ws + L.t2; a.out_ps = L.ps14; a.out2 = ws + L.s2; a.out2_ps = L.ps14;
    a.mask = mk(L.m1_2); a.mask_stride = L.np14;
    if ((rc = launch_conv_fwd<14, 32, 64, EPI_CONV1, true, 9, KX::rb2c1>(a, fp, hfp, st, "rb2_conv1"))) return rc;

    // k5: rb2.conv2 + s2 -> h2
    TDM_PROF(4);
    a = ConvArgs{}; a.t = t; a.batch = B; a.np = (int)L.np14;
    a.in = ws + L.t2; a.in_ps = L.ps14; a.w = wp + WP::rb2_c2; a.bias = fp + P::rb2_c2b;
    a.res = ws + L.s2; a.res_ps = L.ps14; a.out = ws + L.h2; a.out_ps = L.ps14;
    a.mask = mk(L.m2_2); a.mask_stride = L.np14;
    if ((rc = launch_conv_fwd<14, 64, 64, EPI_RES, false, 9, KX::rb2c2>(a, fp, hfp, st, "rb2_conv2"))) return rc;

    // k6: rb3.conv1 -> t3
    TDM_PROF(5);
    a = ConvArgs{}; a.t = t; a.batch = B; a.np = (int)L.np14;
    a.in = ws + L.h2; a.in_ps = L.ps14; a.w = wp + WP::rb3_c1; a.bias = fp + P::rb3_c1b;
    a.tw = fp + P::rb3_tw; a.tb = fp + P::rb3_tb; a.out = ws + L.t3; a.out_ps = L.ps14;
    a.mask = mk(L.m1_3); a.mask_stride = L.np14;
    if ((rc = launch_conv_fwd<14, 64, 64, EPI_CONV1, false, 9, KX::rb3c1>(a, fp, hfp, st, "rb3_conv1"))) return rc;

    // k7: rb3.conv2 + h2 -> upsampled into cat planes 0..7
    TDM_PROF(6);
    a = ConvArgs{}; a.t = t; a.batch = B; a.np = (int)L.np14;
    a.in = ws + L.t3; a.in_ps = L.ps14; a.w = wp + WP::rb3_c2; a.bias = fp + P::rb3_c2b;
    a.res = ws + L.h2; a.res_ps = L.ps14;
    a.mask = mk(L.m2_3); a.mask_stride = L.np14;
    if (sa.train) {
        // training: the weight gradient of rb4.conv1 needs the concatenated input as a tensor
        a.out = ws + L.cat; a.out_ps = L.ps28;
        if ((rc = launch_conv<14, 64, 64, EPI_RES_UP, false, 9, KX::rb3c2>(a, st, "rb3_conv2"))) return rc;
    } else {
        // sampling: h3 stays at 14x14; rb4.conv1's gather producers upsample it into the smem tile
        a.out = ws + L.h3; a.out_ps = L.ps14;
        if ((rc = launch_conv_fwd<14, 64, 64, EPI_RES, false, 9, KX::rb3c2>(a, fp, hfp, st, "rb3_conv2"))) return rc;
    }

    // k8 + k9 fused (sampling with a host mirror): the whole rb4 block + out conv + reverse step in one kernel,
    // t4 and s4 stay in shared memory (resblock_tc.cuh).  TDM_UNFUSED=1 keeps the layer-by-layer kernels (the
    // per-layer parity tests read t4 / s4 back from the workspace).
    if (fused) {
        TDM_PROF(7);
        RbArgs r{};
        r.in = ws + L.cat + 8 * L.ps28; r.in_ps = L.ps28;
        r.in2 = ws + L.h3; r.in2_ps = L.ps14;
        r.w1 = wp + WP::rb4_c1; r.wsk = wp + WP::rb4_sk; r.w2 = wp + WP::rb4_c2;
        r.t = t; r.x = sa.fuse_step ? x : nullptr; r.fout = fout;
        r.z = sa.z; r.betas = sa.betas; r.alphas = sa.alphas; r.sqrt_om = sa.sqrt_om;
        r.seed = sa.seed; r.sample_offset = sa.sample_offset; r.step_id = sa.step_id; r.fuse_step = sa.fuse_step;
        r.np = (int)L.np28; r.batch = B;
        std::memcpy(r.cp.bias1, hfp + P::rb4_c1b, 32 * sizeof(float));
        std::memcpy(r.cp.tw, hfp + P::rb4_tw, 32 * sizeof(float));
        std::memcpy(r.cp.tb, hfp + P::rb4_tb, 32 * sizeof(float));
        std::memcpy(r.cp.sbias, hfp + P::rb4_sb, 32 * sizeof(float));
        std::memcpy(r.cp.bias2, hfp + P::rb4_c2b, 32 * sizeof(float));
        std::memcpy(r.cp.aux, hfp + P::out_w, 32 * sizeof(float));
        r.cp.aux[32] = hfp[P::out_b];
        if ((rc = launch_resblock<RB_KIND_RB4>(r, st, "rb4_fused"))) return rc;
        TDM_PROF(8);
        TDM_PROF(9);
        return TDM_OK;
    }

    // k8: rb4.conv1 (+skip) -> t4, s4
    TDM_PROF(7);
    a = ConvArgs{}; a.t = t; a.batch = B; a.np = (int)L.np28;
    a.w = wp + WP::rb4_c1; a.bias = fp + P::rb4_c1b;
    a.tw = fp + P::rb4_tw; a.tb = fp + P::rb4_tb; a.sbias = fp + P::rb4_sb;
    a.out = ws + L.t4; a.out_ps = L.ps28; a.out2 = ws + L.s4; a.out2_ps = L.ps28;
    a.mask = mk(L.m1_4); a.mask_stride = L.np28;
    if (sa.train) {
        a.in = ws + L.cat; a.in_ps = L.ps28;
        if ((rc = launch_conv<28, 96, 32, EPI_CONV1, true, 9, KX::rb4c1>(a, st, "rb4_conv1"))) return rc;
    } else {
        a.in = ws + L.cat + 8 * L.ps28; a.in_ps = L.ps28;   // h1 planes (bulk); planes 0..7 gathered from h3
        a.in2 = ws + L.h3; a.in2_ps = L.ps14;
        if ((rc = launch_conv_fwd<28, 96, 32, EPI_CONV1, true, 9, KX::rb4c1, kGatherWarps>(a, fp, hfp, st, "rb4_conv1"))) return rc;
    }

    // k9: rb4.conv2 + s4, out conv, optional reverse step
    TDM_PROF(8);
    a = ConvArgs{}; a.t = t; a.batch = B; a.np = (int)L.np28;
    a.in = ws + L.t4; a.in_ps = L.ps28; a.w = wp + WP::rb4_c2; a.bias = fp + P::rb4_c2b;
    a.res = ws + L.s4; a.res_ps = L.ps28; a.aux_w = fp + P::out_w; a.aux_b = fp + P::out_b;
    a.fout = fout; a.x = sa.fuse_step ? x : nullptr;
    a.mask = mk(L.m2_4); a.mask_stride = L.np28;
    if (sa.train) { a.out = ws + L.h4; a.out_ps = L.ps28; }
    a.fuse_step = sa.fuse_step; a.z = sa.z; a.betas = sa.betas; a.alphas = sa.alphas;
    a.sqrt_om = sa.sqrt_om; a.seed = sa.seed; a.sample_offset = sa.sample_offset; a.step_id = sa.step_id;
    if ((rc = launch_conv_fwd<28, 32, 32, EPI_FINAL, false, 9, KX::rb4c2, 0, TDM_G_RB4C2>(a, fp, hfp, st, "rb4_conv2"))) return rc;
    TDM_PROF(9);
    return TDM_OK;
}

}  // namespace tdm

using namespace tdm;

extern "C" int64_t tdm_unet_param_count(void) { return P::count; }
extern "C" int64_t tdm_unet_wpack_bytes(void) { return WP::total; }
extern "C" int64_t tdm_unet_workspace_bytes(int64_t batch, int for_backward) {
    if (batch <= 0) return 0;
    return make_ws(batch, for_backward != 0).total;
}

extern "C" int tdm_unet_debug_layout(int64_t batch, int64_t* host_out16) {
    TDM_CHECK_ARG(batch > 0 && host_out16, "tdm_unet_debug_layout: bad arguments");
    const UNetWs w = make_ws(batch, false);
    const int64_t v[16] = {w.nt28, w.nt14, w.ps28, w.ps14, w.t1, w.cat, w.p1, w.t2,
                           w.s2,   w.h2,   w.t3,   w.t4,   w.s4, w.total, w.h3, Geo<28>::GUARD};
    for (int i = 0; i < 16; ++i) host_out16[i] = v[i];
    return TDM_OK;
}

extern "C" int tdm_unet_pack_weights(const float* flat_params, void* wpack, void* stream) {
    TDM_CHECK_ARG(flat_params && wpack, "tdm_unet_pack_weights: null pointer");
    drop_host_params(wpack);   // the device copy is about to change: a host mirror of the old values is stale
    launch_pdl(pack_weights_kernel, dim3(32, kPackJobs), dim3(256), 0, (cudaStream_t)stream, 
        flat_params, reinterpret_cast<uint8_t*>(wpack));
    TDM_CHECK_LAUNCH("tdm_unet_pack_weights");
    return TDM_OK;
}

#ifdef TDM_TIMELINE
// development aid: copy the in-kernel timeline (conv_tc.cuh) to host memory
extern "C" int tdm_debug_read_timeline(long long* host_out) {
    TDM_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 96 * 16));
    return TDM_OK;
}
#endif

extern "C" int tdm_unet_set_fused(int on) {
    const int prev = fused_blocks_enabled() ? 1 : 0;
    g_fused.store(on ? 1 : 0, std::memory_order_relaxed);
    return prev;
}

extern "C" int tdm_unet_forget_host_params(const void* wpack) {
    drop_host_params(wpack);
    return TDM_OK;
}

extern "C" int tdm_unet_pack_weights_host(const float* flat_params, const float* flat_params_host, void* wpack,
                                          void* stream) {
    TDM_CHECK_ARG(flat_params_host, "tdm_unet_pack_weights_host: null host mirror");
    const int rc = tdm_unet_pack_weights(flat_params, wpack, stream);   // also drops a previous mirror
    if (rc != TDM_OK) return rc;
    HostMirror m;
    m.wpack = wpack;
    m.v.reset(new float[P::count]);
    std::memcpy(m.v.get(), flat_params_host, sizeof(float) * P::count);
    std::lock_guard<std::mutex> lk(g_hm_mu);
    g_hm.push_back(std::move(m));
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
