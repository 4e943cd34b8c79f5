// elementwise.cu — the HBM-bound diffusion math: q_sample, the reverse (ancestral) step, Philox
// normal fill and the final [-1,1] -> [0,1] map.  One pass over memory each, 128-bit accesses,
// schedule coefficients gathered per sample from the reference's fp32 tables.
//
// Bit-exactness: the injected-noise variants restate the reference's separate mul / sub / add /
// div / sqrt ATen ops with the *_rn intrinsics so nvcc cannot contract them into FMAs
// (contraction moves q_sample by up to 4.8e-7).
#include "common.cuh"
#include "diffusion_math.cuh"

namespace tdm {

// One block = one sample x one chunk of kThreads*kUnroll float4s (grid = (batch, chunks)): the sample
// index, its timestep and the schedule coefficients are block-uniform, so nothing per element needs an
// integer division or an IEEE divide/sqrt.  A warp touches 512 contiguous bytes per access.
constexpr int kThreads = 64;
constexpr int kUnroll = 4;  // float4s per thread -> 64 B in flight per stream per thread
constexpr int kChunk4 = kThreads * kUnroll;

__device__ __forceinline__ float4 ldcs4(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));  // streaming: read once
}
__device__ __forceinline__ void stcs4(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}

// ---------------------------------------------------------------------------------------------
// q_sample (src/mnist.py:36-42, src/shakespeare.py:37-44)
// ---------------------------------------------------------------------------------------------
template <bool kPhilox>
__global__ void __launch_bounds__(kThreads)
q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                const int64_t* __restrict__ t, const float* __restrict__ sqrt_acp,
                const float* __restrict__ sqrt_om, float* __restrict__ noise_out,
                float* __restrict__ out, uint32_t inner4, uint64_t seed, uint64_t sample_offset,
                uint32_t stream_id, const int64_t* __restrict__ stream_dev) {
    const int64_t b = blockIdx.x;
    const uint32_t q0 = blockIdx.y * kChunk4 + threadIdx.x;
    if (kPhilox && stream_dev) stream_id += (uint32_t)__ldg(stream_dev);
    const int64_t tb = __ldg(t + b);
    const float ca = __ldg(sqrt_acp + tb), cb = __ldg(sqrt_om + tb);
    const int64_t base = b * (int64_t)inner4;
    float4 xv[kUnroll], nv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const uint32_t q = q0 + u * kThreads;
        if (q < inner4) {
            xv[u] = ldcs4(x0 + (base + q) * 4);
            if constexpr (kPhilox) nv[u] = philox_normal4(seed, sample_offset + (uint64_t)b, q, stream_id, kDomainQSample);
            else nv[u] = ldcs4(noise + (base + q) * 4);
        }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const uint32_t q = q0 + u * kThreads;
        if (q < inner4) {
            float4 o;
            o.x = __fadd_rn(__fmul_rn(ca, xv[u].x), __fmul_rn(cb, nv[u].x));
            o.y = __fadd_rn(__fmul_rn(ca, xv[u].y), __fmul_rn(cb, nv[u].y));
            o.z = __fadd_rn(__fmul_rn(ca, xv[u].z), __fmul_rn(cb, nv[u].z));
            o.w = __fadd_rn(__fmul_rn(ca, xv[u].w), __fmul_rn(cb, nv[u].w));
            stcs4(out + (base + q) * 4, o);
            if constexpr (kPhilox) stcs4(noise_out + (base + q) * 4, nv[u]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// reverse step (src/mnist.py:167-180, src/shakespeare.py:343-352)
// ---------------------------------------------------------------------------------------------
template <bool kPhilox>
__global__ void __launch_bounds__(kThreads)
reverse_step_kernel(const float* __restrict__ x, const float* __restrict__ eps,
                    const float* __restrict__ z, const int64_t* __restrict__ t,
                    const float* __restrict__ betas, const float* __restrict__ alphas,
                    const float* __restrict__ sqrt_om, float* __restrict__ out, uint32_t inner4,
                    uint64_t seed, uint64_t sample_offset, uint32_t step_id) {
    // the reference decides "last step" from t[0] for the whole batch (src/mnist.py:176)
    const bool add_noise = __ldg(t) != 0;
    const int64_t b = blockIdx.x;
    const uint32_t q0 = blockIdx.y * kChunk4 + threadIdx.x;
    const int64_t tb = __ldg(t + b);
    const StepCoef c = step_coef(tb, betas, alphas, sqrt_om);   // block-uniform
    const int64_t base = b * (int64_t)inner4;
    float4 xv[kUnroll], ev[kUnroll], zv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const uint32_t q = q0 + u * kThreads;
        if (q < inner4) {
            xv[u] = ldcs4(x + (base + q) * 4);
            ev[u] = ldcs4(eps + (base + q) * 4);
            zv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (add_noise) {
                if constexpr (kPhilox)
                    zv[u] = philox_normal4(seed, sample_offset + (uint64_t)b, q, step_id + (uint32_t)tb, kDomainReverse);
                else
                    zv[u] = ldcs4(z + (base + q) * 4);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const uint32_t q = q0 + u * kThreads;
        if (q < inner4) {
            float4 o;
            o.x = rstep1(c, xv[u].x, ev[u].x, zv[u].x, add_noise);
            o.y = rstep1(c, xv[u].y, ev[u].y, zv[u].y, add_noise);
            o.z = rstep1(c, xv[u].z, ev[u].z, zv[u].z, add_noise);
            o.w = rstep1(c, xv[u].w, ev[u].w, zv[u].w, add_noise);
            stcs4(out + (base + q) * 4, o);
        }
    }
}

__global__ void __launch_bounds__(kThreads)
randn_kernel(float* __restrict__ out, uint32_t inner4, uint64_t seed, uint64_t sample_offset,
             uint32_t stream_id) {
    const int64_t b = blockIdx.x;
    const uint32_t q0 = blockIdx.y * kChunk4 + threadIdx.x;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const uint32_t q = q0 + u * kThreads;
        if (q < inner4)
            stcs4(out + (b * (int64_t)inner4 + q) * 4,
                  philox_normal4(seed, sample_offset + (uint64_t)b, q, stream_id, kDomainInit));
    }
}

__global__ void __launch_bounds__(256)
unit_range_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
    // (clamp(x,-1,1) + 1) / 2 with the reference's op order (src/mnist.py:194).  The division by two is a multiplication by
    // 0.5 (exact either way, same bits); as an IEEE division it made this 55 instructions per element (ncu, round 2).
    int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = ldcs4(x + i);
        v.x = __fmul_rn(__fadd_rn(fminf(fmaxf(v.x, -1.f), 1.f), 1.f), 0.5f);
        v.y = __fmul_rn(__fadd_rn(fminf(fmaxf(v.y, -1.f), 1.f), 1.f), 0.5f);
        v.z = __fmul_rn(__fadd_rn(fminf(fmaxf(v.z, -1.f), 1.f), 1.f), 0.5f);
        v.w = __fmul_rn(__fadd_rn(fminf(fmaxf(v.w, -1.f), 1.f), 1.f), 0.5f);
        stcs4(out + i, v);
    } else {
        for (; i < n; ++i) out[i] = __fmul_rn(__fadd_rn(fminf(fmaxf(x[i], -1.f), 1.f), 1.f), 0.5f);
    }
}

__global__ void add_i64_kernel(int64_t* __restrict__ t, int64_t n, int64_t delta) {
    pdl_wait();   // the kernel before us still reads t
    pdl_launch_dependents();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] += delta;
}

// ---------------------------------------------------------------------------------------------
// Input pipeline (src/mnist.py:139-147): the reference's DataLoader applies torchvision's
// ToTensor (uint8 -> fp32, true division by 255) and Normalize ((x - mean) / std) to each image on
// the host.  Here the uint8 dataset is resident in HBM and one pass gathers a (shuffled) batch and
// normalises it.  A uint8 pixel has 256 possible values, so every block first builds the 256
// results with the reference's exact op sequence (*_rn: no contraction) and the per-element work is
// one shared-memory lookup: 1 byte read + 4 bytes written per pixel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
u8_gather_normalize_kernel(const uint8_t* __restrict__ images, const int64_t* __restrict__ index,
                           float* __restrict__ out, int64_t n, int row_elems, float mean, float stdv) {
    __shared__ float lut[256];
    lut[threadIdx.x] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)threadIdx.x, 255.f), mean), stdv);
    __syncthreads();
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
        const int64_t src = index ? index[r] : r;
        const uint8_t* in = images + src * row_elems;
        float* o = out + r * row_elems;
        for (int j = threadIdx.x * 4; j < row_elems; j += 256 * 4) {
            const uint32_t v = __ldcs(reinterpret_cast<const unsigned int*>(in + j));
            float4 f;
            f.x = lut[v & 0xff];
            f.y = lut[(v >> 8) & 0xff];
            f.z = lut[(v >> 16) & 0xff];
            f.w = lut[v >> 24];
            *reinterpret_cast<float4*>(o + j) = f;   // default policy: the training step reads it next
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Output step (src/mnist.py:194-199): (clamp(x,-1,1)+1)/2, torchvision.utils.save_image's make_grid
// (single channel tripled, `padding` zero pixels around every image, nrow images per row) and its
// float -> uint8 conversion mul(255).add_(0.5).clamp_(0,255).to(uint8), written as the HWC uint8
// array PIL encodes.  One thread per grid pixel; only the uint8 grid crosses PCIe.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
image_grid_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ grid, int n, int h, int w,
                     int xmaps, int padding, int gh, int gw, int from_signed) {
    const int px = blockIdx.x * 256 + threadIdx.x;
    if (px >= gh * gw) return;
    const int gy = px / gw, gx = px - gy * gw;
    float v = 0.f;   // make_grid's pad_value
    bool inside = false;
    int k = 0, iy = gy, ix = gx;
    if (n == 1) {
        inside = true;   // make_grid returns a single image as it is, without a border
    } else {
        const int ch = h + padding, cw = w + padding;
        const int cy = gy / ch, cx = gx / cw;
        iy = gy - cy * ch - padding;
        ix = gx - cx * cw - padding;
        k = cy * xmaps + cx;
        inside = iy >= 0 && ix >= 0 && cx < xmaps && k < n;
    }
    if (inside) {
        v = x[((int64_t)k * h + iy) * w + ix];
        if (from_signed) v = __fdiv_rn(__fadd_rn(fminf(fmaxf(v, -1.f), 1.f), 1.f), 2.f);
    }
    v = fminf(fmaxf(__fadd_rn(__fmul_rn(v, 255.f), 0.5f), 0.f), 255.f);
    const uint8_t b = (uint8_t)(int)v;   // truncation, like Tensor.to(torch.uint8)
    uint8_t* o = grid + (int64_t)px * 3;
    o[0] = b; o[1] = b; o[2] = b;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline dim3 grid_for(int64_t batch, int64_t inner4) {
    return dim3((unsigned)batch, (unsigned)((inner4 + kChunk4 - 1) / kChunk4));
}
#define TDM_CHECK_EW_SHAPE(who)                                                                          \
    TDM_CHECK_ARG(batch <= 0x7fffffffLL && inner / 4 <= 65535LL * kChunk4, who ": tensor too large for one launch")

}  // namespace tdm

using namespace tdm;

extern "C" int tdm_q_sample(const float* x0, const float* noise, const int64_t* t,
                            const float* sqrt_acp, const float* sqrt_om_acp, float* out,
                            int64_t batch, int64_t inner, int n_steps, void* stream) {
    if (batch == 0) return TDM_OK;
    TDM_CHECK_ARG(x0 && noise && t && sqrt_acp && sqrt_om_acp && out, "tdm_q_sample: null pointer");
    TDM_CHECK_ARG(batch >= 0 && inner > 0 && n_steps > 0, "tdm_q_sample: bad sizes");
    TDM_CHECK_ARG(inner % 4 == 0, "tdm_q_sample: inner (%lld) must be a multiple of 4", (long long)inner);
    TDM_CHECK_ARG(aligned16(x0) && aligned16(noise) && aligned16(out), "tdm_q_sample: 16-byte alignment required");
    TDM_CHECK_EW_SHAPE("tdm_q_sample");
    q_sample_kernel<false><<<grid_for(batch, inner / 4), kThreads, 0, (cudaStream_t)stream>>>(
        x0, noise, t, sqrt_acp, sqrt_om_acp, nullptr, out, (uint32_t)(inner / 4), 0, 0, 0, nullptr);
    TDM_CHECK_LAUNCH("tdm_q_sample");
    return TDM_OK;
}

extern "C" int tdm_q_sample_philox(const float* x0, const int64_t* t, const float* sqrt_acp,
                                   const float* sqrt_om_acp, float* noise_out, float* out,
                                   int64_t batch, int64_t inner, int n_steps, uint64_t seed,
                                   uint64_t sample_offset, uint32_t stream_id,
                                   const int64_t* stream_id_dev, void* stream) {
    if (batch == 0) return TDM_OK;
    TDM_CHECK_ARG(x0 && t && sqrt_acp && sqrt_om_acp && noise_out && out, "tdm_q_sample_philox: null pointer");
    TDM_CHECK_ARG(batch >= 0 && inner > 0 && n_steps > 0, "tdm_q_sample_philox: bad sizes");
    TDM_CHECK_ARG(inner % 4 == 0, "tdm_q_sample_philox: inner must be a multiple of 4");
    TDM_CHECK_ARG(aligned16(x0) && aligned16(noise_out) && aligned16(out), "tdm_q_sample_philox: 16-byte alignment required");
    TDM_CHECK_EW_SHAPE("tdm_q_sample_philox");
    q_sample_kernel<true><<<grid_for(batch, inner / 4), kThreads, 0, (cudaStream_t)stream>>>(
        x0, nullptr, t, sqrt_acp, sqrt_om_acp, noise_out, out, (uint32_t)(inner / 4), seed,
        sample_offset, stream_id, stream_id_dev);
    TDM_CHECK_LAUNCH("tdm_q_sample_philox");
    return TDM_OK;
}

extern "C" int tdm_reverse_step(const float* x, const float* eps, const float* z, const int64_t* t,
                                const float* betas, const float* alphas, const float* sqrt_om_acp,
                                float* out, int64_t batch, int64_t inner, int n_steps, uint64_t seed,
                                uint64_t sample_offset, uint32_t step_id, void* stream) {
    if (batch == 0) return TDM_OK;
    TDM_CHECK_ARG(x && eps && t && betas && alphas && sqrt_om_acp && out, "tdm_reverse_step: null pointer");
    TDM_CHECK_ARG(batch >= 0 && inner > 0 && n_steps > 0, "tdm_reverse_step: bad sizes");
    TDM_CHECK_ARG(inner % 4 == 0, "tdm_reverse_step: inner must be a multiple of 4");
    TDM_CHECK_ARG(aligned16(x) && aligned16(eps) && aligned16(out) && (!z || aligned16(z)),
                  "tdm_reverse_step: 16-byte alignment required");
    TDM_CHECK_EW_SHAPE("tdm_reverse_step");
    if (z)
        reverse_step_kernel<false><<<grid_for(batch, inner / 4), kThreads, 0, (cudaStream_t)stream>>>(
            x, eps, z, t, betas, alphas, sqrt_om_acp, out, (uint32_t)(inner / 4), 0, 0, 0);
    else
        reverse_step_kernel<true><<<grid_for(batch, inner / 4), kThreads, 0, (cudaStream_t)stream>>>(
            x, eps, nullptr, t, betas, alphas, sqrt_om_acp, out, (uint32_t)(inner / 4), seed,
            sample_offset, step_id);
    TDM_CHECK_LAUNCH("tdm_reverse_step");
    return TDM_OK;
}

extern "C" int tdm_randn_philox(float* out, int64_t batch, int64_t inner, uint64_t seed,
                                uint64_t sample_offset, uint32_t stream_id, void* stream) {
    if (batch == 0) return TDM_OK;
    TDM_CHECK_ARG(out, "tdm_randn_philox: null pointer");
    TDM_CHECK_ARG(batch >= 0 && inner > 0 && inner % 4 == 0, "tdm_randn_philox: inner must be a positive multiple of 4");
    TDM_CHECK_ARG(aligned16(out), "tdm_randn_philox: 16-byte alignment required");
    TDM_CHECK_EW_SHAPE("tdm_randn_philox");
    randn_kernel<<<grid_for(batch, inner / 4), kThreads, 0, (cudaStream_t)stream>>>(
        out, (uint32_t)(inner / 4), seed, sample_offset, stream_id);
    TDM_CHECK_LAUNCH("tdm_randn_philox");
    return TDM_OK;
}

extern "C" int tdm_to_unit_range(const float* x, float* out, int64_t n, void* stream) {
    if (n == 0) return TDM_OK;
    TDM_CHECK_ARG(x && out && n >= 0, "tdm_to_unit_range: bad arguments");
    TDM_CHECK_ARG(aligned16(x) && aligned16(out), "tdm_to_unit_range: 16-byte alignment required");
    if (n == 0) return TDM_OK;
    const int64_t nthreads = (n + 3) / 4;
    unit_range_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0,
                        (cudaStream_t)stream>>>(x, out, n);
    TDM_CHECK_LAUNCH("tdm_to_unit_range");
    return TDM_OK;
}

extern "C" int tdm_u8_gather_normalize(const uint8_t* images, const int64_t* index, float* out, int64_t n,
                                       int64_t row_elems, float mean, float stdv, void* stream) {
    TDM_CHECK_ARG(n >= 0 && row_elems > 0 && row_elems % 4 == 0 && row_elems <= (1 << 24),
                  "tdm_u8_gather_normalize: row_elems must be a positive multiple of 4");
    TDM_CHECK_ARG(stdv != 0.f, "tdm_u8_gather_normalize: std must be non-zero");
    if (n == 0) return TDM_OK;
    TDM_CHECK_ARG(images && out, "tdm_u8_gather_normalize: null pointer");
    TDM_CHECK_ARG((reinterpret_cast<uintptr_t>(images) & 3u) == 0 && aligned16(out),
                  "tdm_u8_gather_normalize: images must be 4-byte and out 16-byte aligned");
    const int64_t cap = (int64_t)num_sms() * 8;
    const unsigned grid = (unsigned)(n < cap ? n : cap);
    u8_gather_normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, index, out, n, (int)row_elems, mean, stdv);
    TDM_CHECK_LAUNCH("tdm_u8_gather_normalize");
    return TDM_OK;
}

extern "C" int tdm_image_grid_shape(int64_t n, int h, int w, int nrow, int padding, int* out_h, int* out_w) {
    TDM_CHECK_ARG(n >= 1 && h >= 1 && w >= 1 && nrow >= 1 && padding >= 0 && out_h && out_w,
                  "tdm_image_grid_shape: bad arguments");
    if (n == 1) {
        *out_h = h;
        *out_w = w;
        return TDM_OK;
    }
    const int64_t xmaps = nrow < n ? nrow : n;
    const int64_t ymaps = (n + xmaps - 1) / xmaps;
    const int64_t gh = (h + padding) * ymaps + padding, gw = (w + padding) * xmaps + padding;
    TDM_CHECK_ARG(gh * gw < (1LL << 30), "tdm_image_grid_shape: grid too large");
    *out_h = (int)gh;
    *out_w = (int)gw;
    return TDM_OK;
}

extern "C" int tdm_image_grid_u8(const float* x, uint8_t* grid_hwc, int64_t n, int h, int w, int nrow,
                                 int padding, int from_signed, void* stream) {
    int gh = 0, gw = 0;
    if (int rc = tdm_image_grid_shape(n, h, w, nrow, padding, &gh, &gw)) return rc;
    TDM_CHECK_ARG(x && grid_hwc, "tdm_image_grid_u8: null pointer");
    const int xmaps = (int)(nrow < n ? nrow : n);
    const int64_t px = (int64_t)gh * gw;
    image_grid_u8_kernel<<<(unsigned)((px + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        x, grid_hwc, (int)n, h, w, xmaps, padding, gh, gw, from_signed);
    TDM_CHECK_LAUNCH("tdm_image_grid_u8");
    return TDM_OK;
}

extern "C" int tdm_timestep_advance(int64_t* t, int64_t batch, int64_t delta, void* stream) {
    if (batch == 0) return TDM_OK;
    TDM_CHECK_ARG(t && batch > 0, "tdm_timestep_advance: bad arguments");
    launch_pdl(add_i64_kernel, dim3((unsigned)((batch + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, t, batch, delta);
    TDM_CHECK_LAUNCH("tdm_timestep_advance");
    return TDM_OK;
}
