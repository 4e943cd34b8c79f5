// common.cuh — error plumbing, launch accounting and the counter-based RNG shared by all kernels.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/tdm_b200.h"

namespace tdm {

// ---- host-side error state (thread local) ---------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define TDM_CHECK_ARG(cond, ...)          \
    do {                                  \
        if (!(cond)) {                    \
            ::tdm::set_error(__VA_ARGS__); \
            return TDM_ERR_ARG;           \
        }                                 \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A kernel launched through launch_pdl may become resident while its
// predecessor in the stream is still draining; everything before pdl_wait() (barrier init, TMEM allocation)
// then overlaps the predecessor's tail and the launch latency.  Rules kept by every kernel that uses it:
//   * no global-memory access before pdl_wait();
//   * pdl_launch_dependents() only AFTER pdl_wait(), so a kernel never overlaps anything but its direct
//     predecessor (which therefore has itself waited for everything older).
// Kernels launched the ordinary way see both instructions as no-ops.  Set TDM_NO_PDL=1 to launch everything
// the ordinary way (debugging aid).
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("TDM_NO_PDL");
        return !(e && e[0] == '1');
    }();
    return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

#define TDM_CHECK_LAUNCH(name)                                                          \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            ::tdm::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
            return TDM_ERR_CUDA;                                                        \
        }                                                                               \
        ::tdm::count_launch();                                                          \
    } while (0)

#define TDM_CHECK_CUDA(call)                                                             \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            ::tdm::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
            return TDM_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

int num_sms();  // cached cudaDevAttrMultiProcessorCount of the current device

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: set it once per (kernel instantiation,
// device ordinal), lock-free (a racing thread at worst sets it twice).  The static lives in the enclosing
// function, i.e. one mask per template instantiation of the launcher.
#define TDM_SET_MAX_DYN_SMEM(kern, bytes)                                                              \
    do {                                                                                               \
        static std::atomic<unsigned long long> done__{0ull};                                           \
        int dev__ = 0;                                                                                 \
        TDM_CHECK_CUDA(cudaGetDevice(&dev__));                                                         \
        const unsigned long long bit__ = 1ull << (dev__ & 63);                                         \
        if (!(done__.load(std::memory_order_acquire) & bit__)) {                                       \
            TDM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)); \
            done__.fetch_or(bit__, std::memory_order_release);                                         \
        }                                                                                              \
    } while (0)

// ---- Philox4x32-10 (Salmon et al. 2011), hand-rolled so the numpy oracle can mirror it --------
// counter = (c0, c1, c2, c3), key = (k0, k1).  Same constants as Random123 / cuRAND.
struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                         uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0;
        const uint64_t p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += W0;
        k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// 23-bit uniform in (0,1): (bits>>9 + 0.5) * 2^-23 — every value is exactly representable in fp32,
// the smallest is 2^-24 and the largest 1 - 2^-24, so it is never 0 and never 1.  (A 24-bit version
// rounds its top value to exactly 1.0, log gives 0 and the Box-Muller radius degenerates.)
__host__ __device__ __forceinline__ float u01(uint32_t bits) {
    return ((float)(bits >> 9) + 0.5f) * (1.0f / 8388608.0f);
}

// Four N(0,1) from one Philox block via two Box–Muller pairs:
//   (n0, n1) = r(x) * (cos, sin)(2*pi*u(y)),  (n2, n3) = r(z) * (cos, sin)(2*pi*u(w)).
// Device code uses the MUFU fast paths; the oracle uses libm and the tests carry the tolerance.
#ifdef __CUDACC__
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t sample, uint32_t quad,
                                                 uint32_t step, uint32_t domain) {
    // counter = (quad, sample_lo, step, domain | sample_hi<<8): sample index may exceed 2^32
    const uint32_t c3 = domain | ((uint32_t)(sample >> 32) << 8);
    Philox4 r = philox4x32_10(quad, (uint32_t)sample, step, c3, (uint32_t)seed,
                              (uint32_t)(seed >> 32));
    // radius sqrt(-2 ln u): MUFU square root (sqrt.approx: 0 -> 0, no NaN) instead of the ~10-instruction
    // IEEE sqrtf; the oracle uses libm and the tests carry the 2e-5 tolerance
    const float r0 = sqrt_approx(-2.0f * __logf(u01(r.x)));
    const float r1 = sqrt_approx(-2.0f * __logf(u01(r.z)));
    float s0, c0, s1, c1;
    __sincosf(6.283185307179586f * u01(r.y), &s0, &c0);
    __sincosf(6.283185307179586f * u01(r.w), &s1, &c1);
    return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}
#endif

constexpr uint32_t kDomainQSample = 0;   // training noise
constexpr uint32_t kDomainReverse = 1;   // per-step posterior noise z
constexpr uint32_t kDomainInit = 2;      // x_T

}  // namespace tdm
