// diffusion_math.cuh — the reverse-step scalar math shared by the standalone elementwise kernel
// and the fused UNet epilogue.  Restates src/mnist.py:169-180 op by op (no FMA contraction).
#pragma once
#include "common.cuh"

namespace tdm {

struct StepCoef {
    float recip_sqrt_alpha;  // 1/sqrt(alphas[t])              src/mnist.py:171
    float eps_coef;          // betas[t]/sqrt(1-alphas_cumprod[t])   src/mnist.py:174
    float sigma;             // sqrt(betas[t])                  src/mnist.py:179-180
};

__device__ __forceinline__ StepCoef step_coef(int64_t tb, const float* __restrict__ betas,
                                              const float* __restrict__ alphas,
                                              const float* __restrict__ sqrt_om) {
    const float beta = __ldg(betas + tb);
    StepCoef c;
    c.recip_sqrt_alpha = __fdiv_rn(1.0f, __fsqrt_rn(__ldg(alphas + tb)));
    c.eps_coef = __fdiv_rn(beta, __ldg(sqrt_om + tb));
    c.sigma = __fsqrt_rn(beta);
    return c;
}

__device__ __forceinline__ float rstep1(const StepCoef& c, float x, float e, float z, bool add_noise) {
    const float mean = __fmul_rn(c.recip_sqrt_alpha, __fsub_rn(x, __fmul_rn(c.eps_coef, e)));
    return add_noise ? __fadd_rn(mean, __fmul_rn(c.sigma, z)) : mean;
}

}  // namespace tdm
