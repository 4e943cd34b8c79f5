// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16, both operands in un-swizzled smem) as a
// function of N, and with the A operand in TMEM.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tinydiffusionmodels_b200/csrc \
//        tools/micro/mma_rate.cu -o gpurun_out/mma_rate && gpurun_out/mma_rate
// Answers "is a small-N MMA bound by the shared-memory operand reads rather than the tensor pipe?".
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc05.cuh"
using namespace tdm;

template <int N, bool A_TMEM, int A_OFF = 0, int NCHAIN = 2>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int iters) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < (8192 + 256 * 32) / 4; i += 128) ((uint32_t*)smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (warp == 1) {
        // A: 128 rows x 16 k: two planes of [128][8] -> LBO = 2048, SBO = 128.  B: N rows, planes after A.
        const uint64_t adesc = make_smem_desc(smem_u32(smem) + A_OFF * 16, 2048 + 512, 128);   // A_OFF rows off the 128-byte core-matrix grid
        const uint64_t bdesc = make_smem_desc(smem_u32(smem) + 8192, 256 * 16, 128);
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            if constexpr (A_TMEM) {
                asm volatile(
                    "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
                    "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                    ::"r"(tm + (uint32_t)((i % NCHAIN) * (512 / (NCHAIN < 2 ? 2 : NCHAIN)))), "r"(tm + 256u + 128u), "l"(bdesc), "r"(idesc), "r"(1u)
                    : "memory");
            } else {
                umma_bf16_elect(tm + (uint32_t)((i % NCHAIN) * (512 / (NCHAIN < 2 ? 2 : NCHAIN))), adesc, bdesc, idesc, 1u);
            }
        }
        umma_commit_elect(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, bool A_TMEM, int A_OFF = 0, int NCHAIN = 2>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    const int iters = 4096;
    cudaFuncSetAttribute(rate_kernel<N, A_TMEM, A_OFF, NCHAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int grid : {1, 148}) {
        rate_kernel<N, A_TMEM, A_OFF, NCHAIN><<<grid, 128, 65536>>>(d, iters);
        rate_kernel<N, A_TMEM, A_OFF, NCHAIN><<<grid, 128, 65536>>>(d, iters);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        const double cyc = (double)mx / iters;
        printf("%-10s N=%3d grid=%3d  %.1f cycles/MMA  (tensor-ideal %.1f; smem bytes/MMA %d -> %.0f B/clk)  %s\n", name, N, grid,
               cyc, N / 2.0, (A_TMEM ? 0 : 4096) + N * 32, ((A_TMEM ? 0 : 4096) + N * 32) / cyc,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaFree(d);
}

int main() {
    run<16, false>("A=smem"); run<32, false>("A=smem"); run<64, false>("A=smem"); run<96, false>("A=smem");
    run<128, false>("A=smem"); run<192, false>("A=smem"); run<256, false>("A=smem");
    run<32, true>("A=tmem"); run<64, true>("A=tmem"); run<96, true>("A=tmem"); run<128, true>("A=tmem"); run<256, true>("A=tmem");
    // A tile starting 1..7 rows off the 8-row core-matrix grid (what a conv tap at row offset +-1, +-29 does)
    run<32, false, 1>("A+1row"); run<32, false, 4>("A+4rows"); run<32, false, 7>("A+7rows");
    run<64, false, 1>("A+1row"); run<96, false, 1>("A+1row"); run<128, false, 1>("A+1row");
    // dependent accumulation: every MMA adds into the SAME accumulator (a conv tile's 18 taps) vs 2 / 4 chains
    run<32, false, 0, 1>("1 chain"); run<32, false, 0, 2>("2 chains"); run<32, false, 0, 4>("4 chains");
    run<64, false, 0, 1>("1 chain"); run<64, false, 0, 2>("2 chains");
    run<96, false, 0, 1>("1 chain"); run<96, false, 0, 2>("2 chains");
    run<128, false, 0, 1>("1 chain"); run<128, false, 0, 2>("2 chains");
    return 0;
}
