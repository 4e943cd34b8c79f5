// Micro-benchmark: do tcgen05.ld (accumulator drain by epilogue warps) and tcgen05.mma contend?
//   warp 1 issues ITERS MMAs (M=128, N, K=16, operands in smem) into columns [0, N) / [256, 256+N) alternately;
//   warps 2..5 (one per TMEM lane quarter), if enabled, keep reading LDCOLS columns starting at column 384 with
//   tcgen05.ld 32x32b.x32 until the MMA warp is done.  Reports cycles per MMA and the drain rate achieved.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tinydiffusionmodels_b200/csrc \
//        tools/micro/tmem_contention.cu -o tools/micro/tmem_contention
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc05.cuh"
using namespace tdm;

template <int N, bool DRAIN>
__global__ void __launch_bounds__(192) k(long long* out, int iters) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    __shared__ volatile int done;
    __shared__ long long ld_count[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(&slot);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); done = 0; }
    for (int i = threadIdx.x; i < (8192 + 256 * 32) / 4; i += 192) ((uint32_t*)smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (warp == 1) {
        const uint64_t adesc = make_smem_desc(smem_u32(smem), 2048, 128);
        const uint64_t bdesc = make_smem_desc(smem_u32(smem) + 8192, 256 * 16, 128);
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) umma_bf16_elect(tm + (uint32_t)((i & 1) * 128), adesc, bdesc, idesc, 1u);
        umma_commit_elect(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; done = 1; }
    } else if (warp >= 2 && DRAIN) {
        const int q = warp & 3;
        const uint32_t taddr = tm + ((uint32_t)(q * 32) << 16) + 256;
        long long n = 0;
        uint32_t acc = 0;
        while (!done) {
#pragma unroll
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(taddr + c0, r);
                tmem_ld_wait();
                acc ^= r[0] ^ r[31];
            }
            n += 8;
        }
        if (lane == 0) ld_count[q] = n + (acc == 0x12345678u);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (threadIdx.x == 0 && DRAIN) out[blockIdx.x * 2 + 1] = ld_count[0] + ld_count[1] + ld_count[2] + ld_count[3];
    if (warp == 0) tmem_dealloc<512>(tm);
}

template <int N, bool DRAIN>
void run() {
    long long* d; cudaMalloc(&d, 2 * 148 * sizeof(long long)); cudaMemset(d, 0, 2 * 148 * sizeof(long long));
    const int iters = 8192;
    cudaFuncSetAttribute(k<N, DRAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<N, DRAIN><<<148, 192, 65536>>>(d, iters);
    k<N, DRAIN><<<148, 192, 65536>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0, lds = 0; for (int i = 0; i < 148; ++i) { mx = h[2*i] > mx ? h[2*i] : mx; lds += h[2*i+1]; }
    const double cyc = (double)mx / iters;
    // each counted tcgen05.ld moves 32 lanes x 32 columns x 4 B = 4 KB per warp
    printf("N=%3d drain=%d  %.1f cycles/MMA  (alone: %.0f)  drain %.1f B/clk/SM  %s\n", N, (int)DRAIN, cyc,
           N / 2.0 > (4096 + 32.0 * N) / 128 ? N / 2.0 : (4096 + 32.0 * N) / 128,
           DRAIN ? (double)lds / 148 * 4096.0 / (double)mx : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<32, false>(); run<32, true>(); run<96, false>(); run<96, true>(); run<128, false>(); run<128, true>();
    run<256, false>(); run<256, true>();
    return 0;
}
