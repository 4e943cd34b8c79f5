// Micro-benchmark: how far can one thread run ahead of the tensor pipe?  Cycles to ISSUE k back-to-back tcgen05.mma
// (M=128, N=32 -> 40 cycles of execution each) before the issuing thread is back-pressured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tinydiffusionmodels_b200/csrc tools/micro/mma_queue.cu -o tools/micro/mma_queue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc05.cuh"
using namespace tdm;

template <int N>
__global__ void __launch_bounds__(64) k(long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < (8192 + 256 * 32) / 4; i += 64) ((uint32_t*)smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (warp == 1 && elect_one()) {
        const uint64_t adesc = make_smem_desc(smem_u32(smem), 2048, 128);
        const uint64_t bdesc = make_smem_desc(smem_u32(smem) + 8192, 256 * 16, 128);
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        long long t[65];
        t[0] = clock64();
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            umma_bf16(tm + (uint32_t)((i & 1) * 256), adesc, bdesc, idesc, 1u);
            t[i + 1] = clock64();
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long te = clock64();
        for (int i = 0; i <= 64; ++i) out[i] = t[i] - t[0];
        out[65] = te - t[0];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}

// cycles per "tile" of 18 MMAs when every tile ends with NCOMMIT tcgen05.commit (to rotating mbarriers), as the
// convolution kernels do (stage release + accumulator ready)
template <int N, int NCOMMIT>
__global__ void __launch_bounds__(64) ktile(long long* out, int tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bars[16];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(bars + i, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < (8192 + 256 * 32) / 4; i += 64) ((uint32_t*)smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (warp == 1 && elect_one()) {
        const uint64_t adesc = make_smem_desc(smem_u32(smem), 2048, 128);
        const uint64_t bdesc = make_smem_desc(smem_u32(smem) + 8192, 256 * 16, 128);
        constexpr uint32_t idesc = make_idesc_bf16(128, N);
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
#pragma unroll
            for (int i = 0; i < 18; ++i) umma_bf16(tm + (uint32_t)((t & 3) * 64), adesc, bdesc, idesc, i != 0);
#pragma unroll
            for (int c = 0; c < NCOMMIT; ++c) umma_commit(bars + ((2 * t + c) & 7));
        }
        umma_commit(bars + 15);
        mbar_wait(bars + 15, 0);
        out[0] = clock64() - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}
// the descriptor pattern of a nine-tap 32->32 convolution tile: 4 input planes of 192 rows (LBO = 3072 B), tap offsets
// (ky-1)*29 + (kx-1) rows, weights [tap][4 planes][32 rows] (LBO = 512 B), 4 rotating stages
__global__ void __launch_bounds__(64) kconv(long long* out, int tiles) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bars[16];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(&slot);
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(bars + i, 1); mbar_fence_init(); }
    for (int i = threadIdx.x; i < (4 * 12288 + 18432) / 4; i += 64) ((uint32_t*)smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (warp == 1 && elect_one()) {
        const uint64_t w_base = make_smem_desc(smem_u32(smem) + 4 * 12288, 512, 128);
        constexpr uint32_t idesc = make_idesc_bf16(128, 32);
        const long long t0 = clock64();
        for (int t = 0; t < tiles; ++t) {
            const uint64_t in_base = make_smem_desc(smem_u32(smem) + (t & 3) * 12288, 3072, 128);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                    umma_bf16(tm + (uint32_t)((t & 3) * 32), desc_add(in_base, (2 * ks) * 3072 + (32 + (tap / 3 - 1) * 29 + tap % 3 - 1) * 16),
                              desc_add(w_base, ((tap * 4 + 2 * ks) * 32) * 16), idesc, (tap | ks) != 0);
            umma_commit(bars + ((2 * t) & 7));
            umma_commit(bars + ((2 * t + 1) & 7));
        }
        umma_commit(bars + 15);
        mbar_wait(bars + 15, 0);
        out[0] = clock64() - t0;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}
// the same conv-pattern issue loop while 16 other warps of the SM stream 16-byte global stores / loads
// (what the epilogue warps do): does traffic through the L1 data path slow shared-memory-bound MMAs?
template <int MODE>   // 0 idle, 1 STG.128 stream, 2 LDG.128 stream, 3 bulk g2s copies into a scratch smem area
__global__ void __launch_bounds__(64 + 512) kconv_traffic(long long* out, int tiles, uint4* gbuf) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bars[16];
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(&slot);
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(bars + i, 1); mbar_fence_init(); done = 0; }
    for (int i = threadIdx.x; i < (4 * 12288 + 18432) / 4; i += 576) ((uint32_t*)smem)[i] = 0;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tm = slot;
    if (warp == 1) {
        if (elect_one()) {
            const uint64_t w_base = make_smem_desc(smem_u32(smem) + 4 * 12288, 512, 128);
            constexpr uint32_t idesc = make_idesc_bf16(128, 32);
            const long long t0 = clock64();
            for (int t = 0; t < tiles; ++t) {
                const uint64_t in_base = make_smem_desc(smem_u32(smem) + (t & 3) * 12288, 3072, 128);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks)
                        umma_bf16(tm + (uint32_t)((t & 3) * 32), desc_add(in_base, (2 * ks) * 3072 + (32 + (tap / 3 - 1) * 29 + tap % 3 - 1) * 16),
                                  desc_add(w_base, ((tap * 4 + 2 * ks) * 32) * 16), idesc, (tap | ks) != 0);
                umma_commit(bars + ((2 * t) & 7));
                umma_commit(bars + ((2 * t + 1) & 7));
            }
            umma_commit(bars + 15);
            mbar_wait(bars + 15, 0);
            out[blockIdx.x] = clock64() - t0;
            done = 1;
        }
    } else if (warp >= 2 && MODE != 0) {
        uint4* p = gbuf + (size_t)blockIdx.x * (1 << 16) + (warp - 2) * 4096 + lane;
        uint4 v = make_uint4(warp, lane, 0, 0);
        int i = 0;
        uint32_t acc = 0;
        if (MODE == 3 && warp == 2 && lane == 0) {   // one thread keeps 12 KB bulk copies coming (like the producer warp)
            uint8_t* scratch = smem + 4 * 12288 + 18432;
            uint32_t ph = 0;
            while (!done) {
                mbar_arrive_expect_tx(bars + 14, 12288);
                bulk_g2s(scratch, reinterpret_cast<const uint8_t*>(gbuf) + (size_t)blockIdx.x * (1 << 20) + (i & 63) * 12288, 12288, bars + 14);
                mbar_wait(bars + 14, ph);
                ph ^= 1; ++i;
            }
        } else if (MODE != 3) {
            while (!done) {
                if (MODE == 1) p[(i & 127) * 32] = v;
                else acc ^= p[(i & 127) * 32].x;
                ++i;
            }
            if (acc == 0x12345u) p[0] = v;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tm);
}
template <int MODE>
void run_traffic(const char* what) {
    long long* d; cudaMalloc(&d, 148 * sizeof(long long));
    uint4* g; cudaMalloc(&g, (size_t)148 * (1 << 20)); cudaMemset(g, 0, (size_t)148 * (1 << 20));
    cudaFuncSetAttribute(kconv_traffic<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    const int tiles = 512;
    kconv_traffic<MODE><<<148, 576, 98304>>>(d, tiles, g); kconv_traffic<MODE><<<148, 576, 98304>>>(d, tiles, g);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("conv-pattern tile on 148 SMs, other warps: %-28s %.1f cycles per tile %s\n", what, (double)mx / tiles, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d); cudaFree(g);
}

void run_conv() {
    long long* d; cudaMalloc(&d, sizeof(long long));
    cudaFuncSetAttribute(kconv, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);
    const int tiles = 512;
    kconv<<<1, 64, 98304>>>(d, tiles); kconv<<<1, 64, 98304>>>(d, tiles);
    cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("conv-pattern tile (18 MMAs, N=32, shifted A rows, per-tap B): %.1f cycles per tile (%.1f per MMA) %s\n", (double)h / tiles,
           (double)h / tiles / 18, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

template <int N, int NCOMMIT>
void run_tile() {
    long long* d; cudaMalloc(&d, sizeof(long long));
    cudaFuncSetAttribute(ktile<N, NCOMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int tiles = 512;
    ktile<N, NCOMMIT><<<1, 64, 65536>>>(d, tiles); ktile<N, NCOMMIT><<<1, 64, 65536>>>(d, tiles);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%d, %d commits per 18-MMA tile: %.1f cycles per tile (%.1f per MMA)\n", N, NCOMMIT, (double)h / tiles, (double)h / tiles / 18);
    cudaFree(d);
}

template <int N>
void run() {
    long long* d; cudaMalloc(&d, 66 * sizeof(long long));
    cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<N><<<1, 64, 65536>>>(d); k<N><<<1, 64, 65536>>>(d);
    cudaDeviceSynchronize();
    long long h[66]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%d: cycles after issuing k MMAs:", N);
    for (int i : {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32, 48, 64}) printf(" k=%d:%lld", i, h[i]);
    printf("  | all 64 retired: %lld\n", h[65]);
    cudaFree(d);
}
int main() {
    run<32>(); run<128>();
    run_tile<32, 0>(); run_tile<32, 1>(); run_tile<32, 2>(); run_tile<96, 2>(); run_conv();
    run_traffic<0>("idle"); run_traffic<1>("16 warps of STG.128"); run_traffic<2>("16 warps of LDG.128"); run_traffic<3>("12 KB bulk copies g2s");
    return 0;
}
