// peer.cu — peer-mapped gradient buffers for the data-parallel training step (one process per GPU).
//
// The reference's DDP-less loop has no multi-GPU path; ours shards the batch and has to sum one 726 KB flat
// gradient per step.  Instead of a separate collective, every rank keeps its gradient in a buffer the other
// ranks map through CUDA IPC, and the optimizer kernel (tdm_adamw_flat_peer, unet_bwd.cu) reads all ranks'
// gradients over NVLink while it applies the update: "all-reduce + AdamW" is one kernel.
//
// Buffer layout (tdm_peer_buffer_bytes):   [ flags: 64 x u64 | grad parity 0: Npad x f32 | grad parity 1: Npad x f32 ]
//   flags[r] on rank q  = the last step k for which rank r has finished writing grad[k & 1]   (written BY rank r)
// Two gradient slots because a fast rank starts the next backward while a slow peer may still be reading.
#include <cstring>
#include "common.cuh"

using namespace tdm;

namespace {
constexpr int64_t kFlagBytes = 512;
inline int64_t npad(int64_t n) { return (n + 63) / 64 * 64; }
}  // namespace

extern "C" int64_t tdm_peer_buffer_bytes(int64_t n) { return n <= 0 ? 0 : kFlagBytes + 2 * npad(n) * 4; }
extern "C" int64_t tdm_peer_grad_offset(int64_t n, int parity) { return kFlagBytes + (parity & 1) * npad(n) * 4; }

extern "C" int tdm_peer_alloc(int64_t bytes, void** out_ptr) {
    TDM_CHECK_ARG(bytes > 0 && out_ptr, "tdm_peer_alloc: bad arguments");
    void* p = nullptr;
    TDM_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));   // a whole allocation of its own: IPC handles map allocations
    TDM_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
    *out_ptr = p;
    return TDM_OK;
}

extern "C" int tdm_peer_free(void* ptr) {
    if (ptr) TDM_CHECK_CUDA(cudaFree(ptr));
    return TDM_OK;
}

extern "C" int tdm_peer_export(const void* ptr, void* host_handle64) {
    TDM_CHECK_ARG(ptr && host_handle64, "tdm_peer_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    TDM_CHECK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    std::memcpy(host_handle64, &h, 64);
    return TDM_OK;
}

extern "C" int tdm_peer_import(const void* host_handle64, void** out_ptr) {
    TDM_CHECK_ARG(host_handle64 && out_ptr, "tdm_peer_import: null pointer");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, host_handle64, 64);
    void* p = nullptr;
    TDM_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *out_ptr = p;
    return TDM_OK;
}

extern "C" int tdm_peer_close(void* ptr) {
    if (ptr) TDM_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return TDM_OK;
}
