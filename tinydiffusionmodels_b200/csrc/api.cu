// api.cu — library-wide state behind the C ABI: version, thread-local error text, launch counter.
#include <atomic>
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace tdm {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached_dev = dev;
        cached_sms = n;
    }
    return cached_sms;
}

}  // namespace tdm

extern "C" int tdm_version(void) { return 1; }
extern "C" const char* tdm_last_error(void) { return tdm::g_err; }
extern "C" int64_t tdm_launch_count(void) { return tdm::g_launches.load(std::memory_order_relaxed); }
