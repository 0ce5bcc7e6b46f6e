// Library-wide state: version, thread-local error string, device check.
#include "common.cuh"

#include <atomic>

namespace mobgt {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
long long *g_timeline_dev = nullptr;
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace mobgt

extern "C" int32_t mobgt_version(void) { return 100; }

extern "C" int32_t mobgt_last_error(char *buf, size_t buflen) {
    if (!buf || buflen == 0) return MOBGT_ERR_NULL;
    strncpy(buf, mobgt::g_err, buflen - 1);
    buf[buflen - 1] = 0;
    return MOBGT_OK;
}

extern "C" int32_t mobgt_device_check(void) {
    int dev = 0, major = 0, minor = 0;
    MOBGT_CUDA_OK(cudaGetDevice(&dev));
    MOBGT_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    MOBGT_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    MOBGT_REQUIRE(major == 10, MOBGT_ERR_CUDA, "libmobgt is built for sm_100a only; device is sm_%d%d", major, minor);
    return MOBGT_OK;
}

// Number of libmobgt kernels launched by this process so far (bench.py reports the per-step delta).
extern "C" int32_t mobgt_launch_count(int64_t *out) {
    if (!out) return MOBGT_ERR_NULL;
    *out = mobgt::g_launches.load(std::memory_order_relaxed);
    return MOBGT_OK;
}

// Debug hook: register (or clear, with NULL) a device buffer of 256 int64 for the attention kernels' clock64() timeline.
extern "C" int32_t mobgt_debug_set_timeline(void *dev_buf256) {
    mobgt::g_timeline_dev = static_cast<long long *>(dev_buf256);
    return MOBGT_OK;
}
