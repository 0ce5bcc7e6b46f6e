// Library-wide state: version, thread-local error string, device check.
#include "common.cuh"

#include <atomic>
#include <map>
#include <mutex>

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

// A side stream + two events per (device, slot): lets an entry point run two INDEPENDENT kernels next to each other —
// fork: the side stream waits for everything enqueued on the caller's stream so far; join: the caller's stream waits for the
// side stream.  Nothing synchronises with the host, and the pattern is legal inside a stream capture (the side stream joins the
// capture between fork and join).  The objects are created on first use, which must not fall inside a capture (the trainer
// warms every captured step up eagerly first).
int32_t get_fork_join(int slot, ForkJoin *out) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, ForkJoin> cache;
    int dev = 0;
    MOBGT_CUDA_OK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, slot});
    if (it == cache.end()) {
        ForkJoin fj{};
        MOBGT_CUDA_OK(cudaStreamCreateWithFlags(&fj.side, cudaStreamNonBlocking));
        MOBGT_CUDA_OK(cudaEventCreateWithFlags(&fj.fork, cudaEventDisableTiming));
        MOBGT_CUDA_OK(cudaEventCreateWithFlags(&fj.join, cudaEventDisableTiming));
        it = cache.emplace(std::make_pair(dev, slot), fj).first;
    }
    *out = it->second;
    return MOBGT_OK;
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
