// Shared helpers for libmobgt (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/mobgt.h"

namespace mobgt {

void set_error(const char *fmt, ...);
void count_launch();   // bumps the process-wide kernel-launch counter (mobgt_launch_count)

inline int32_t cuda_fail(cudaError_t e, const char *what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MOBGT_ERR_CUDA;
}

#define MOBGT_CUDA_OK(expr)                                        \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::mobgt::cuda_fail(_e, #expr); \
    } while (0)

#define MOBGT_LAUNCH_OK(name)                                            \
    do {                                                                 \
        cudaError_t _e = cudaGetLastError();                             \
        if (_e != cudaSuccess) return ::mobgt::cuda_fail(_e, "launch " name); \
        ::mobgt::count_launch();                                         \
    } while (0)

#define MOBGT_REQUIRE(cond, code, ...)   \
    do {                                 \
        if (!(cond)) {                   \
            ::mobgt::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

constexpr int kNumSMs = 148;

// Debug timeline (mobgt_debug_set_timeline): when a device buffer of 256 int64 is registered, thread 0 of ONE CTA of the
// attention kernels stamps clock64() at its pipeline stages (scripts/timeline.py prints the deltas).  NULL = off.
extern long long *g_timeline_dev;
#define MOBGT_STAMP(tl, slot)                                                                              \
    do {                                                                                                   \
        if ((tl) != nullptr && threadIdx.x == 0 && blockIdx.x == gridDim.x / 2 && (slot) < 256) (tl)[(slot)] = clock64(); \
    } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace mobgt
