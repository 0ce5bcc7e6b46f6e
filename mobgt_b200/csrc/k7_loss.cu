// K7 — the training losses of MobGT over the POI / category logits (sm_100a, HBM-bound).
//
// Replaces:
//   * toyotagraph: log_softmax over the POI logits (model_fqandtoyo.py:1425) + NLLLoss(ignore_index=0) (data.py:165,
//     model_fqandtoyo.py:1470-1471) — in torch: log_softmax fwd (read + write [B,V]), nll gather, log_softmax bwd (two reads +
//     one write), nll bwd (a zero-filled [B,V] + scatter);
//   * GradientTailLoss (model_fqandtoyo.py:545-550; alpha = 0.2 on the POI logits of foursquaregraph / gowalla :1447-1460,
//     alpha = 0.1 on the category logits of toyotagraph :1464-1469) — in torch: zeros + scatter one-hot + sigmoid + two logs +
//     6 elementwise ops and their autograd mirror, each a full pass over [B,V].
// Here every loss is ONE read of the logits in forward (per-row-chunk partials, merged in a fixed order: bitwise reproducible)
// and ONE read + ONE write in backward (d logits, already scaled by the upstream gradient read from device memory, so the
// call is CUDA-graph capturable and needs no host value).
#include <cuda_bf16.h>

#include "common.cuh"

namespace mobgt {

template <typename T>
__device__ __forceinline__ float ldf(const T *p);
template <>
__device__ __forceinline__ float ldf<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stf(T *p, float v);
template <>
__device__ __forceinline__ void stf<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int kLossThreads = 256;

// merge two online-softmax states (m, s): s * exp(m) is the partial sum of exp(x)
__device__ __forceinline__ void lse_merge(float &m, float &s, float m2, float s2) {
    const float M = fmaxf(m, m2);
    if (M == -INFINITY) { m = M; s = 0.f; return; }
    s = s * __expf(m - M) + s2 * __expf(m2 - M);
    m = M;
}

// ---- log_softmax + NLL, forward: partial (max, sum exp) of chunk `blockIdx.x` of row `blockIdx.y`
template <typename T>
__global__ void __launch_bounds__(kLossThreads) k7_lse_partial_kernel(const T *__restrict__ x, int64_t stride, int V, int chunk,
                                                                     float2 *__restrict__ part) {
    const int row = blockIdx.y, c0 = blockIdx.x * chunk, c1 = min(V, c0 + chunk);
    const T *xr = x + (size_t)row * stride;
    float m = -INFINITY, s = 0.f;
    for (int c = c0 + threadIdx.x; c < c1; c += kLossThreads) {
        const float v = ldf(xr + c);
        if (v > m) { s = s * __expf(m - v) + 1.f; m = v; }      // m == -inf: s == 0, exp(-inf) == 0
        else if (m > -INFINITY) s += __expf(v - m);
    }
    __shared__ float sm[kLossThreads / 32], ss[kLossThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lse_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kLossThreads / 32; ++w) lse_merge(m, s, sm[w], ss[w]);
        part[(size_t)row * gridDim.x + blockIdx.x] = make_float2(m, s);
    }
}

// The same partial for bf16 logits whose rows are 16-byte aligned (the training head: class count padded to a multiple of 8):
// 8 logits per 16-byte load, the thread's running (m, s) updated once per 8 elements — local max first, then eight exp2 of
// differences — instead of a 2-byte load, a branch and an exp per element (30 -> ~10 us on the [256, 60 008] logits of c2).
__global__ void __launch_bounds__(kLossThreads) k7_lse_partial_bf16x8_kernel(const __nv_bfloat16 *__restrict__ x, int64_t stride, int V,
                                                                            int chunk, float2 *__restrict__ part) {
    const int row = blockIdx.y, c0 = blockIdx.x * chunk, c1 = min(V, c0 + chunk);        // chunk % 8 == 0
    const uint4 *xr = reinterpret_cast<const uint4 *>(x + (size_t)row * stride);
    constexpr float kL2e = 1.4426950408889634f;
    float m = -INFINITY, s = 0.f;                                                       // m in log2 units
    for (int c = c0 + threadIdx.x * 8; c < c1; c += kLossThreads * 8) {
        const uint4 v = __ldg(xr + (c >> 3));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            f[2 * e] = __uint_as_float(w[e] << 16) * kL2e;
            f[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u) * kL2e;
        }
        const int nv = min(8, c1 - c);
        float mx = -INFINITY;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (e >= nv) f[e] = -INFINITY;
            mx = fmaxf(mx, f[e]);
        }
        if (mx == -INFINITY) continue;
        const float mn = fmaxf(m, mx);
        float a = s * exp2f(m - mn);                                                     // m == -inf: s == 0
#pragma unroll
        for (int e = 0; e < 8; ++e) a += exp2f(f[e] - mn);
        s = a;
        m = mn;
    }
    m *= 0.6931471805599453f;                                                           // back to natural-log units for lse_merge
    __shared__ float sm[kLossThreads / 32], ss[kLossThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lse_merge(m, s, __shfl_xor_sync(0xffffffffu, m, o), __shfl_xor_sync(0xffffffffu, s, o));
    if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = m; ss[threadIdx.x >> 5] = s; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kLossThreads / 32; ++w) lse_merge(m, s, sm[w], ss[w]);
        part[(size_t)row * gridDim.x + blockIdx.x] = make_float2(m, s);
    }
}

// one CTA: lse[row] from the row's partials (fixed order), row loss = lse - x[target]; loss = sum / #valid rows
template <typename T>
__global__ void __launch_bounds__(kLossThreads) k7_nll_finish_kernel(const T *__restrict__ x, int64_t stride, int B, int V, int nchunk,
                                                                    const float2 *__restrict__ part,
                                                                    const int64_t *__restrict__ target, int64_t ignore_index,
                                                                    float *__restrict__ lse, float *__restrict__ loss) {
    float acc = 0.f, cnt = 0.f;
    for (int row = threadIdx.x; row < B; row += kLossThreads) {
        float m = -INFINITY, s = 0.f;
        for (int c = 0; c < nchunk; ++c) {
            const float2 p = part[(size_t)row * nchunk + c];
            lse_merge(m, s, p.x, p.y);
        }
        const float l = m + logf(s);
        lse[row] = l;
        const int64_t t = target[row];
        if (t != ignore_index && t >= 0 && t < V) {
            acc += l - ldf(x + (size_t)row * stride + t);
            cnt += 1.f;
        }
    }
    __shared__ float sa[kLossThreads], sc[kLossThreads];
    sa[threadIdx.x] = acc;
    sc[threadIdx.x] = cnt;
    __syncthreads();
    for (int o = kLossThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        loss[0] = sa[0] / sc[0];          // all rows ignored: 0 / 0 = NaN, as torch's mean reduction
        loss[1] = sc[0];
    }
}

// backward: d logits[r][c] = ( softmax[r][c] - [c == target_r] ) * g / #valid   (0 for ignored rows)
template <typename T>
__global__ void __launch_bounds__(kLossThreads) k7_nll_bwd_kernel(const T *__restrict__ x, int64_t stride, int V,
                                                                 const int64_t *__restrict__ target, int64_t ignore_index,
                                                                 const float *__restrict__ lse, const float *__restrict__ aux,
                                                                 const float *__restrict__ gout, T *__restrict__ dx, int64_t dstride,
                                                                 int Vpad) {
    const int row = blockIdx.y;
    const int64_t t = target[row];
    const bool live = t != ignore_index && t >= 0 && t < V;
    const float scale = live ? __ldg(gout) / __ldg(aux + 1) : 0.f;
    const float l = lse[row];
    const T *xr = x + (size_t)row * stride;
    T *dr = dx + (size_t)row * dstride;
    const int c0 = blockIdx.x * (kLossThreads * 8);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * kLossThreads + threadIdx.x;
        if (c < V) {
            const float p = live ? __expf(ldf(xr + c) - l) : 0.f;
            stf(dr + c, (p - (c == (int)t ? 1.f : 0.f)) * scale);
        } else if (c < Vpad) {
            stf(dr + c, 0.f);          // padding classes of a row (see mobgt.h): no gradient
        }
    }
}

// the same for bf16 logits / gradients with 16-byte aligned rows (both strides multiples of 8): 8 classes per 16-byte load / store
__global__ void __launch_bounds__(kLossThreads) k7_nll_bwd_bf16x8_kernel(const __nv_bfloat16 *__restrict__ x, int64_t stride, int V,
                                                                        const int64_t *__restrict__ target, int64_t ignore_index,
                                                                        const float *__restrict__ lse, const float *__restrict__ aux,
                                                                        const float *__restrict__ gout, __nv_bfloat16 *__restrict__ dx,
                                                                        int64_t dstride, int Vpad) {
    const int row = blockIdx.y;
    const int64_t t = target[row];
    const bool live = t != ignore_index && t >= 0 && t < V;
    const float scale = live ? __ldg(gout) / __ldg(aux + 1) : 0.f;
    constexpr float kL2e = 1.4426950408889634f;
    const float l2 = lse[row] * kL2e;
    const int c = (blockIdx.x * kLossThreads + threadIdx.x) * 8;
    if (c >= Vpad) return;
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(x + (size_t)row * stride + c));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float g[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cc = c + 2 * e + h;
            const float xv = __uint_as_float(h ? (w[e] & 0xFFFF0000u) : (w[e] << 16));
            const float p = (live && cc < V) ? exp2f(fmaf(xv, kL2e, -l2)) : 0.f;      // padding classes (cc >= V): no gradient
            g[h] = (p - ((live && cc == (int)t) ? 1.f : 0.f)) * scale;
        }
        __nv_bfloat162 pk = __floats2bfloat162_rn(g[0], g[1]);
        o[e] = *reinterpret_cast<uint32_t *>(&pk);
    }
    *reinterpret_cast<uint4 *>(dx + (size_t)row * dstride + c) = make_uint4(o[0], o[1], o[2], o[3]);
}

// ---- GradientTailLoss (model_fqandtoyo.py:545-550, k = 1, beta = 1):
//   f(x) = -alpha (1 - p) log p   on the target class,   -p log(1 - p)   elsewhere,   p = sigmoid(x) ;  loss = mean over [B,V]
// (the reference evaluates both branches everywhere and multiplies by the one-hot mask; only the selected branch is evaluated
// here, which differs only where the other branch would be 0 * inf)
__device__ __forceinline__ float gtl_f(float x, bool hot, float alpha) {
    const float p = 1.f / (1.f + __expf(-x));
    return hot ? -alpha * (1.f - p) * logf(p) : -p * logf(1.f - p);
}
// d f / d x with dp/dx = p (1 - p)
__device__ __forceinline__ float gtl_df(float x, bool hot, float alpha) {
    const float p = 1.f / (1.f + __expf(-x));
    const float q = 1.f - p;
    return hot ? alpha * q * (p * logf(p) - q) : p * (p - q * logf(q));
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads) k7_gtl_partial_kernel(const T *__restrict__ x, int64_t stride, int V, int chunk,
                                                                     const int64_t *__restrict__ target, float alpha,
                                                                     float *__restrict__ part) {
    const int row = blockIdx.y, c0 = blockIdx.x * chunk, c1 = min(V, c0 + chunk);
    const int t = (int)target[row];
    const T *xr = x + (size_t)row * stride;
    float acc = 0.f;
    for (int c = c0 + threadIdx.x; c < c1; c += kLossThreads) acc += gtl_f(ldf(xr + c), c == t, alpha);
    __shared__ float sa[kLossThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sa[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kLossThreads / 32; ++w) acc += sa[w];
        part[(size_t)row * gridDim.x + blockIdx.x] = acc;
    }
}

__global__ void __launch_bounds__(kLossThreads) k7_sum_finish_kernel(const float *__restrict__ part, int n, float inv_count,
                                                                    float *__restrict__ loss) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += kLossThreads) acc += part[i];
    __shared__ float sa[kLossThreads];
    sa[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kLossThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sa[threadIdx.x] += sa[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = sa[0] * inv_count;
}

template <typename T>
__global__ void __launch_bounds__(kLossThreads) k7_gtl_bwd_kernel(const T *__restrict__ x, int64_t stride, int V,
                                                                 const int64_t *__restrict__ target, float alpha, float inv_count,
                                                                 const float *__restrict__ gout, T *__restrict__ dx, int64_t dstride,
                                                                 int Vpad) {
    const int row = blockIdx.y;
    const int t = (int)target[row];
    const float scale = __ldg(gout) * inv_count;
    const T *xr = x + (size_t)row * stride;
    T *dr = dx + (size_t)row * dstride;
    const int c0 = blockIdx.x * (kLossThreads * 8);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int c = c0 + u * kLossThreads + threadIdx.x;
        if (c < V) stf(dr + c, gtl_df(ldf(xr + c), c == t, alpha) * scale);
        else if (c < Vpad) stf(dr + c, 0.f);
    }
}

// chunks per row: enough CTAs to fill the 148 SMs a few times over, each chunk at least 2 048 columns
static inline int loss_chunks(int B, int V) {
    int want = ceil_div(4 * kNumSMs, B > 0 ? B : 1);
    int maxc = ceil_div(V, 2048);
    int c = want < 1 ? 1 : want;
    if (c > maxc) c = maxc;
    return c < 1 ? 1 : c;
}

}  // namespace mobgt

using namespace mobgt;

extern "C" int64_t mobgt_loss_workspace_bytes(int32_t B, int32_t V) {
    if (B < 1 || V < 1) return -1;
    return (int64_t)B * loss_chunks(B, V) * (int64_t)sizeof(float2);
}

#define MOBGT_LOSS_COMMON(name)                                                                                       \
    MOBGT_REQUIRE(dtype == MOBGT_F32 || dtype == MOBGT_BF16, MOBGT_ERR_BAD_DTYPE, name ": logits dtype %d", dtype);   \
    MOBGT_REQUIRE(B >= 1 && V >= 1 && row_stride >= V, MOBGT_ERR_BAD_SHAPE, name ": B=%d V=%d row_stride=%lld", B, V, \
                  (long long)row_stride);                                                                             \
    MOBGT_REQUIRE(B <= 65535, MOBGT_ERR_BAD_SHAPE, name ": B=%d > 65535 rows", B)

extern "C" int32_t mobgt_lsm_nll_fwd(const void *logits, int32_t dtype, int64_t row_stride, const int64_t *target,
                                     int64_t ignore_index, int32_t B, int32_t V, void *workspace, int64_t workspace_bytes,
                                     float *lse, float *loss, void *stream) {
    MOBGT_REQUIRE(logits && target && workspace && lse && loss, MOBGT_ERR_NULL, "mobgt_lsm_nll_fwd: null pointer");
    MOBGT_LOSS_COMMON("mobgt_lsm_nll_fwd");
    const int nchunk = loss_chunks(B, V);
    MOBGT_REQUIRE(workspace_bytes >= (int64_t)B * nchunk * (int64_t)sizeof(float2), MOBGT_ERR_WORKSPACE_TOO_SMALL,
                  "mobgt_lsm_nll_fwd: workspace %lld bytes", (long long)workspace_bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunk = ceil_div(V, nchunk);
    float2 *part = static_cast<float2 *>(workspace);
    dim3 grid((unsigned)nchunk, (unsigned)B);
    if (dtype == MOBGT_F32) {
        k7_lse_partial_kernel<float><<<grid, kLossThreads, 0, s>>>(static_cast<const float *>(logits), row_stride, V, chunk, part);
        MOBGT_LAUNCH_OK("k7_lse_partial_kernel");
        k7_nll_finish_kernel<float><<<1, kLossThreads, 0, s>>>(static_cast<const float *>(logits), row_stride, B, V, nchunk, part,
                                                              target, ignore_index, lse, loss);
    } else {
        if (row_stride % 8 == 0 && ((uintptr_t)logits & 15) == 0) {
            const int chunk8 = round_up(chunk, 8);          // (covers V with the same nchunk: chunk8 >= chunk)
            k7_lse_partial_bf16x8_kernel<<<grid, kLossThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(logits), row_stride, V, chunk8, part);
        } else {
            k7_lse_partial_kernel<__nv_bfloat16><<<grid, kLossThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(logits), row_stride, V,
                                                                              chunk, part);
        }
        MOBGT_LAUNCH_OK("k7_lse_partial_kernel");
        k7_nll_finish_kernel<__nv_bfloat16><<<1, kLossThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(logits), row_stride, B, V,
                                                                      nchunk, part, target, ignore_index, lse, loss);
    }
    MOBGT_LAUNCH_OK("k7_nll_finish_kernel");
    return MOBGT_OK;
}

extern "C" int32_t mobgt_lsm_nll_bwd(const void *logits, int32_t dtype, int64_t row_stride, const int64_t *target,
                                     int64_t ignore_index, int32_t B, int32_t V, const float *lse, const float *loss,
                                     const float *grad_out, void *dlogits, int64_t d_row_stride, void *stream) {
    MOBGT_REQUIRE(logits && target && lse && loss && grad_out && dlogits, MOBGT_ERR_NULL, "mobgt_lsm_nll_bwd: null pointer");
    MOBGT_LOSS_COMMON("mobgt_lsm_nll_bwd");
    MOBGT_REQUIRE(d_row_stride >= V, MOBGT_ERR_BAD_SHAPE, "mobgt_lsm_nll_bwd: d_row_stride=%lld", (long long)d_row_stride);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // a d_row_stride slightly larger than V (classes padded to a multiple of 8 / 16 for the GEMMs): the padding is zero-filled
    const int Vpad = (d_row_stride - V < 16) ? (int)d_row_stride : V;
    dim3 grid((unsigned)ceil_div(Vpad, kLossThreads * 8), (unsigned)B);
    if (dtype == MOBGT_F32)
        k7_nll_bwd_kernel<float><<<grid, kLossThreads, 0, s>>>(static_cast<const float *>(logits), row_stride, V, target, ignore_index,
                                                              lse, loss, grad_out, static_cast<float *>(dlogits), d_row_stride, Vpad);
    else if (row_stride % 8 == 0 && d_row_stride % 8 == 0 && Vpad % 8 == 0 && (((uintptr_t)logits | (uintptr_t)dlogits) & 15) == 0)
        k7_nll_bwd_bf16x8_kernel<<<dim3((unsigned)ceil_div(Vpad, kLossThreads * 8), (unsigned)B), kLossThreads, 0, s>>>(
            static_cast<const __nv_bfloat16 *>(logits), row_stride, V, target, ignore_index, lse, loss, grad_out,
            static_cast<__nv_bfloat16 *>(dlogits), d_row_stride, Vpad);
    else
        k7_nll_bwd_kernel<__nv_bfloat16><<<grid, kLossThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(logits), row_stride, V, target,
                                                                      ignore_index, lse, loss, grad_out,
                                                                      static_cast<__nv_bfloat16 *>(dlogits), d_row_stride, Vpad);
    MOBGT_LAUNCH_OK("k7_nll_bwd_kernel");
    return MOBGT_OK;
}

extern "C" int32_t mobgt_gtl_fwd(const void *logits, int32_t dtype, int64_t row_stride, const int64_t *target, float alpha,
                                 int32_t B, int32_t V, void *workspace, int64_t workspace_bytes, float *loss, void *stream) {
    MOBGT_REQUIRE(logits && target && workspace && loss, MOBGT_ERR_NULL, "mobgt_gtl_fwd: null pointer");
    MOBGT_LOSS_COMMON("mobgt_gtl_fwd");
    const int nchunk = loss_chunks(B, V);
    MOBGT_REQUIRE(workspace_bytes >= (int64_t)B * nchunk * (int64_t)sizeof(float), MOBGT_ERR_WORKSPACE_TOO_SMALL,
                  "mobgt_gtl_fwd: workspace %lld bytes", (long long)workspace_bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int chunk = ceil_div(V, nchunk);
    float *part = static_cast<float *>(workspace);
    dim3 grid((unsigned)nchunk, (unsigned)B);
    if (dtype == MOBGT_F32)
        k7_gtl_partial_kernel<float><<<grid, kLossThreads, 0, s>>>(static_cast<const float *>(logits), row_stride, V, chunk, target,
                                                                  alpha, part);
    else
        k7_gtl_partial_kernel<__nv_bfloat16><<<grid, kLossThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(logits), row_stride, V,
                                                                          chunk, target, alpha, part);
    MOBGT_LAUNCH_OK("k7_gtl_partial_kernel");
    k7_sum_finish_kernel<<<1, kLossThreads, 0, s>>>(part, B * nchunk, 1.0f / ((float)B * (float)V), loss);
    MOBGT_LAUNCH_OK("k7_sum_finish_kernel");
    return MOBGT_OK;
}

extern "C" int32_t mobgt_gtl_bwd(const void *logits, int32_t dtype, int64_t row_stride, const int64_t *target, float alpha,
                                 int32_t B, int32_t V, const float *grad_out, void *dlogits, int64_t d_row_stride,
                                 void *stream) {
    MOBGT_REQUIRE(logits && target && grad_out && dlogits, MOBGT_ERR_NULL, "mobgt_gtl_bwd: null pointer");
    MOBGT_LOSS_COMMON("mobgt_gtl_bwd");
    MOBGT_REQUIRE(d_row_stride >= V, MOBGT_ERR_BAD_SHAPE, "mobgt_gtl_bwd: d_row_stride=%lld", (long long)d_row_stride);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int Vpad = (d_row_stride - V < 16) ? (int)d_row_stride : V;
    dim3 grid((unsigned)ceil_div(Vpad, kLossThreads * 8), (unsigned)B);
    const float inv = 1.0f / ((float)B * (float)V);
    if (dtype == MOBGT_F32)
        k7_gtl_bwd_kernel<float><<<grid, kLossThreads, 0, s>>>(static_cast<const float *>(logits), row_stride, V, target, alpha, inv,
                                                              grad_out, static_cast<float *>(dlogits), d_row_stride, Vpad);
    else
        k7_gtl_bwd_kernel<__nv_bfloat16><<<grid, kLossThreads, 0, s>>>(static_cast<const __nv_bfloat16 *>(logits), row_stride, V, target,
                                                                      alpha, inv, grad_out, static_cast<__nv_bfloat16 *>(dlogits),
                                                                      d_row_stride, Vpad);
    MOBGT_LAUNCH_OK("k7_gtl_bwd_kernel");
    return MOBGT_OK;
}
