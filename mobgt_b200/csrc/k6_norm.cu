// K6 — encoder row ops next to the attention path (first slice): LayerNorm forward / backward and the
// column sum that produces the bias gradient of every Linear (sm_100a, HBM-bound streaming kernels).
//
// Replace, inside EncoderLayer.forward (model_fqandtoyo.py:1731-1743) and the final LN (:1360-1364):
//   * nn.LayerNorm forward  — torch: vectorized_layer_norm_kernel (53 us at [33 024, 192])
//   * its backward          — torch: GammaBetaBackwardCUDAKernelTemplate + layer_norm_grad_input_kernel (252 + 27 us; 18.5 % of
//                             the training step's kernel time, profiles/r01z_launches_summary.txt)
//   * grad_bias = dy.sum(0) — torch: reduce_kernel<bf16> (66 us per Linear; 9.5 % of the step)
// One warp per row, the row lives in registers (D / 32 values per lane), statistics by warp shuffles; dgamma / dbeta and the
// column sums are accumulated per lane across all rows a warp owns, folded per CTA in shared memory and reduced over the CTAs
// in a fixed order by a second tiny kernel (deterministic, no atomics).
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace mobgt {

constexpr int kNormWarps = 8;        // warps (rows in flight) per CTA
constexpr int kNormCtas = 6 * kNumSMs;      // forward: 6 CTAs x 8 warps resident per SM (40 registers / thread)
constexpr int kNormCtasBwd = 4 * kNumSMs;   // backward: 4 CTAs per SM (64 registers / thread)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// counter-based dropout mask: element `idx` of call `seed` is kept iff hash >= p * 2^32 (lowbias32 mix); the backward
// regenerates the mask from (seed, idx) instead of storing it
__device__ __forceinline__ bool drop_keep(uint32_t idx, uint32_t seed_lo, uint32_t seed_hi, uint32_t thresh) {
    uint32_t z = idx + seed_lo;
    z ^= z >> 16; z *= 0x7feb352du;
    z ^= z >> 15; z *= 0x846ca68bu;
    z ^= z >> 16; z ^= seed_hi;
    z *= 0x9E3779B1u; z ^= z >> 15;
    return z >= thresh;
}

// A step counter kept in device memory (optional) is folded into the host seed, so that a CUDA graph that replays these
// kernels with baked-in arguments still draws a fresh mask every replay (the counter is incremented inside the graph).
__device__ __forceinline__ void mix_device_seed(const unsigned long long *seed_dev, uint32_t &lo, uint32_t &hi) {
    if (seed_dev != nullptr) {
        const unsigned long long sd = *seed_dev;
        lo ^= (uint32_t)sd * 0x9E3779B1u;
        hi += (uint32_t)(sd >> 32) ^ ((uint32_t)sd * 0x85EBCA6Bu);
    }
}

// out = LayerNorm(s),  s = x + dropout(y)   (y == NULL: s = x).  s is written when s_out != NULL (the new residual stream).
template <int PER>
__global__ void __launch_bounds__(kNormWarps * 32) k6_layernorm_fwd_kernel(const float *__restrict__ x, const __nv_bfloat16 *__restrict__ y,
                                                                          uint32_t drop_thresh, float drop_scale, uint32_t seed_lo,
                                                                          uint32_t seed_hi, const unsigned long long *__restrict__ seed_dev,
                                                                          const float *__restrict__ gamma,
                                                                          const float *__restrict__ beta, float eps, int N,
                                                                          float *__restrict__ s_out, float *__restrict__ out,
                                                                          __nv_bfloat16 *__restrict__ out16,
                                                                          float *__restrict__ mean, float *__restrict__ rstd) {
    constexpr int D = PER * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    mix_device_seed(seed_dev, seed_lo, seed_hi);
    float g[PER], b[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        g[i] = gamma[i * 32 + lane];
        b[i] = beta[i * 32 + lane];
    }
    for (int row = blockIdx.x * kNormWarps + warp; row < N; row += gridDim.x * kNormWarps) {
        const size_t base = (size_t)row * D;
        float v[PER];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            v[i] = x[base + i * 32 + lane];
            if (y != nullptr) {
                const float yv = __bfloat162float(y[base + i * 32 + lane]);
                if (drop_thresh == 0u) v[i] += yv;
                else if (drop_keep((uint32_t)(base + i * 32 + lane), seed_lo, seed_hi, drop_thresh))
                    v[i] += __bfloat162float(__float2bfloat16_rn(yv * drop_scale));      // bf16 dropout output, as nn.Dropout on a bf16 tensor
            }
            s += v[i];
        }
        if (s_out != nullptr) {
#pragma unroll
            for (int i = 0; i < PER; ++i) s_out[base + i * 32 + lane] = v[i];
        }
        const float mu = warp_sum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const float d = v[i] - mu;
            q += d * d;
        }
        const float rs = rsqrtf(warp_sum(q) * (1.0f / D) + eps);      // biased variance, as torch.nn.LayerNorm
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const float o = (v[i] - mu) * rs * g[i] + b[i];
            if (out != nullptr) out[base + i * 32 + lane] = o;
            if (out16 != nullptr) out16[base + i * 32 + lane] = __float2bfloat16_rn(o);
        }
        if (lane == 0) {
            mean[row] = mu;
            rstd[row] = rs;
        }
    }
}

// ds = rstd * ( g dy - mean(g dy) - xhat mean(g dy xhat) ) + ds_ext ;  dgamma = sum_rows dy xhat ;  dbeta = sum_rows dy
// (x = the LayerNorm input s).  dx = ds (residual branch) and, when dyb_out != NULL, dyb_out = bf16(ds * mask * scale) is the
// gradient of the sub-layer output that went through the dropout.
template <int PER>
__global__ void __launch_bounds__(kNormWarps * 32) k6_layernorm_bwd_kernel(const float *__restrict__ dy, const __nv_bfloat16 *__restrict__ dy16,
                                                                          const float *__restrict__ ds_ext, const float *__restrict__ x,
                                                                          const float *__restrict__ gamma, const float *__restrict__ mean,
                                                                          const float *__restrict__ rstd, int N, uint32_t drop_thresh,
                                                                          float drop_scale, uint32_t seed_lo, uint32_t seed_hi,
                                                                          const unsigned long long *__restrict__ seed_dev,
                                                                          float *__restrict__ dx, __nv_bfloat16 *__restrict__ dyb_out,
                                                                          float *__restrict__ partial) {
    constexpr int D = PER * 32;
    __shared__ float sred[kNormWarps][3 * D];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    mix_device_seed(seed_dev, seed_lo, seed_hi);
    float g[PER], dg[PER], db[PER], dc[PER];     // dc: column sums of dyb_out = the bias gradient of the Linear that produced y
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        g[i] = gamma[i * 32 + lane];
        dg[i] = 0.f;
        db[i] = 0.f;
        dc[i] = 0.f;
    }
    for (int row = blockIdx.x * kNormWarps + warp; row < N; row += gridDim.x * kNormWarps) {
        const size_t base = (size_t)row * D;
        const float mu = mean[row], rs = rstd[row];
        float xh[PER], gy[PER];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            // the fp32 consumer (residual stream) and the bf16 consumer (next GEMM) of the output both send a gradient
            float d = dy != nullptr ? dy[base + i * 32 + lane] : 0.f;
            if (dy16 != nullptr) d += __bfloat162float(dy16[base + i * 32 + lane]);
            xh[i] = (x[base + i * 32 + lane] - mu) * rs;
            gy[i] = d * g[i];
            s1 += gy[i];
            s2 += gy[i] * xh[i];
            dg[i] += d * xh[i];
            db[i] += d;
        }
        const float m1 = warp_sum(s1) * (1.0f / D), m2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            float ds = rs * (gy[i] - m1 - xh[i] * m2);
            if (ds_ext != nullptr) ds += ds_ext[base + i * 32 + lane];
            dx[base + i * 32 + lane] = ds;
            if (dyb_out != nullptr) {
                const bool keep = drop_thresh == 0u || drop_keep((uint32_t)(base + i * 32 + lane), seed_lo, seed_hi, drop_thresh);
                const __nv_bfloat16 yb = __float2bfloat16_rn(keep ? ds * (drop_thresh == 0u ? 1.0f : drop_scale) : 0.f);
                dyb_out[base + i * 32 + lane] = yb;
                dc[i] += __bfloat162float(yb);        // the sum of what a column-sum kernel would read back
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        sred[warp][i * 32 + lane] = dg[i];
        sred[warp][D + i * 32 + lane] = db[i];
        sred[warp][2 * D + i * 32 + lane] = dc[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 3 * D; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kNormWarps; ++w) s += sred[w][c];
        partial[(size_t)blockIdx.x * 3 * D + c] = s;
    }
}

// out[c] = sum over parts of partial[part][c]: block = 32 columns x 8 rows, row r sums the parts p = r (mod 8) in order, the
// 8 row sums are folded in order (fixed summation tree: deterministic)
__global__ void __launch_bounds__(256) k6_reduce_parts_kernel(const float *__restrict__ partial, int nparts, int C,
                                                              float *__restrict__ out0, int C0, float *__restrict__ out1,
                                                              int C1 = 1 << 30, float *__restrict__ out2 = nullptr) {
    __shared__ float sfold[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (c < C)
        for (int p = ty; p < nparts; p += 8) s += partial[(size_t)p * C + c];
    sfold[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < C) {
        float t = sfold[0][tx];
#pragma unroll
        for (int r = 1; r < 8; ++r) t += sfold[r][tx];
        if (c < C0) out0[c] = t;
        else if (c < C0 + C1) out1[c - C0] = t;
        else if (out2 != nullptr) out2[c - C0 - C1] = t;
    }
}

// column sums of a row-major [N, C] matrix (bf16 or f32, row stride in elements): thread = 8 (bf16) / 4 (f32) adjacent columns
// (one 16-byte load per row), CTA = (column group, row strip); per-CTA strip sums -> partial[strip][C]
template <typename T>
__global__ void __launch_bounds__(256) k6_colsum_kernel(const T *__restrict__ src, int64_t stride, int N, int C, int rows_per_strip,
                                                        float *__restrict__ partial) {
    constexpr int V = 16 / (int)sizeof(T);
    __shared__ float sacc[8][32 * V];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + lane) * V;
    const int r0 = blockIdx.y * rows_per_strip, r1 = min(N, r0 + rows_per_strip);
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    if (col < C) {
        // 4 rows (8 warps apart) in flight per iteration: the loads are independent, the adds keep a fixed order
        for (int r = r0 + warp; r < r1; r += 32) {
            uint4 w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                w[j] = (r + 8 * j < r1) ? __ldg(reinterpret_cast<const uint4 *>(src + (size_t)(r + 8 * j) * stride + col))
                                        : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t u[4] = {w[j].x, w[j].y, w[j].z, w[j].w};
                if constexpr (sizeof(T) == 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] += __uint_as_float(u[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        acc[2 * i] += __uint_as_float(u[i] << 16);
                        acc[2 * i + 1] += __uint_as_float(u[i] & 0xFFFF0000u);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) sacc[warp][lane * V + i] = acc[i];
    __syncthreads();
    for (int c = threadIdx.x; c < 32 * V; c += blockDim.x) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sacc[w][c];
        const int gc = blockIdx.x * 32 * V + c;
        if (gc < C) partial[(size_t)blockIdx.y * C + gc] = s;
    }
}

// ---- GELU of the FFN (nn.GELU(), exact erf form: model_fqandtoyo.py:1650): backward on bf16 [N, C] activations, fp32 math.
// (The forward stays the library's elementwise kernel: a hand-written one measured slower, 48 vs 39 us at [33024, 1024].)
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = __expf(-0.5f * x * x) * 0.39894228040143267794f;
    return fmaf(x, pdf, cdf);
}

// dH = dA * gelu'(h) (bf16) fused with the column sums of dH — the bias gradient of the Linear that produced h — in the
// layout of k6_colsum_kernel: thread = 8 adjacent columns, CTA = (256 columns, row strip), fixed summation order.
__global__ void __launch_bounds__(256) k6_gelu_bwd_colsum_kernel(const __nv_bfloat16 *__restrict__ dA, const __nv_bfloat16 *__restrict__ h,
                                                                 __nv_bfloat16 *__restrict__ dH, int N, int C, int rows_per_strip,
                                                                 float *__restrict__ partial) {
    __shared__ float sacc[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int col = (blockIdx.x * 32 + lane) * 8;
    const int r0 = blockIdx.y * rows_per_strip, r1 = min(N, r0 + rows_per_strip);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (col < C) {
        for (int r = r0 + warp; r < r1; r += 16) {   // 2 rows (8 warps apart) in flight per iteration
            uint4 g[2], x[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const bool ok = r + 8 * j < r1;
                const size_t o = (size_t)(r + 8 * j) * C + col;
                g[j] = ok ? __ldg(reinterpret_cast<const uint4 *>(dA + o)) : make_uint4(0u, 0u, 0u, 0u);
                x[j] = ok ? __ldg(reinterpret_cast<const uint4 *>(h + o)) : make_uint4(0u, 0u, 0u, 0u);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (r + 8 * j >= r1) continue;
                const uint32_t gu[4] = {g[j].x, g[j].y, g[j].z, g[j].w}, xu[4] = {x[j].x, x[j].y, x[j].z, x[j].w};
                uint32_t o4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float d0 = __uint_as_float(gu[k] << 16) * gelu_grad_f(__uint_as_float(xu[k] << 16));
                    const float d1 = __uint_as_float(gu[k] & 0xFFFF0000u) * gelu_grad_f(__uint_as_float(xu[k] & 0xFFFF0000u));
                    o4[k] = sm100::pack_bf16(d0, d1);
                    acc[2 * k] += __uint_as_float(o4[k] << 16);            // the sum runs over the bf16 values that are stored
                    acc[2 * k + 1] += __uint_as_float(o4[k] & 0xFFFF0000u);
                }
                *reinterpret_cast<uint4 *>(dH + (size_t)(r + 8 * j) * C + col) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sacc[warp][lane * 8 + i] = acc[i];
    __syncthreads();
    {
        const int c = threadIdx.x;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sacc[w][c];
        const int gc = blockIdx.x * 256 + c;
        if (gc < C) partial[(size_t)blockIdx.y * C + gc] = s;
    }
}

}  // namespace mobgt

using namespace mobgt;

#define MOBGT_NORM_DISPATCH(PERV, CALL)     \
    switch (PERV) {                         \
        case 2: { constexpr int P_ = 2; CALL; } break;   \
        case 4: { constexpr int P_ = 4; CALL; } break;   \
        case 6: { constexpr int P_ = 6; CALL; } break;   \
        case 8: { constexpr int P_ = 8; CALL; } break;   \
        case 10: { constexpr int P_ = 10; CALL; } break; \
        case 12: { constexpr int P_ = 12; CALL; } break; \
        case 16: { constexpr int P_ = 16; CALL; } break; \
        default: MOBGT_REQUIRE(false, MOBGT_ERR_UNSUPPORTED, "layernorm: D=%d is not built (D/32 in {2,4,6,8,10,12,16})", D); \
    }

extern "C" int64_t mobgt_layernorm_bwd_workspace_bytes(int32_t D) {
    if (D <= 0 || D % 32 != 0 || D > 512) return -1;
    return (int64_t)kNormCtas * 3 * D * (int64_t)sizeof(float);
}

static uint32_t drop_threshold(float p) {
    if (!(p > 0.f)) return 0u;
    const double t = (double)p * 4294967296.0;
    return t >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t;
}

extern "C" int32_t mobgt_add_dropout_layernorm_fwd(const float *x, const void *y_bf16, float drop_p, uint64_t seed, const float *gamma,
                                                   const float *beta, float eps, int32_t N, int32_t D, float *s_out, float *out,
                                                   void *out_bf16, float *mean, float *rstd, const void *seed_dev, void *stream) {
    MOBGT_REQUIRE(x && gamma && beta && (out || out_bf16) && mean && rstd, MOBGT_ERR_NULL, "mobgt_add_dropout_layernorm_fwd: null pointer");
    MOBGT_REQUIRE(D > 0 && D % 32 == 0 && D <= 512, MOBGT_ERR_BAD_SHAPE, "mobgt_add_dropout_layernorm_fwd: D=%d", D);
    MOBGT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, MOBGT_ERR_BAD_SHAPE, "mobgt_add_dropout_layernorm_fwd: p=%f", (double)drop_p);
    MOBGT_REQUIRE((int64_t)N * D < (1ll << 32), MOBGT_ERR_BAD_SHAPE, "mobgt_add_dropout_layernorm_fwd: N*D must be < 2^32");
    if (N <= 0) return MOBGT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = min(kNormCtas, ceil_div(N, kNormWarps));
    const uint32_t th = y_bf16 ? drop_threshold(drop_p) : 0u;
    const float sc = 1.0f / (1.0f - drop_p);
    MOBGT_NORM_DISPATCH(D / 32, (k6_layernorm_fwd_kernel<P_><<<grid, kNormWarps * 32, 0, s>>>(
                                    x, static_cast<const __nv_bfloat16 *>(y_bf16), th, sc, (uint32_t)seed, (uint32_t)(seed >> 32),
                                    static_cast<const unsigned long long *>(seed_dev), gamma, beta, eps, N, s_out, out,
                                    static_cast<__nv_bfloat16 *>(out_bf16), mean, rstd)));
    MOBGT_LAUNCH_OK("k6_layernorm_fwd_kernel");
    return MOBGT_OK;
}

extern "C" int32_t mobgt_layernorm_fwd(const float *x, const float *gamma, const float *beta, float eps, int32_t N, int32_t D,
                                       float *out, void *out_bf16, float *mean, float *rstd, void *stream) {
    return mobgt_add_dropout_layernorm_fwd(x, nullptr, 0.f, 0ull, gamma, beta, eps, N, D, nullptr, out, out_bf16, mean, rstd, nullptr,
                                           stream);
}

extern "C" int32_t mobgt_add_dropout_layernorm_bwd(const float *dy, const void *dy_bf16, const float *ds_ext, const float *s_saved,
                                                   const float *gamma, const float *mean, const float *rstd, int32_t N, int32_t D,
                                                   float drop_p, uint64_t seed, float *dx, void *dyb_out, float *dgamma, float *dbeta,
                                                   float *dyb_colsum, void *workspace, int64_t workspace_bytes, const void *seed_dev,
                                                   void *stream) {
    MOBGT_REQUIRE((dy || dy_bf16) && s_saved && gamma && mean && rstd && dx && dgamma && dbeta && workspace, MOBGT_ERR_NULL,
                  "mobgt_add_dropout_layernorm_bwd: null pointer");
    MOBGT_REQUIRE(D > 0 && D % 32 == 0 && D <= 512, MOBGT_ERR_BAD_SHAPE, "mobgt_add_dropout_layernorm_bwd: D=%d", D);
    MOBGT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, MOBGT_ERR_BAD_SHAPE, "mobgt_add_dropout_layernorm_bwd: p=%f", (double)drop_p);
    MOBGT_REQUIRE(workspace_bytes >= (int64_t)kNormCtas * 3 * D * 4, MOBGT_ERR_WORKSPACE_TOO_SMALL, "mobgt_add_dropout_layernorm_bwd: workspace");
    MOBGT_REQUIRE(dyb_colsum == nullptr || dyb_out != nullptr, MOBGT_ERR_NULL, "mobgt_add_dropout_layernorm_bwd: dyb_colsum needs dyb_out");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = N > 0 ? min(kNormCtasBwd, ceil_div(N, kNormWarps)) : 0;
    float *partial = static_cast<float *>(workspace);
    if (grid > 0) {
        MOBGT_NORM_DISPATCH(D / 32, (k6_layernorm_bwd_kernel<P_><<<grid, kNormWarps * 32, 0, s>>>(
                                        dy, static_cast<const __nv_bfloat16 *>(dy_bf16), ds_ext, s_saved, gamma, mean, rstd, N,
                                        drop_threshold(drop_p), 1.0f / (1.0f - drop_p), (uint32_t)seed, (uint32_t)(seed >> 32),
                                        static_cast<const unsigned long long *>(seed_dev), dx, static_cast<__nv_bfloat16 *>(dyb_out),
                                        partial)));
        MOBGT_LAUNCH_OK("k6_layernorm_bwd_kernel");
    }
    k6_reduce_parts_kernel<<<ceil_div(3 * D, 32), 256, 0, s>>>(partial, grid, 3 * D, dgamma, D, dbeta, D, dyb_colsum);
    MOBGT_LAUNCH_OK("k6_reduce_parts_kernel");
    return MOBGT_OK;
}

extern "C" int32_t mobgt_layernorm_bwd(const float *dy, const void *dy_bf16, const float *x, const float *gamma, const float *mean,
                                       const float *rstd, int32_t N, int32_t D, float *dx, float *dgamma, float *dbeta, void *workspace,
                                       int64_t workspace_bytes, void *stream) {
    return mobgt_add_dropout_layernorm_bwd(dy, dy_bf16, nullptr, x, gamma, mean, rstd, N, D, 0.f, 0ull, dx, nullptr, dgamma, dbeta,
                                           nullptr, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int64_t mobgt_colsum_workspace_bytes(int32_t N, int32_t C) {
    if (N < 0 || C <= 0) return -1;
    const int strips = max(1, min(256, ceil_div(N, 128)));
    return (int64_t)strips * C * (int64_t)sizeof(float);
}

extern "C" int32_t mobgt_colsum(const void *src, int32_t src_dtype, int64_t src_stride, int32_t N, int32_t C, float *out,
                                void *workspace, int64_t workspace_bytes, void *stream) {
    MOBGT_REQUIRE(src && out && workspace, MOBGT_ERR_NULL, "mobgt_colsum: null pointer");
    MOBGT_REQUIRE(src_dtype == MOBGT_F32 || src_dtype == MOBGT_BF16, MOBGT_ERR_BAD_DTYPE, "mobgt_colsum: dtype");
    const int V = src_dtype == MOBGT_F32 ? 4 : 8;
    MOBGT_REQUIRE(C > 0 && C % V == 0 && src_stride % V == 0 && ((uintptr_t)src & 15) == 0, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_colsum: C=%d and the row stride must be multiples of %d elements, src 16-byte aligned", C, V);
    const int strips = max(1, min(256, ceil_div(N, 128)));
    MOBGT_REQUIRE(workspace_bytes >= (int64_t)strips * C * 4, MOBGT_ERR_WORKSPACE_TOO_SMALL, "mobgt_colsum: workspace");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
    const int rows_per_strip = max(1, ceil_div(max(N, 1), strips));
    dim3 grid((unsigned)ceil_div(C, 32 * V), (unsigned)strips);
    if (src_dtype == MOBGT_F32)
        k6_colsum_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float *>(src), src_stride, N, C, rows_per_strip, partial);
    else
        k6_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(src), src_stride, N, C, rows_per_strip,
                                                             partial);
    MOBGT_LAUNCH_OK("k6_colsum_kernel");
    k6_reduce_parts_kernel<<<ceil_div(C, 32), 256, 0, s>>>(partial, strips, C, out, C, out);
    MOBGT_LAUNCH_OK("k6_reduce_parts_kernel");
    return MOBGT_OK;
}

extern "C" int32_t mobgt_gelu_bwd_colsum(const void *da_bf16, const void *h_bf16, int32_t N, int32_t C, void *dh_bf16, float *dbias,
                                         void *workspace, int64_t workspace_bytes, void *stream) {
    MOBGT_REQUIRE(da_bf16 && h_bf16 && dh_bf16 && dbias && workspace, MOBGT_ERR_NULL, "mobgt_gelu_bwd_colsum: null pointer");
    MOBGT_REQUIRE(N >= 0 && C > 0 && C % 8 == 0, MOBGT_ERR_BAD_SHAPE, "mobgt_gelu_bwd_colsum: N=%d C=%d (C must be a multiple of 8)", N, C);
    const int strips = max(1, min(256, ceil_div(N, 128)));
    MOBGT_REQUIRE(workspace_bytes >= (int64_t)strips * C * 4, MOBGT_ERR_WORKSPACE_TOO_SMALL, "mobgt_gelu_bwd_colsum: workspace");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *partial = static_cast<float *>(workspace);
    const int rows_per_strip = max(1, ceil_div(max(N, 1), strips));
    dim3 grid((unsigned)ceil_div(C, 256), (unsigned)strips);
    k6_gelu_bwd_colsum_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(da_bf16), static_cast<const __nv_bfloat16 *>(h_bf16),
                                                   static_cast<__nv_bfloat16 *>(dh_bf16), N, C, rows_per_strip, partial);
    MOBGT_LAUNCH_OK("k6_gelu_bwd_colsum_kernel");
    k6_reduce_parts_kernel<<<ceil_div(C, 32), 256, 0, s>>>(partial, strips, C, dbias, C, dbias);
    MOBGT_LAUNCH_OK("k6_reduce_parts_kernel");
    return MOBGT_OK;
}
