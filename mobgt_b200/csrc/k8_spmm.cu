// K8 — CSR SpMM for the global GCN tables (sm_100a, L2-gather bound).  SURVEY.md §8(f) #1.
//
// Replaces `torch.spmm(adj, support)` of GraphConvolution.forward (modelGNN.py:39-46) as used by the two global GCNs the
// model recomputes every forward (`poi_distance_model(X, D_A)`, `poi_cat_model(C_X, C_A)`, model_fqandtoyo.py:1236-1237;
// GCN.forward modelGNN.py:65-73).  The reference multiplies a DENSE P x P row-normalised adjacency; the adjacency is 0.03 %
// (c2 world) to 13 % dense, so the product is a CSR gather:
//
//     Y[r, :] = act( sum_{j in row r} val[j] * S[col[j], :] + bias )        act = LeakyReLU(slope) or identity
//
// One sub-group of LPR = D/4 lanes per output row (D = 16 / 32 / 64 / 128 -> 4 / 8 / 16 / 32 lanes, each owning one float4 of
// the row), 32/LPR rows per warp.  The sub-group reads its row's (col, val) pairs LPR at a time with one coalesced load and
// broadcasts them by shuffle, so LPR independent 16-byte gathers of S are in flight per lane; S (<= 30 MB) is L2-resident.
// The bias add and the activation are fused into the store.  The backward is the same kernel on the CSR of A^T (built once:
// the adjacency is a constant of the dataset), so there are no atomics and the result is bitwise reproducible.
#include "common.cuh"

namespace mobgt {

template <int LPR>
__global__ void __launch_bounds__(256) k8_spmm_kernel(const int32_t *__restrict__ crow, const int32_t *__restrict__ col,
                                                      const float *__restrict__ val, int nrows, const float *__restrict__ S,
                                                      const float *__restrict__ bias, float slope, int act,
                                                      float *__restrict__ Y) {
    constexpr int D = 4 * LPR, RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sub = lane % LPR;
    const int row = warp * RPW + lane / LPR;
    const bool live = row < nrows;
    const int j0 = live ? __ldg(crow + row) : 0, j1 = live ? __ldg(crow + row + 1) : 0;
    // all sub-groups of a warp iterate together (shuffles need converged lanes): loop to the longest row of the warp
    int len = j1 - j0;
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 *S4 = reinterpret_cast<const float4 *>(S);
    for (int base = 0; base < len; base += LPR) {
        const int j = j0 + base + sub;
        const bool ok = j < j1;
        const int c = ok ? __ldg(col + j) : 0;
        const float v = ok ? __ldg(val + j) : 0.f;
        float4 s[LPR > 8 ? 8 : LPR];
        constexpr int STEP = LPR > 8 ? 8 : LPR;
#pragma unroll
        for (int t0 = 0; t0 < LPR; t0 += STEP) {
            float vv[STEP];
#pragma unroll
            for (int t = 0; t < STEP; ++t) {
                const int cc = __shfl_sync(0xffffffffu, c, t0 + t, LPR);
                vv[t] = __shfl_sync(0xffffffffu, v, t0 + t, LPR);
                s[t] = __ldg(S4 + (size_t)cc * LPR + sub);            // vv == 0 past the row end: row 0 is read, adds nothing
            }
#pragma unroll
            for (int t = 0; t < STEP; ++t) {
                acc.x = fmaf(vv[t], s[t].x, acc.x);
                acc.y = fmaf(vv[t], s[t].y, acc.y);
                acc.z = fmaf(vv[t], s[t].z, acc.z);
                acc.w = fmaf(vv[t], s[t].w, acc.w);
            }
        }
    }
    if (!live) return;
    if (bias != nullptr) {
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias) + sub);
        acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
    }
    if (act) {
        acc.x = acc.x > 0.f ? acc.x : acc.x * slope;
        acc.y = acc.y > 0.f ? acc.y : acc.y * slope;
        acc.z = acc.z > 0.f ? acc.z : acc.z * slope;
        acc.w = acc.w > 0.f ? acc.w : acc.w * slope;
    }
    reinterpret_cast<float4 *>(Y)[(size_t)row * LPR + sub] = acc;
    (void)D;
}

}  // namespace mobgt

using namespace mobgt;

extern "C" int32_t mobgt_spmm_csr(const int32_t *crow, const int32_t *col, const float *val, int32_t nrows, const float *S,
                                  int32_t D, const float *bias, float leaky_slope, int32_t activation, float *Y,
                                  void *stream) {
    MOBGT_REQUIRE(crow && col && val && S && Y, MOBGT_ERR_NULL, "mobgt_spmm_csr: null pointer");
    MOBGT_REQUIRE(nrows >= 0, MOBGT_ERR_BAD_SHAPE, "mobgt_spmm_csr: nrows=%d", nrows);
    MOBGT_REQUIRE(D == 16 || D == 32 || D == 64 || D == 128, MOBGT_ERR_UNSUPPORTED,
                  "mobgt_spmm_csr: width %d (built for 16 / 32 / 64 / 128: the GCN widths of MobGT)", D);
    MOBGT_REQUIRE((((uintptr_t)S | (uintptr_t)Y | (uintptr_t)bias) & 15) == 0, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_spmm_csr: S, Y and bias must be 16-byte aligned");
    if (nrows == 0) return MOBGT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int lpr = D / 4, rpw = 32 / lpr;
    const int warps = ceil_div(nrows, rpw);
    const int blocks = ceil_div(warps, 8);
    switch (lpr) {
        case 4: k8_spmm_kernel<4><<<blocks, 256, 0, s>>>(crow, col, val, nrows, S, bias, leaky_slope, activation, Y); break;
        case 8: k8_spmm_kernel<8><<<blocks, 256, 0, s>>>(crow, col, val, nrows, S, bias, leaky_slope, activation, Y); break;
        case 16: k8_spmm_kernel<16><<<blocks, 256, 0, s>>>(crow, col, val, nrows, S, bias, leaky_slope, activation, Y); break;
        default: k8_spmm_kernel<32><<<blocks, 256, 0, s>>>(crow, col, val, nrows, S, bias, leaky_slope, activation, Y); break;
    }
    MOBGT_LAUNCH_OK("k8_spmm_kernel");
    return MOBGT_OK;
}
