// K2 — attention-bias build, forward and backward (sm_100a, HBM-bound).
//
// Replaces model_fqandtoyo.py:1143-1216 (fp32 statement: model.py:126-190): the
// rel_pos_encoder / poi_pos_encoder gathers, the graph-token virtual distance, and the multi-hop
// edge encoding (edge_encoder gather along each path, per-hop edge_dis_encoder 8x8 dot, sum over
// hops divided by the clamped path length) fused into ONE pass that reads the compact packed
// indices (i16 rel_pos, i16 poi_pos, u8 edge_input[hops]) and writes bias[B,H,T,Tp].
//
//   bias[g,h,a,b] = [a>=1, b>=1] ( R[rp[g,a-1,b-1],h] + Ppos[pp[g,a-1,b-1],h] + Edge[g,h,a-1,b-1] )
//                 + [a>=1, b==0] t[h]            (row 0 gets no bias: model_fqandtoyo.py:1160-1165)
//   Edge[g,h,i,j] = ( sum_k sum_h' E[ei[g,i,j,k],h'] * W[k,h',h] ) / sp,   sp = clamp(M,1,hops) with M = rp-1
//   (and -inf where M >= rel_pos_max, collator.py:354-358; padding columns are masked inside the
//    attention kernel from seqlens, so the 2*attn_bias term of :1144/:1216 needs no storage).
//
// The 8x8 dot is hoisted out of the pair loop: EW[k][v][h] = sum_h' E[v,h'] W[k,h',h] is a
// [hops,128,H] table (80 KB fp32 at hops=20,H=8) built by a prologue kernel and staged in shared
// memory, so the per-pair work is `walk length` 32-byte shared-memory gathers.  E[0] is a padding row
// (nn.Embedding(padding_idx=0), model_fqandtoyo.py:784) and is structurally zero, so the hop loop stops
// at the first ei == 0.
//
// Backward: dBias (fp32, summed over layers) -> dR, dPpos, dt and dEW by shared-memory-privatised
// histograms (persistent CTAs), then dE = dEW . W^T and dW = E^T . dEW in an epilogue kernel.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mobgt {

constexpr int kH = 8;          // heads (one float4 pair per table row); other head counts: UNSUPPORTED
constexpr int kEdgeVocab = 128;  // edge_encoder rows (model_fqandtoyo.py:784)

struct K2Common {
    const int32_t *n;        // [B]
    const int64_t *sq_off;   // [B+1]
    const int16_t *rel_pos;  // packed [sum n^2]  (M+1)
    const int16_t *poi_pos;  // packed [sum n^2]
    const uint8_t *edge_in;  // packed [sum n^2, hops]  (e+1)
    int B, T, Tp, hops, rel_pos_max;
};

// EW[k][v][h] = sum_h' E[v][h'] * W[k][h'][h]
__global__ void k2_prep_kernel(const float *__restrict__ E, const float *__restrict__ W, int hops, float *__restrict__ EW) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hops * kEdgeVocab * kH) return;
    const int h = idx % kH, v = (idx / kH) % kEdgeVocab, k = idx / (kH * kEdgeVocab);
    float acc = 0.f;
#pragma unroll
    for (int hp = 0; hp < kH; ++hp) acc += E[v * kH + hp] * W[(k * kH + hp) * kH + h];
    EW[idx] = acc;
}

__device__ __forceinline__ void add8(float (&a)[8], const float *__restrict__ p) {
    const float4 x = *reinterpret_cast<const float4 *>(p);
    const float4 y = *reinterpret_cast<const float4 *>(p + 4);
    a[0] += x.x; a[1] += x.y; a[2] += x.z; a[3] += x.w;
    a[4] += y.x; a[5] += y.y; a[6] += y.z; a[7] += y.w;
}

template <typename OutT>
__device__ __forceinline__ void store_out(OutT *p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16 *p, float v) { *p = __float2bfloat16_rn(v); }

// persistent CTAs (the EW table is staged once per CTA); one thread per (a, b) token pair, 8 heads each.
template <typename OutT>
__global__ void __launch_bounds__(256) k2_bias_fwd_kernel(const K2Common c, const float *__restrict__ R,
                                                          const float *__restrict__ Pp, const float *__restrict__ tvd,
                                                          const float *__restrict__ EWg, OutT *__restrict__ out) {
    extern __shared__ __align__(16) float EW[];  // [hops][128][8]
    const int tabn = c.hops * kEdgeVocab * kH;
    for (int i = threadIdx.x * 4; i < tabn; i += blockDim.x * 4)
        *reinterpret_cast<float4 *>(EW + i) = *reinterpret_cast<const float4 *>(EWg + i);
    __syncthreads();
    const int tiles = ceil_div(c.T * c.T, (int)blockDim.x);
    const size_t hs = (size_t)c.T * c.Tp;
    for (int w = blockIdx.x; w < tiles * c.B; w += gridDim.x) {
        const int g = w / tiles, tile = w - g * tiles;
        const int n = c.n[g];
        const int Tg = n + 1;
        const int cell = tile * blockDim.x + threadIdx.x;
        if (cell >= Tg * Tg) continue;
        const int a = cell / Tg, b = cell - a * Tg;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (a >= 1 && b == 0) {
#pragma unroll
            for (int h = 0; h < 8; ++h) acc[h] = tvd[h];
        } else if (a >= 1) {
            const int64_t pc = c.sq_off[g] + (int64_t)(a - 1) * n + (b - 1);
            const int rp = c.rel_pos[pc];
            const int M = rp - 1;
            if (M >= c.rel_pos_max) {
#pragma unroll
                for (int h = 0; h < 8; ++h) acc[h] = -INFINITY;
            } else {
                const uint8_t *ei = c.edge_in + pc * c.hops;
                for (int k = 0; k < c.hops; ++k) {
                    const int v = ei[k];
                    if (v == 0) break;
                    add8(acc, EW + ((size_t)k * kEdgeVocab + v) * kH);
                }
                const float inv = 1.0f / (float)min(max(M, 1), c.hops);
                const int pp = c.poi_pos[pc];
                const float4 r0 = *reinterpret_cast<const float4 *>(R + rp * kH), r1 = *reinterpret_cast<const float4 *>(R + rp * kH + 4);
                const float4 p0 = *reinterpret_cast<const float4 *>(Pp + pp * kH), p1 = *reinterpret_cast<const float4 *>(Pp + pp * kH + 4);
                const float rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
                const float qq[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
                for (int h = 0; h < 8; ++h) acc[h] = (rr[h] + qq[h]) + acc[h] * inv;
            }
        }
        OutT *o = out + ((size_t)g * kH * c.T + a) * c.Tp + b;
#pragma unroll
        for (int h = 0; h < 8; ++h) store_out<OutT>(o + h * hs, acc[h]);
    }
}

// ---- backward --------------------------------------------------------------------------------------
// Histogram add with warp-level pre-aggregation: the keys of one stream (one hop slot, or rel_pos / poi_pos) are
// heavily duplicated inside a warp (most walks carry the same edge feature), and same-address shared-memory atomics
// serialise.  Up to two dominant key groups are peeled with a warp reduction (one atomic per head for the whole group);
// whatever is left goes through individual atomics.  Must be called by all 32 lanes (key < 0 = nothing to add).
__device__ __forceinline__ void warp_hist_add(float *hist, int key, const float (&w)[8]) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned act = __ballot_sync(full, key >= 0);
#pragma unroll 1
    for (int peel = 0; peel < 2 && act; ++peel) {
        const int l = __ffs(act) - 1;
        const int kk = __shfl_sync(full, key, l);
        const unsigned grp = __ballot_sync(full, key == kk);
        if (__popc(grp) < 3) break;
        const bool in = key == kk;
#pragma unroll
        for (int h = 0; h < 8; ++h) {
            float s = in ? w[h] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(full, s, o);
            if (lane == l) atomicAdd(hist + kk * 8 + h, s);
        }
        if (in) key = -1;
        act &= ~grp;
    }
    if (key >= 0) {
#pragma unroll
        for (int h = 0; h < 8; ++h) atomicAdd(hist + key * 8 + h, w[h]);
    }
}

// persistent CTAs; shared-memory histograms dEW[hops][128][8], dR[512][8], dP[bins][8], dt[8]
__global__ void __launch_bounds__(256) k2_bias_bwd_kernel(const K2Common c, const float *__restrict__ dB, int num_bins,
                                                          float *__restrict__ dEW, float *__restrict__ dR,
                                                          float *__restrict__ dP, float *__restrict__ dt) {
    extern __shared__ __align__(16) float sm[];
    const int nEW = c.hops * kEdgeVocab * kH, nR = 512 * kH, nP = num_bins * kH;
    float *sEW = sm, *sR = sEW + nEW, *sP = sR + nR, *st = sP + nP;
    const int tot = nEW + nR + nP + kH;
    for (int i = threadIdx.x; i < tot; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int tiles = ceil_div(c.T * c.T, (int)blockDim.x);
    const size_t hs = (size_t)c.T * c.Tp;
    const unsigned full = 0xffffffffu;
    for (int w = blockIdx.x; w < tiles * c.B; w += gridDim.x) {
        const int g = w / tiles, tile = w - g * tiles;
        const int n = c.n[g], Tg = n + 1;
        const int cell = tile * blockDim.x + threadIdx.x;
        const int a = cell / Tg, b = cell - a * Tg;
        const bool in_graph = cell < Tg * Tg && a >= 1;
        float d[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (in_graph) {
            const float *src = dB + ((size_t)g * kH * c.T + a) * c.Tp + b;
#pragma unroll
            for (int h = 0; h < 8; ++h) d[h] = src[h * hs];
        }
        // column 0 of rows >= 1: graph-token virtual distance
        warp_hist_add(st, (in_graph && b == 0) ? 0 : -1, d);
        bool pair = in_graph && b >= 1;
        int rp = 0, pp = 0, M = 0;
        int64_t pc = 0;
        if (pair) {
            pc = c.sq_off[g] + (int64_t)(a - 1) * n + (b - 1);
            rp = c.rel_pos[pc];
            M = rp - 1;
            if (M >= c.rel_pos_max) pair = false;     // -inf entries carry no gradient
            else pp = c.poi_pos[pc];
        }
        warp_hist_add(sR, pair ? rp : -1, d);
        warp_hist_add(sP, pair ? pp : -1, d);
        const float inv = 1.0f / (float)min(max(M, 1), c.hops);
#pragma unroll
        for (int h = 0; h < 8; ++h) d[h] *= inv;
        const uint8_t *ei = c.edge_in + pc * c.hops;
        for (int k = 0; k < c.hops; ++k) {
            int v = 0;
            if (pair) {
                v = ei[k];
                if (v == 0) pair = false;              // end of the walk
            }
            if (__ballot_sync(full, pair) == 0) break;
            warp_hist_add(sEW, pair ? k * kEdgeVocab + v : -1, d);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nEW; i += blockDim.x) if (sEW[i] != 0.f) atomicAdd(dEW + i, sEW[i]);
    for (int i = threadIdx.x; i < nR; i += blockDim.x) if (sR[i] != 0.f) atomicAdd(dR + i, sR[i]);
    for (int i = threadIdx.x; i < nP; i += blockDim.x) if (sP[i] != 0.f) atomicAdd(dP + i, sP[i]);
    if (threadIdx.x < kH && st[threadIdx.x] != 0.f) atomicAdd(dt + threadIdx.x, st[threadIdx.x]);
}

// dE[v][h'] = sum_k sum_h dEW[k][v][h] W[k][h'][h] ;  dW[k][h'][h] = sum_v E[v][h'] dEW[k][v][h]
__global__ void k2_bias_bwd_finish_kernel(const float *__restrict__ dEW, const float *__restrict__ E,
                                          const float *__restrict__ W, int hops, float *__restrict__ dE,
                                          float *__restrict__ dW) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nE = kEdgeVocab * kH, nW = hops * kH * kH;
    if (idx < nE) {
        const int v = idx / kH, hp = idx % kH;
        float acc = 0.f;
        for (int k = 0; k < hops; ++k)
#pragma unroll
            for (int h = 0; h < kH; ++h) acc += dEW[((size_t)k * kEdgeVocab + v) * kH + h] * W[(k * kH + hp) * kH + h];
        dE[idx] = (v == 0) ? 0.f : acc;   // padding_idx row
    } else if (idx < nE + nW) {
        const int j = idx - nE;
        const int h = j % kH, hp = (j / kH) % kH, k = j / (kH * kH);
        float acc = 0.f;
        for (int v = 0; v < kEdgeVocab; ++v) acc += E[v * kH + hp] * dEW[((size_t)k * kEdgeVocab + v) * kH + h];
        dW[j] = acc;
    }
}

}  // namespace mobgt

using namespace mobgt;

extern "C" int32_t mobgt_bias_fwd(const int32_t *n, const int64_t *sq_off, const int16_t *rel_pos, const int16_t *poi_pos,
                                  const uint8_t *edge_in, int32_t B, int32_t T, int32_t Tp, int32_t hops, int32_t H,
                                  int32_t rel_pos_max, const float *R, const float *Ppos, const float *E, const float *W,
                                  const float *tvd, void *workspace, void *out, int32_t out_dtype, void *stream) {
    MOBGT_REQUIRE(n && sq_off && rel_pos && poi_pos && edge_in && R && Ppos && E && W && tvd && workspace && out,
                  MOBGT_ERR_NULL, "mobgt_bias_fwd: null pointer");
    MOBGT_REQUIRE(H == kH, MOBGT_ERR_UNSUPPORTED, "mobgt_bias_fwd: num_heads=%d (only 8 is built)", H);
    MOBGT_REQUIRE(hops >= 1 && hops <= MOBGT_MAX_HOPS, MOBGT_ERR_BAD_SHAPE, "mobgt_bias_fwd: hops=%d", hops);
    MOBGT_REQUIRE(T >= 2 && Tp >= T && Tp % 8 == 0, MOBGT_ERR_BAD_SHAPE, "mobgt_bias_fwd: T=%d Tp=%d", T, Tp);
    MOBGT_REQUIRE(out_dtype == MOBGT_F32 || out_dtype == MOBGT_BF16, MOBGT_ERR_BAD_DTYPE, "mobgt_bias_fwd: dtype");
    if (B <= 0) return MOBGT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *EW = static_cast<float *>(workspace);
    const int tabn = hops * kEdgeVocab * kH;
    k2_prep_kernel<<<ceil_div(tabn, 256), 256, 0, s>>>(E, W, hops, EW);
    MOBGT_LAUNCH_OK("k2_prep_kernel");
    K2Common c{n, sq_off, rel_pos, poi_pos, edge_in, B, T, Tp, hops, rel_pos_max};
    const size_t smem = (size_t)tabn * sizeof(float);
    const int tiles_total = ceil_div(T * T, 256) * B;
    dim3 grid((unsigned)min(2 * kNumSMs, tiles_total));
    if (out_dtype == MOBGT_F32) {
        MOBGT_CUDA_OK(cudaFuncSetAttribute(k2_bias_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k2_bias_fwd_kernel<float><<<grid, 256, smem, s>>>(c, R, Ppos, tvd, EW, static_cast<float *>(out));
    } else {
        MOBGT_CUDA_OK(cudaFuncSetAttribute(k2_bias_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k2_bias_fwd_kernel<__nv_bfloat16><<<grid, 256, smem, s>>>(c, R, Ppos, tvd, EW, static_cast<__nv_bfloat16 *>(out));
    }
    MOBGT_LAUNCH_OK("k2_bias_fwd_kernel");
    return MOBGT_OK;
}

extern "C" int32_t mobgt_bias_bwd(const int32_t *n, const int64_t *sq_off, const int16_t *rel_pos, const int16_t *poi_pos,
                                  const uint8_t *edge_in, int32_t B, int32_t T, int32_t Tp, int32_t hops, int32_t H,
                                  int32_t rel_pos_max, int32_t num_bins, const float *dBias, const float *E, const float *W,
                                  void *workspace, float *dR, float *dPpos, float *dE, float *dW, float *dtvd, void *stream) {
    MOBGT_REQUIRE(n && sq_off && rel_pos && poi_pos && edge_in && dBias && E && W && workspace && dR && dPpos && dE && dW && dtvd,
                  MOBGT_ERR_NULL, "mobgt_bias_bwd: null pointer");
    MOBGT_REQUIRE(H == kH, MOBGT_ERR_UNSUPPORTED, "mobgt_bias_bwd: num_heads=%d (only 8 is built)", H);
    MOBGT_REQUIRE(hops >= 1 && hops <= MOBGT_MAX_HOPS && num_bins >= 1 && num_bins <= 1024, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_bias_bwd: hops=%d num_bins=%d", hops, num_bins);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *dEW = static_cast<float *>(workspace);
    const int nEW = hops * kEdgeVocab * kH;
    MOBGT_CUDA_OK(cudaMemsetAsync(dEW, 0, (size_t)nEW * 4, s));
    MOBGT_CUDA_OK(cudaMemsetAsync(dR, 0, 512 * kH * 4, s));
    MOBGT_CUDA_OK(cudaMemsetAsync(dPpos, 0, (size_t)num_bins * kH * 4, s));
    MOBGT_CUDA_OK(cudaMemsetAsync(dtvd, 0, kH * 4, s));
    if (B > 0) {
        K2Common c{n, sq_off, rel_pos, poi_pos, edge_in, B, T, Tp, hops, rel_pos_max};
        const size_t smem = (size_t)(nEW + 512 * kH + num_bins * kH + kH) * sizeof(float);
        MOBGT_CUDA_OK(cudaFuncSetAttribute(k2_bias_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k2_bias_bwd_kernel<<<2 * kNumSMs, 256, smem, s>>>(c, dBias, num_bins, dEW, dR, dPpos, dtvd);
        MOBGT_LAUNCH_OK("k2_bias_bwd_kernel");
    }
    k2_bias_bwd_finish_kernel<<<ceil_div(kEdgeVocab * kH + hops * kH * kH, 256), 256, 0, s>>>(dEW, E, W, hops, dE, dW);
    MOBGT_LAUNCH_OK("k2_bias_bwd_finish_kernel");
    return MOBGT_OK;
}
