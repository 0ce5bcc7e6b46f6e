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

constexpr int kDomEdge = 4;    // edge count 1 -> attn_edge_type 3 (wrapper.py:49-53) -> +1 (collator.py:86-93): "one transition"
constexpr int kRelRows = 512;  // rel_pos_encoder rows (model_fqandtoyo.py:786): keys 0..511

// walk length the distance key rp = M + 1 predicts: min(M, hops) for a reachable off-diagonal pair, else 0
__host__ __device__ __forceinline__ int expected_walk(int rp, int hops) {
    const int M = rp - 1;
    return (M < 510) ? min(max(M, 0), hops) : 0;
}

// Prologue tables (workspace):
//   EW[k][v][h] = sum_h' E[v][h'] * W[k][h'][h]                                            [hops][128][8]
//   RL[rp][h]   = R[rp][h] + ( sum_{k < L(rp)} EW[k][dom][h] ) / clamp(rp-1, 1, hops)       [512][8]
// RL is the whole rel_pos-keyed part of the bias of a pair whose walk is the EXPECTED one: L(rp) hops that all carry the
// dominant edge feature (a single transition).  -inf where rp-1 >= rel_pos_max (collator.py:354-358).
__global__ void k2_prep_kernel(const float *__restrict__ E, const float *__restrict__ W, const float *__restrict__ R, int hops,
                               int rel_pos_max, float *__restrict__ EW, float *__restrict__ RL) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int nEW = hops * kEdgeVocab * kH;
    if (idx < nEW) {
        const int h = idx % kH, v = (idx / kH) % kEdgeVocab, k = idx / (kH * kEdgeVocab);
        float acc = 0.f;
#pragma unroll
        for (int hp = 0; hp < kH; ++hp) acc += E[v * kH + hp] * W[(k * kH + hp) * kH + h];
        EW[idx] = acc;
    } else if (idx < nEW + kRelRows * kH) {
        const int j = idx - nEW;
        const int h = j % kH, rp = j / kH;
        const int L = expected_walk(rp, hops);
        float ps = 0.f;
        for (int k = 0; k < L; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int hp = 0; hp < kH; ++hp) acc += E[kDomEdge * kH + hp] * W[(k * kH + hp) * kH + h];
            ps += acc;
        }
        const int M = rp - 1;
        const float inv = 1.0f / (float)min(max(M, 1), hops);
        RL[j] = (M >= rel_pos_max) ? -INFINITY : R[j] + ps * inv;
    }
}

__device__ __forceinline__ void add8(float (&a)[8], const float *__restrict__ p) {
    const float4 x = *reinterpret_cast<const float4 *>(p);
    const float4 y = *reinterpret_cast<const float4 *>(p + 4);
    a[0] += x.x; a[1] += x.y; a[2] += x.z; a[3] += x.w;
    a[4] += y.x; a[5] += y.y; a[6] += y.z; a[7] += y.w;
}
__device__ __forceinline__ void sub8(float (&a)[8], const float *__restrict__ p) {
    const float4 x = *reinterpret_cast<const float4 *>(p);
    const float4 y = *reinterpret_cast<const float4 *>(p + 4);
    a[0] -= x.x; a[1] -= x.y; a[2] -= x.z; a[3] -= x.w;
    a[4] -= y.x; a[5] -= y.y; a[6] -= y.z; a[7] -= y.w;
}

__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }

template <typename OutT>
__device__ __forceinline__ void store2(OutT *p, float a, float b);
template <>
__device__ __forceinline__ void store2<float>(float *p, float a, float b) { *reinterpret_cast<float2 *>(p) = make_float2(a, b); }
template <>
__device__ __forceinline__ void store2<__nv_bfloat16>(__nv_bfloat16 *p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162 *>(p) = __floats2bfloat162_rn(a, b);
}
template <typename OutT>
__device__ __forceinline__ void store1(OutT *p, float a);
template <>
__device__ __forceinline__ void store1<float>(float *p, float a) { *p = a; }
template <>
__device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16 *p, float a) { *p = __float2bfloat16_rn(a); }

// expected walk bytes of a pair with walk length L: word q = dominant feature in bytes [4q, 4q+4) ∩ [0, L)
__device__ __forceinline__ uint32_t expected_word(int L, int q) {
    const int r = min(max(L - 4 * q, 0), 4);
    const uint32_t mask = r >= 4 ? 0xFFFFFFFFu : ((1u << (8 * r)) - 1u);
    return (0x01010101u * (uint32_t)kDomEdge) & mask;
}

// Forward.  Persistent CTAs; RL [512][8], Ppos [bins][8] and the expected-walk words XW [hops+1][8] are staged once per
// CTA in shared memory.  One thread per PAIR OF ADJACENT CELLS (a, b) (a, b+1), b even, of the Tp-pitched plane row, so
// every head plane is written as 4-byte (bf16x2) / 8-byte (f32x2) words and a warp covers 64 consecutive columns.  All
// index bytes of a thread's two cells (2 x (2 + 2 + hops) B) are fetched up front as independent loads.
//
// Per cell the work is TWO table gathers: bias = RL[rp] + Ppos[pp].  The walk bytes are only compared (XOR) against the
// walk the distance predicts; the bytes that differ — another edge feature (0.7 % of the hops of trajectory graphs) or a
// walk cut short by the reference's node-0 quirk (algos.pyx:57-62) — are corrected one by one from the EW table in
// global memory (L1-resident):   Edge = sum_k EW[k][e_k]  =  PS[L] + sum_{k : e_k != x_k} ( EW[k][e_k] - EW[k][x_k] ).
// The identity holds for ANY byte pattern (EW[k][0] = 0: edge_encoder row 0 is the padding row, model_fqandtoyo.py:784).
template <typename OutT, int HOPW>
__global__ void __launch_bounds__(512, 2) k2_bias_fwd_kernel(const K2Common c, const float *__restrict__ RLg,
                                                             const float *__restrict__ Pg, int num_bins,
                                                             const float *__restrict__ tvd,
                                                             const float *__restrict__ EWg, OutT *__restrict__ out) {
    extern __shared__ __align__(16) float sm[];
    const int nR = kRelRows * kH, nP = num_bins * kH;
    float *RL = sm, *Pp = RL + nR;
    uint32_t *XW = reinterpret_cast<uint32_t *>(Pp + nP);          // [hops + 1][8]
    for (int i = threadIdx.x * 4; i < nR; i += blockDim.x * 4)
        *reinterpret_cast<float4 *>(RL + i) = *reinterpret_cast<const float4 *>(RLg + i);
    for (int i = threadIdx.x * 4; i < nP; i += blockDim.x * 4)
        *reinterpret_cast<float4 *>(Pp + i) = *reinterpret_cast<const float4 *>(Pg + i);
    constexpr int hopw = HOPW;                    // hops = 4 * HOPW
    for (int i = threadIdx.x; i < (c.hops + 1) * 8; i += blockDim.x) XW[i] = (i & 7) < hopw ? expected_word(i >> 3, i & 7) : 0u;
    float tv[8];
#pragma unroll
    for (int h = 0; h < 8; ++h) tv[h] = tvd[h];
    __syncthreads();
    const int half = c.Tp >> 1;                   // cell pairs per plane row
    const int per_graph = c.T * half;
    const int tiles = ceil_div(per_graph, (int)blockDim.x);
    const size_t hs = (size_t)c.T * c.Tp;
    for (int w = blockIdx.x; w < tiles * c.B; w += gridDim.x) {
        const int g = w / tiles, tile = w - g * tiles;
        const int n = c.n[g];
        const int Tg = n + 1;
        const int f = tile * blockDim.x + threadIdx.x;
        const int a = f / half, b = (f - a * half) * 2;
        if (a >= Tg || b >= Tg) continue;
        const bool two = b + 1 < Tg;
        float acc[2][8];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int h = 0; h < 8; ++h) acc[s][h] = 0.f;
        if (a >= 1) {
            const int64_t rowp = c.sq_off[g] + (int64_t)(a - 1) * n;   // pair (a-1, j) lives at rowp + j
            // cell s of this thread is the pair j = b - 1 + s  (b == 0, s == 0: the virtual-distance column)
            int rp[2], pp[2];
            uint32_t ew[2][HOPW];
            bool is_pair[2];
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                const int j = b - 1 + s;
                is_pair[s] = j >= 0 && j < n;
                rp[s] = 0; pp[s] = 0;
#pragma unroll
                for (int q = 0; q < HOPW; ++q) ew[s][q] = 0u;
                if (is_pair[s]) {
                    const int64_t pc = rowp + j;
                    rp[s] = c.rel_pos[pc];
                    pp[s] = c.poi_pos[pc];
                    const uint32_t *ei = reinterpret_cast<const uint32_t *>(c.edge_in + pc * c.hops);
#pragma unroll
                    for (int q = 0; q < HOPW; ++q) ew[s][q] = __ldg(ei + q);
                }
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (!is_pair[s]) {
                    if (b == 0 && s == 0) {
#pragma unroll
                        for (int h = 0; h < 8; ++h) acc[0][h] = tv[h];
                    }
                    continue;
                }
                const int rk = min(max(rp[s], 0), kRelRows - 1);
                const int L = expected_walk(rk, c.hops);
                const uint4 x0 = *reinterpret_cast<const uint4 *>(XW + L * 8), x1 = *reinterpret_cast<const uint4 *>(XW + L * 8 + 4);
                const uint32_t xw[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                uint32_t any = 0u;
#pragma unroll
                for (int q = 0; q < HOPW; ++q) any |= ew[s][q] ^ xw[q];
                const float *r = RL + rk * kH, *q_ = Pp + min(max(pp[s], 0), num_bins - 1) * kH;
                const float4 r0 = *reinterpret_cast<const float4 *>(r), r1 = *reinterpret_cast<const float4 *>(r + 4);
                const float4 p0 = *reinterpret_cast<const float4 *>(q_), p1 = *reinterpret_cast<const float4 *>(q_ + 4);
                acc[s][0] = r0.x + p0.x; acc[s][1] = r0.y + p0.y; acc[s][2] = r0.z + p0.z; acc[s][3] = r0.w + p0.w;
                acc[s][4] = r1.x + p1.x; acc[s][5] = r1.y + p1.y; acc[s][6] = r1.z + p1.z; acc[s][7] = r1.w + p1.w;
                if (any != 0u) {
                    float corr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int q = 0; q < HOPW; ++q) {
                        {
                            uint32_t d = ew[s][q] ^ xw[q];
                            while (d != 0u) {
                                const int e = (__ffs(d) - 1) >> 3;
                                const int v = (ew[s][q] >> (8 * e)) & 0xFF, x = (xw[q] >> (8 * e)) & 0xFF;
                                const float *row = EWg + (size_t)(q * 4 + e) * kEdgeVocab * kH;
                                add8(corr, row + min(v, kEdgeVocab - 1) * kH);
                                sub8(corr, row + x * kH);
                                d &= ~(0xFFu << (8 * e));
                            }
                        }
                    }
                    const float inv = 1.0f / (float)min(max(rk - 1, 1), c.hops);
#pragma unroll
                    for (int h = 0; h < 8; ++h) acc[s][h] += corr[h] * inv;
                }
            }
        }
        OutT *o = out + ((size_t)g * kH * c.T + a) * c.Tp + b;
        if (two) {
#pragma unroll
            for (int h = 0; h < 8; ++h) store2<OutT>(o + h * hs, acc[0][h], acc[1][h]);
        } else {
#pragma unroll
            for (int h = 0; h < 8; ++h) store1<OutT>(o + h * hs, acc[0][h]);
        }
    }
}

// ---- backward --------------------------------------------------------------------------------------
// dBias (fp32, summed over layers) -> table gradients.  Every cell contributes its 8-vector d[h] to
//   dt (column 0), dR[rp], dPpos[pp] and dEW[k][ei[k]] * 1/sp for every hop k of its walk:
// 22 keyed adds whose keys are heavily duplicated (one distance / one edge feature dominates), and shared-memory fp32
// atomicAdd is a compare-and-swap loop (ATOMS.CAST.SPIN) that serialises on equal addresses.  So the main path uses
// NO atomics:
//   * lanes = (4 cells) x (8 heads): lane (p4, h) owns head h of cell p4 of the current group, so one keyed add is a
//     plain LDS / FADD / STS of 8 consecutive floats per cell, into WARP-PRIVATE histograms;
//   * the poi_pos and walk histograms are replicated per p4 (no two lanes ever share an address); the rel_pos histogram
//     is updated in 4 lock-step turns (one p4 per turn);
//   * the 20 hop adds of a walk collapse into ONE: the cell is added to A[L] (L = walk length) as if every hop carried
//     the dominant edge feature (kDomEdge: "one transition", by far the most common); only the hops with another
//     feature are visited: N[k] += d (warp-private) and dEW[k][v] += d (shared-memory atomic into the CTA's dEW
//     histogram, low contention), and the finish resolves dEW[k][dom] += sum_{L > k} A[L] - N[k].
// Per-CTA totals go to a partial buffer in the workspace and are reduced in a fixed order by the finish kernels
// (reproducible except for the rare-path atomics).

struct K2BwdPlan {
    int Rrows;      // rel_pos histogram rows: keys 0..Rrows-2 direct, key 511 -> row Rrows-1
    int nEW, nR, nP, nA, nN, stride;   // floats; partial row = [EW | R | P | A | N | t(8)]
    int warps;
};

__host__ __device__ inline K2BwdPlan k2_bwd_plan(int T, int hops, int num_bins) {
    K2BwdPlan p;
    p.Rrows = min(T, 511) + 1;
    p.nEW = hops * kEdgeVocab * kH;
    p.nR = p.Rrows * kH;
    p.nP = num_bins * kH;
    p.nA = (hops + 1) * kH;
    p.nN = hops * kH;
    p.stride = p.nEW + p.nR + p.nP + p.nA + p.nN + kH;
    const int per_warp = (p.nR + 4 * (p.nP + p.nA + p.nN) + kH) * 4;
    int w = (int)((227 * 1024 - 1024 - p.nEW * 4) / per_warp);
    p.warps = w > 8 ? 8 : w;
    return p;
}

// DT = float: one fp32 dBias buffer (nlayers == 1).  DT = bf16: nlayers per-layer dS planes (layer_stride elements apart,
// written by mobgt_attn_bwd mode 2), summed here in fp32.
template <typename DT>
__global__ void __launch_bounds__(256, 1) k2_bias_bwd_kernel(const K2Common c, const DT *__restrict__ dB, int nlayers,
                                                             int64_t layer_stride, int num_bins,
                                                             const K2BwdPlan pl, float *__restrict__ partial,
                                                             float *__restrict__ dR_overflow) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int p4 = lane >> 3, h = lane & 7;
    float *sEW = sm;
    const int per_warp = pl.nR + 4 * (pl.nP + pl.nA + pl.nN) + kH;
    float *sR = sm + pl.nEW + warp * per_warp;
    float *sP = sR + pl.nR + p4 * pl.nP;
    float *sA = sR + pl.nR + 4 * pl.nP + p4 * pl.nA;
    float *sN = sR + pl.nR + 4 * (pl.nP + pl.nA) + p4 * pl.nN;
    float *sT = sR + pl.nR + 4 * (pl.nP + pl.nA + pl.nN);
    for (int i = threadIdx.x; i < pl.nEW + nwarp * per_warp; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const unsigned full = 0xffffffffu;
    const size_t hs = (size_t)c.T * c.Tp;
    const int hopw = c.hops >> 2;
    const int gbase = lane & ~7;
    float tacc = 0.f;
    for (int u = blockIdx.x * nwarp + warp; u < c.B * c.T; u += gridDim.x * nwarp) {
        const int g = u / c.T, a = u - g * c.T;
        const int n = c.n[g];
        if (a == 0 || a > n) continue;                       // row 0 carries no parameter (model_fqandtoyo.py:1160-1165)
        const int Tg = n + 1;
        const DT *row = dB + ((size_t)g * kH * c.T + a) * c.Tp + h * hs;
        const int64_t rowp = c.sq_off[g] + (int64_t)(a - 1) * n - 1;     // pair of cell b lives at rowp + b
        for (int b0 = 0; b0 < Tg; b0 += 16) {
            const int bq = b0 + 4 * p4;
            float dj[4] = {0.f, 0.f, 0.f, 0.f};
            if (bq < Tg) {
                if constexpr (sizeof(DT) == 4) {
                    const float4 dv = *reinterpret_cast<const float4 *>(row + bq);
                    dj[0] = dv.x; dj[1] = dv.y; dj[2] = dv.z; dj[3] = dv.w;
                } else {
                    for (int l = 0; l < nlayers; ++l) {
                        const uint2 dv = *reinterpret_cast<const uint2 *>(row + (size_t)l * layer_stride + bq);
                        dj[0] += bf16lo(dv.x); dj[1] += __uint_as_float(dv.x & 0xFFFF0000u);
                        dj[2] += bf16lo(dv.y); dj[3] += __uint_as_float(dv.y & 0xFFFF0000u);
                    }
                }
            }
            // keys of the 4 cells of this lane's p4 group, spread over the 8 head lanes:
            //   every lane h: word h of the walk bytes; lane 0 also rel_pos, lane 1 also poi_pos
            uint32_t wj[4];
            int kj[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = bq + j;
                wj[j] = 0u;
                kj[j] = 0;
                if (b >= 1 && b < Tg) {
                    const int64_t pc = rowp + b;
                    if (h < hopw) wj[j] = __ldg(reinterpret_cast<const uint32_t *>(c.edge_in + pc * c.hops) + h);
                    if (h == 0) kj[j] = c.rel_pos[pc];
                    if (h == 1) kj[j] = c.poi_pos[pc];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int b = bq + j;
                const bool valid = b < Tg;
                float d = valid ? dj[j] : 0.f;
                if (b == 0) tacc += d;
                bool pair = valid && b >= 1;
                const int rp = __shfl_sync(full, kj[j], gbase);
                const int pp = __shfl_sync(full, kj[j], gbase + 1);
                // walk length L and the mask of hops that carry another feature than the dominant one (bit k = hop k),
                // reduced over the 8 word lanes of the cell
                const uint32_t wd = wj[j];
                int L = 0;
                uint32_t hm = 0u;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t v = (wd >> (8 * e)) & 0xFFu;
                    L += (v != 0u);
                    hm |= (v != 0u && v != (uint32_t)kDomEdge) ? (1u << e) : 0u;
                }
                hm <<= 4 * h;
                L += __shfl_xor_sync(full, L, 1);
                hm |= __shfl_xor_sync(full, hm, 1);
                L += __shfl_xor_sync(full, L, 2);
                hm |= __shfl_xor_sync(full, hm, 2);
                L += __shfl_xor_sync(full, L, 4);
                hm |= __shfl_xor_sync(full, hm, 4);
                const int M = rp - 1;
                if (M >= c.rel_pos_max) pair = false;            // -inf entries carry no gradient
                // rel_pos histogram: 4 lock-step turns (cells of one group may share a key)
                int rk = -1;
                if (pair) {
                    rk = (rp == 511) ? pl.Rrows - 1 : (rp < pl.Rrows - 1 ? rp : -2);
                    if (rk == -2) atomicAdd(dR_overflow + rp * kH + h, d);     // key outside the plan (never for K1 output)
                }
#pragma unroll
                for (int turn = 0; turn < 4; ++turn) {
                    if (p4 == turn && rk >= 0) sR[rk * kH + h] += d;
                    __syncwarp();
                }
                if (pair) {
                    sP[min(pp, num_bins - 1) * kH + h] += d;
                    d *= 1.0f / (float)min(max(M, 1), c.hops);
                    sA[L * kH + h] += d;                          // L == 0 (no walk): row 0 of A is ignored
                } else {
                    hm = 0u;
                }
                // hops with a non-dominant feature (few): N[k] += d (taken back from the dominant bin at the end) and
                // dEW[k][v] += d in the CTA's shared histogram
                while (__any_sync(full, hm != 0u)) {
                    const int k = hm ? __ffs(hm) - 1 : 0;
                    const uint32_t wv = __shfl_sync(full, wd, gbase + (k >> 2));
                    if (hm) {
                        const int v = (wv >> (8 * (k & 3))) & 0xFF;
                        sN[k * kH + h] += d;
                        atomicAdd(sEW + ((size_t)k * kEdgeVocab + v) * kH + h, d);
                        hm &= hm - 1u;
                    }
                }
            }
        }
    }
    // column-0 sums: lanes (p4 = 0, h) hold them
    if (p4 == 0) sT[h] = tacc;
    __syncthreads();
    // CTA totals -> partial[blockIdx]
    float *out = partial + (size_t)blockIdx.x * pl.stride;
    for (int i = threadIdx.x; i < pl.nEW; i += blockDim.x) out[i] = sEW[i];
    const float *wbase = sm + pl.nEW;
    for (int i = threadIdx.x; i < pl.nR; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += wbase[w * per_warp + i];
        out[pl.nEW + i] = s;
    }
    for (int i = threadIdx.x; i < pl.nP; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w)
            for (int r = 0; r < 4; ++r) s += wbase[w * per_warp + pl.nR + r * pl.nP + i];
        out[pl.nEW + pl.nR + i] = s;
    }
    for (int i = threadIdx.x; i < pl.nA; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w)
            for (int r = 0; r < 4; ++r) s += wbase[w * per_warp + pl.nR + 4 * pl.nP + r * pl.nA + i];
        out[pl.nEW + pl.nR + pl.nP + i] = s;
    }
    for (int i = threadIdx.x; i < pl.nN; i += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w)
            for (int r = 0; r < 4; ++r) s += wbase[w * per_warp + pl.nR + 4 * (pl.nP + pl.nA) + r * pl.nN + i];
        out[pl.nEW + pl.nR + pl.nP + pl.nA + i] = s;
    }
    if (threadIdx.x < kH) {
        float s = 0.f;
        for (int w = 0; w < nwarp; ++w) s += wbase[w * per_warp + pl.nR + 4 * (pl.nP + pl.nA + pl.nN) + threadIdx.x];
        out[pl.nEW + pl.nR + pl.nP + pl.nA + pl.nN + threadIdx.x] = s;
    }
}

// tot[i] = sum over CTAs (fixed order) of partial[cta][i];  then the walk-length histogram is folded into dEW:
// dEW[k][dom][h] += sum_{L > k} A[L][h] - N[k][h]   (second kernel, after the totals exist)
__global__ void k2_bias_bwd_reduce_kernel(const float *__restrict__ partial, int nparts, int stride, float *__restrict__ tot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= stride) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partial[(size_t)p * stride + i];
    tot[i] = s;
}

__global__ void k2_bias_bwd_scatter_kernel(float *__restrict__ tot, const K2BwdPlan pl, int hops, int num_bins,
                                           float *__restrict__ dR, float *__restrict__ dP, float *__restrict__ dt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float *tR = tot + pl.nEW, *tP = tR + pl.nR, *tA = tP + pl.nP, *tN = tA + pl.nA, *tT = tN + pl.nN;
    if (i < 512 * kH) {
        const int rp = i / kH, hh = i % kH;
        const int row = (rp == 511) ? pl.Rrows - 1 : (rp < pl.Rrows - 1 ? rp : -1);
        dR[i] += (row >= 0) ? tR[row * kH + hh] : 0.f;     // dR holds the (normally empty) overflow atomics
    }
    if (i < num_bins * kH) dP[i] = tP[i];
    if (i < kH) dt[i] = tT[i];
    if (i < hops * kH) {
        const int k = i / kH, hh = i % kH;
        float s = 0.f;
        for (int L = k + 1; L <= hops; ++L) s += tA[L * kH + hh];
        s -= tN[i];
        tot[((size_t)k * kEdgeVocab + kDomEdge) * kH + hh] += s;
    }
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

extern "C" int64_t mobgt_bias_fwd_workspace_bytes(int32_t hops, int32_t H) {
    if (H != kH || hops < 4 || hops > MOBGT_MAX_HOPS || hops % 4 != 0) return -1;
    return (int64_t)(hops * kEdgeVocab + kRelRows) * kH * (int64_t)sizeof(float);
}

extern "C" int32_t mobgt_bias_fwd(const int32_t *n, const int64_t *sq_off, const int16_t *rel_pos, const int16_t *poi_pos,
                                  const uint8_t *edge_in, int32_t B, int32_t T, int32_t Tp, int32_t hops, int32_t H,
                                  int32_t rel_pos_max, int32_t num_bins, const float *R, const float *Ppos, const float *E,
                                  const float *W, const float *tvd, void *workspace, void *out, int32_t out_dtype,
                                  void *stream) {
    MOBGT_REQUIRE(n && sq_off && rel_pos && poi_pos && edge_in && R && Ppos && E && W && tvd && workspace && out,
                  MOBGT_ERR_NULL, "mobgt_bias_fwd: null pointer");
    MOBGT_REQUIRE(H == kH, MOBGT_ERR_UNSUPPORTED, "mobgt_bias_fwd: num_heads=%d (only 8 is built)", H);
    MOBGT_REQUIRE(hops >= 4 && hops <= MOBGT_MAX_HOPS && hops % 4 == 0, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_bias_fwd: hops=%d must be a multiple of 4 in [4,%d]", hops, MOBGT_MAX_HOPS);
    MOBGT_REQUIRE(num_bins >= 1 && num_bins <= 1024, MOBGT_ERR_BAD_SHAPE, "mobgt_bias_fwd: num_bins=%d", num_bins);
    MOBGT_REQUIRE(T >= 2 && Tp >= T && Tp % 8 == 0, MOBGT_ERR_BAD_SHAPE, "mobgt_bias_fwd: T=%d Tp=%d", T, Tp);
    MOBGT_REQUIRE(out_dtype == MOBGT_F32 || out_dtype == MOBGT_BF16, MOBGT_ERR_BAD_DTYPE, "mobgt_bias_fwd: dtype");
    if (B <= 0) return MOBGT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *EW = static_cast<float *>(workspace);                 // [hops][128][8] then RL [512][8]
    const int tabn = hops * kEdgeVocab * kH;
    float *RL = EW + tabn;
    k2_prep_kernel<<<ceil_div(tabn + kRelRows * kH, 256), 256, 0, s>>>(E, W, R, hops, rel_pos_max, EW, RL);
    MOBGT_LAUNCH_OK("k2_prep_kernel");
    K2Common c{n, sq_off, rel_pos, poi_pos, edge_in, B, T, Tp, hops, rel_pos_max};
    const size_t smem = (size_t)(kRelRows * kH + num_bins * kH + (hops + 1) * 8) * sizeof(float);
    const int tiles_total = ceil_div(T * (Tp / 2), 512) * B;
    dim3 grid((unsigned)min(2 * kNumSMs, tiles_total));
    auto launch = [&](auto kern, auto *o) -> int32_t {
        MOBGT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 512, smem, s>>>(c, RL, Ppos, num_bins, tvd, EW, o);
        return MOBGT_OK;
    };
    int32_t rc = MOBGT_OK;
#define MOBGT_K2_CASE(HW)                                                                                             \
    case HW:                                                                                                          \
        rc = (out_dtype == MOBGT_F32) ? launch(k2_bias_fwd_kernel<float, HW>, static_cast<float *>(out))              \
                                      : launch(k2_bias_fwd_kernel<__nv_bfloat16, HW>, static_cast<__nv_bfloat16 *>(out)); \
        break;
    switch (hops / 4) {
        MOBGT_K2_CASE(1) MOBGT_K2_CASE(2) MOBGT_K2_CASE(3) MOBGT_K2_CASE(4) MOBGT_K2_CASE(5) MOBGT_K2_CASE(6) MOBGT_K2_CASE(7)
        MOBGT_K2_CASE(8)
        default: MOBGT_REQUIRE(false, MOBGT_ERR_BAD_SHAPE, "mobgt_bias_fwd: hops=%d", hops);
    }
#undef MOBGT_K2_CASE
    if (rc) return rc;
    MOBGT_LAUNCH_OK("k2_bias_fwd_kernel");
    return MOBGT_OK;
}

extern "C" int64_t mobgt_bias_bwd_workspace_bytes(int32_t T, int32_t hops, int32_t num_bins) {
    if (T < 2 || hops < 4 || hops > MOBGT_MAX_HOPS || num_bins < 1 || num_bins > 1024) return -1;
    const K2BwdPlan pl = k2_bwd_plan(T, hops, num_bins);
    return (int64_t)(kNumSMs + 1) * pl.stride * (int64_t)sizeof(float);
}

extern "C" int32_t mobgt_bias_bwd(const int32_t *n, const int64_t *sq_off, const int16_t *rel_pos, const int16_t *poi_pos,
                                  const uint8_t *edge_in, int32_t B, int32_t T, int32_t Tp, int32_t hops, int32_t H,
                                  int32_t rel_pos_max, int32_t num_bins, const void *dBias, int32_t dbias_dtype,
                                  int32_t n_layers, int64_t layer_stride, const float *E, const float *W,
                                  void *workspace, int64_t workspace_bytes, float *dR, float *dPpos, float *dE, float *dW,
                                  float *dtvd, void *stream) {
    MOBGT_REQUIRE(n && sq_off && rel_pos && poi_pos && edge_in && dBias && E && W && workspace && dR && dPpos && dE && dW && dtvd,
                  MOBGT_ERR_NULL, "mobgt_bias_bwd: null pointer");
    MOBGT_REQUIRE(H == kH, MOBGT_ERR_UNSUPPORTED, "mobgt_bias_bwd: num_heads=%d (only 8 is built)", H);
    MOBGT_REQUIRE(hops >= 4 && hops <= MOBGT_MAX_HOPS && hops % 4 == 0 && num_bins >= 1 && num_bins <= 1024, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_bias_bwd: hops=%d num_bins=%d", hops, num_bins);
    MOBGT_REQUIRE(T >= 2 && Tp >= T && Tp % 8 == 0, MOBGT_ERR_BAD_SHAPE, "mobgt_bias_bwd: T=%d Tp=%d", T, Tp);
    MOBGT_REQUIRE((dbias_dtype == MOBGT_F32 && n_layers == 1) || (dbias_dtype == MOBGT_BF16 && n_layers >= 1 && n_layers <= 64),
                  MOBGT_ERR_BAD_DTYPE, "mobgt_bias_bwd: dbias dtype %d with %d layer planes", dbias_dtype, n_layers);
    const K2BwdPlan pl = k2_bwd_plan(T, hops, num_bins);
    MOBGT_REQUIRE(pl.warps >= 1, MOBGT_ERR_UNSUPPORTED, "mobgt_bias_bwd: no shared-memory plan for T=%d bins=%d", T, num_bins);
    const int64_t need = (int64_t)(kNumSMs + 1) * pl.stride * (int64_t)sizeof(float);
    MOBGT_REQUIRE(workspace_bytes >= need, MOBGT_ERR_WORKSPACE_TOO_SMALL, "mobgt_bias_bwd: workspace %lld < %lld bytes",
                  (long long)workspace_bytes, (long long)need);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *tot = static_cast<float *>(workspace);          // [stride] totals, then [kNumSMs][stride] per-CTA partials
    float *partial = tot + pl.stride;
    MOBGT_CUDA_OK(cudaMemsetAsync(dR, 0, 512 * kH * 4, s));
    const int nparts = B > 0 ? kNumSMs : 0;
    if (B > 0) {
        K2Common c{n, sq_off, rel_pos, poi_pos, edge_in, B, T, Tp, hops, rel_pos_max};
        const int per_warp = pl.nR + 4 * (pl.nP + pl.nA + pl.nN) + kH;
        const size_t smem = (size_t)(pl.nEW + pl.warps * per_warp) * sizeof(float);
        if (dbias_dtype == MOBGT_F32) {
            MOBGT_CUDA_OK(cudaFuncSetAttribute(k2_bias_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k2_bias_bwd_kernel<float><<<kNumSMs, pl.warps * 32, smem, s>>>(c, static_cast<const float *>(dBias), 1, 0, num_bins, pl,
                                                                          partial, dR);
        } else {
            MOBGT_CUDA_OK(cudaFuncSetAttribute(k2_bias_bwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)smem));
            k2_bias_bwd_kernel<__nv_bfloat16><<<kNumSMs, pl.warps * 32, smem, s>>>(
                c, static_cast<const __nv_bfloat16 *>(dBias), n_layers, layer_stride, num_bins, pl, partial, dR);
        }
        MOBGT_LAUNCH_OK("k2_bias_bwd_kernel");
    }
    k2_bias_bwd_reduce_kernel<<<ceil_div(pl.stride, 256), 256, 0, s>>>(partial, nparts, pl.stride, tot);
    MOBGT_LAUNCH_OK("k2_bias_bwd_reduce_kernel");
    k2_bias_bwd_scatter_kernel<<<ceil_div(512 * kH, 256), 256, 0, s>>>(tot, pl, hops, num_bins, dR, dPpos, dtvd);
    MOBGT_LAUNCH_OK("k2_bias_bwd_scatter_kernel");
    k2_bias_bwd_finish_kernel<<<ceil_div(kEdgeVocab * kH + hops * kH * kH, 256), 256, 0, s>>>(tot, E, W, hops, dE, dW);
    MOBGT_LAUNCH_OK("k2_bias_bwd_finish_kernel");
    return MOBGT_OK;
}
