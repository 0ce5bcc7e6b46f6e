// K4 — node-embedding gather / sum, forward, and the deterministic segmented scatter-add backward
// (sm_100a, HBM-bound gathers).
//
// Replaces model_fqandtoyo.py:1222-1344: the per-graph Python loop that looks up
//   global_poidistemb[x-1] (128) | time_embed_model_48[(time_normal*48).long()] (32) | global_catemb[cat_of_poi(x)-1] (32)
// for every node (:1257-1269), and the sum  nf + in_degree_encoder + out_degree_encoder + pe[q+1]  with the
// graph token  graph_token + pe[0]  (:1288-1344).  The two small FuseEmbeddings linears in between
// (160x160, 192x192, :1268-1269) are library GEMMs on the host side.
//
// Tokens are PACKED (no padding rows): graph g owns token rows tok_off[g] .. tok_off[g]+n[g], row 0 of each
// graph is the graph token.  Node v of the batch (packed node index) is token row v + g + 1.
//
// Backward = segmented sum over a stable sort of the index stream (no atomics, fixed summation order):
//   pass A: every chunk of 32 sorted rows sums its runs of equal keys; a run that lies strictly inside one chunk is a
//           complete segment and is stored straight to the table, runs touching a chunk boundary go to head/tail slots;
//   pass B: the chunk where a boundary-crossing segment starts adds the following head slots in order.
#include <cuda_bf16.h>

#include "common.cuh"

namespace mobgt {

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// out[v, 0:Dp | Dp:Dp+Dt | Dp+Dt:Dp+Dt+Dc] = Gd[x[v]-1] | Tm[slot[v]] | Gc[cat_of_poi[x[v]-1]-1]
template <typename OutT>
__global__ void k4_gather_fwd_kernel(const int32_t *__restrict__ x, const int32_t *__restrict__ slot,
                                     const int32_t *__restrict__ cat_of_poi, const float *__restrict__ Gd,
                                     const float *__restrict__ Tm, const float *__restrict__ Gc, int nnode, int Dp, int Dt,
                                     int Dc, OutT *__restrict__ out) {
    const int D = Dp + Dt + Dc;
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int v = blockIdx.x * wpb + (threadIdx.x >> 5); v < nnode; v += gridDim.x * wpb) {
        const int poi = x[v] - 1;
        const float *a = Gd + (size_t)poi * Dp;
        const float *b = Tm + (size_t)slot[v] * Dt;
        const float *c = Gc + (size_t)(cat_of_poi[poi] - 1) * Dc;
        OutT *o = out + (size_t)v * D;
        for (int e = lane; e < D; e += 32) {
            const float val = e < Dp ? a[e] : (e < Dp + Dt ? b[e - Dp] : c[e - Dp - Dt]);
            o[e] = from_f<OutT>(val);
        }
    }
}

// tok[row] = (pos == 0) ? graph_token + pe[0] : nf[node] + Din[indeg[node]] + Dout[outdeg[node]] + pe[pos]
template <typename T>
__global__ void k4_sum_fwd_kernel(const T *__restrict__ nf, const int32_t *__restrict__ tok_graph,
                                  const int32_t *__restrict__ tok_pos, const int32_t *__restrict__ in_deg,
                                  const int32_t *__restrict__ out_deg, const float *__restrict__ Din,
                                  const float *__restrict__ Dout, const float *__restrict__ pe,
                                  const float *__restrict__ graph_token, int ntok, int D, T *__restrict__ tok) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < ntok; r += gridDim.x * wpb) {
        const int pos = tok_pos[r];
        T *o = tok + (size_t)r * D;
        if (pos == 0) {
            for (int e = lane; e < D; e += 32) o[e] = from_f<T>(graph_token[e] + pe[e]);
        } else {
            const int node = r - tok_graph[r] - 1;
            const T *src = nf + (size_t)node * D;
            const float *di = Din + (size_t)in_deg[node] * D;
            const float *dn = Dout + (size_t)out_deg[node] * D;
            const float *pp = pe + (size_t)pos * D;
            for (int e = lane; e < D; e += 32) o[e] = from_f<T>(((to_f<T>(src[e]) + di[e]) + dn[e]) + pp[e]);
        }
    }
}

// ---- deterministic segmented sum -------------------------------------------------------------------
constexpr int kChunk = 32;
constexpr int kHasHead = 1, kHeadContinues = 2, kHasTail = 4;

// src row r (0..nrows) = src + r*src_stride + col0, D columns.  perm / keys are in sorted order.
// grid = nchunks; block = D/4 threads (one float4 column group each).
template <typename T>
__global__ void k4_segsum_a_kernel(const T *__restrict__ src, int64_t src_stride, int col0, int D,
                                   const int32_t *__restrict__ perm, const int32_t *__restrict__ keys, int nrows,
                                   float *__restrict__ table, int nkeys, float *__restrict__ part, int32_t *__restrict__ flags,
                                   int32_t *__restrict__ tail_key) {
    const int c = blockIdx.x;
    const int r0 = c * kChunk, r1 = min(nrows, r0 + kChunk);
    // the chunk's permutation and keys (with one key of context on either side) are fetched once, coalesced, into shared
    // memory: the row loop below then issues nothing but independent row loads, four rows ahead — it used to pay a dependent
    // perm[r] -> row -> keys[r + 1] chain of global loads per row (13 us per launch however small the batch)
    __shared__ int sperm[kChunk], skey[kChunk + 2];
    for (int i = threadIdx.x; i < kChunk + 2; i += blockDim.x) {
        const int r = r0 - 1 + i;
        skey[i] = (r >= 0 && r < nrows) ? keys[r] : -1;
        if (i < kChunk) sperm[i] = (r0 + i < r1) ? perm[r0 + i] : 0;
    }
    __syncthreads();
    const int e0 = threadIdx.x * 4;
    if (e0 >= D) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int fl = 0, tkey = -1;
    int run_start = r0;
    int key = skey[1];
    const int nr = r1 - r0;
    // all rows of the chunk are in flight at once (one global-memory latency per chunk): 8-byte loads of 4 bf16 — 2 registers
    // per row, 32 rows — or 16-byte loads of 4 floats, 16 rows at a time; then the sums run in row order (fixed: reproducible)
    constexpr int kBatch = sizeof(T) == 2 ? kChunk : kChunk / 2;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    for (int i0 = 0; i0 < nr; i0 += kBatch) {
        float v[kBatch][sizeof(T) == 2 ? 2 : 4];        // bf16: two packed words per row; f32: four floats
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            if (i0 + u < nr) {
                const T *row = src + (size_t)sperm[i0 + u] * src_stride + col0 + e0;
                if constexpr (sizeof(T) == 2) {
                    if (vec_ok) {
                        const uint2 w = *reinterpret_cast<const uint2 *>(row);
                        v[u][0] = __uint_as_float(w.x);
                        v[u][1] = __uint_as_float(w.y);
                    } else {
                        const unsigned short *h = reinterpret_cast<const unsigned short *>(row);
                        v[u][0] = __uint_as_float((uint32_t)h[0] | ((uint32_t)h[1] << 16));
                        v[u][1] = __uint_as_float((uint32_t)h[2] | ((uint32_t)h[3] << 16));
                    }
                } else {
                    if (vec_ok) {
                        const float4 w = *reinterpret_cast<const float4 *>(row);
                        v[u][0] = w.x; v[u][1] = w.y; v[u][2] = w.z; v[u][3] = w.w;
                    } else {
                        v[u][0] = to_f<T>(row[0]); v[u][1] = to_f<T>(row[1]); v[u][2] = to_f<T>(row[2]); v[u][3] = to_f<T>(row[3]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
            const int i = i0 + u;
            if (i >= nr) break;
            const int r = r0 + i;
            if constexpr (sizeof(T) == 2) {
                const uint32_t w0 = __float_as_uint(v[u][0]), w1 = __float_as_uint(v[u][1]);
                acc.x += __uint_as_float(w0 << 16);
                acc.y += __uint_as_float(w0 & 0xFFFF0000u);
                acc.z += __uint_as_float(w1 << 16);
                acc.w += __uint_as_float(w1 & 0xFFFF0000u);
            } else {
                acc.x += v[u][0];
                acc.y += v[u][1];
                acc.z += v[u][2];
                acc.w += v[u][3];
            }
            const int nkey = skey[i + 2];                       // keys[r + 1], -1 past the end
            if (r + 1 == r1 || nkey != key) {  // run [run_start, r] ends here
                const bool from_prev = (run_start == r0) && (r0 > 0) && (skey[0] == key);
                const bool to_next = (r + 1 == r1) && (nkey == key);
                if (!from_prev && !to_next) {
                    if (key >= 0 && key < nkeys) *reinterpret_cast<float4 *>(table + (size_t)key * D + e0) = acc;
                } else if (from_prev) {
                    *reinterpret_cast<float4 *>(part + ((size_t)c * 2 + 0) * D + e0) = acc;
                    fl |= kHasHead | (to_next ? kHeadContinues : 0);
                } else {
                    *reinterpret_cast<float4 *>(part + ((size_t)c * 2 + 1) * D + e0) = acc;
                    fl |= kHasTail;
                    tkey = key;
                }
                acc = make_float4(0.f, 0.f, 0.f, 0.f);
                run_start = r + 1;
                key = nkey;
            }
        }
    }
    if (threadIdx.x == 0) {
        flags[c] = fl;
        tail_key[c] = tkey;
    }
}

// Pass B: a run that leaves chunk c through its tail continues through the heads of the following chunks.  The end of the
// run is found by a ballot scan of the chunk flags; the head partials are then summed by kSegRows thread rows over strided
// chunks and folded in a fixed order (deterministic).  Long runs (low-cardinality keys such as degrees: hundreds of chunks)
// no longer walk the chain one chunk at a time.
constexpr int kSegRows = 8;
__global__ void k4_segsum_b_kernel(const float *__restrict__ part, const int32_t *__restrict__ flags,
                                   const int32_t *__restrict__ tail_key, int nchunks, int D, float *__restrict__ table,
                                   int nkeys) {
    extern __shared__ __align__(16) float sfold[];      // [kSegRows][D]
    __shared__ int s_end;
    const int c = blockIdx.x;
    if (!(flags[c] & kHasTail)) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (ty == 0 && tx < 32) {
        int end = nchunks;                               // exclusive end of the chunks whose head belongs to this run
        for (int base = c + 1; base < nchunks; base += 32) {
            const int cc = base + tx;
            const int f = cc < nchunks ? flags[cc] : 0;
            const bool excl = !(f & kHasHead);                              // the run ended before chunk cc
            const bool incl = (f & kHasHead) && !(f & kHeadContinues);      // the run ends inside chunk cc
            const unsigned m = __ballot_sync(0xffffffffu, excl || incl);
            if (m != 0u) {
                const int L = __ffs(m) - 1;
                const bool incl_first = __shfl_sync(0xffffffffu, (int)incl, L) != 0;
                end = base + L + (incl_first ? 1 : 0);
                break;
            }
        }
        if (tx == 0) s_end = end;
    }
    __syncthreads();
    const int end = s_end;
    const int e0 = tx * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e0 < D) {
        if (ty == 0) acc = *reinterpret_cast<const float4 *>(part + ((size_t)c * 2 + 1) * D + e0);
        for (int cc = c + 1 + ty; cc < end; cc += (int)blockDim.y) {
            const float4 h = *reinterpret_cast<const float4 *>(part + ((size_t)cc * 2 + 0) * D + e0);
            acc.x += h.x; acc.y += h.y; acc.z += h.z; acc.w += h.w;
        }
        *reinterpret_cast<float4 *>(sfold + (size_t)ty * D + e0) = acc;
    }
    __syncthreads();
    if (ty == 0 && e0 < D) {
        float4 t = *reinterpret_cast<const float4 *>(sfold + e0);
        for (int r = 1; r < (int)blockDim.y; ++r) {
            const float4 h = *reinterpret_cast<const float4 *>(sfold + (size_t)r * D + e0);
            t.x += h.x; t.y += h.y; t.z += h.z; t.w += h.w;
        }
        const int key = tail_key[c];
        if (key >= 0 && key < nkeys) *reinterpret_cast<float4 *>(table + (size_t)key * D + e0) = t;
    }
}

template <typename T>
static int32_t launch_gather(const int32_t *x, const int32_t *slot, const int32_t *cat_of_poi, const float *Gd,
                             const float *Tm, const float *Gc, int nnode, int Dp, int Dt, int Dc, void *out, cudaStream_t s) {
    const int blocks = min(ceil_div(nnode, 8), 8 * kNumSMs);
    k4_gather_fwd_kernel<T><<<blocks, 256, 0, s>>>(x, slot, cat_of_poi, Gd, Tm, Gc, nnode, Dp, Dt, Dc, static_cast<T *>(out));
    MOBGT_LAUNCH_OK("k4_gather_fwd_kernel");
    return MOBGT_OK;
}

}  // namespace mobgt

using namespace mobgt;

extern "C" int32_t mobgt_embed_gather_fwd(const int32_t *x, const int32_t *slot, const int32_t *cat_of_poi, const float *Gd,
                                          const float *Tm, const float *Gc, int32_t nnode, int32_t Dp, int32_t Dt, int32_t Dc,
                                          void *out, int32_t out_dtype, void *stream) {
    MOBGT_REQUIRE(x && slot && cat_of_poi && Gd && Tm && Gc && out, MOBGT_ERR_NULL, "mobgt_embed_gather_fwd: null pointer");
    MOBGT_REQUIRE(out_dtype == MOBGT_F32 || out_dtype == MOBGT_BF16, MOBGT_ERR_BAD_DTYPE, "mobgt_embed_gather_fwd: dtype");
    if (nnode <= 0) return MOBGT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return out_dtype == MOBGT_F32 ? launch_gather<float>(x, slot, cat_of_poi, Gd, Tm, Gc, nnode, Dp, Dt, Dc, out, s)
                                  : launch_gather<__nv_bfloat16>(x, slot, cat_of_poi, Gd, Tm, Gc, nnode, Dp, Dt, Dc, out, s);
}

extern "C" int32_t mobgt_embed_sum_fwd(const void *nf, const int32_t *tok_graph, const int32_t *tok_pos,
                                       const int32_t *in_deg, const int32_t *out_deg, const float *Din, const float *Dout,
                                       const float *pe, const float *graph_token, int32_t ntok, int32_t D, void *tok,
                                       int32_t dtype, void *stream) {
    MOBGT_REQUIRE(nf && tok_graph && tok_pos && in_deg && out_deg && Din && Dout && pe && graph_token && tok, MOBGT_ERR_NULL,
                  "mobgt_embed_sum_fwd: null pointer");
    MOBGT_REQUIRE(dtype == MOBGT_F32 || dtype == MOBGT_BF16, MOBGT_ERR_BAD_DTYPE, "mobgt_embed_sum_fwd: dtype");
    if (ntok <= 0) return MOBGT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int blocks = min(ceil_div(ntok, 8), 8 * kNumSMs);
    if (dtype == MOBGT_F32)
        k4_sum_fwd_kernel<float><<<blocks, 256, 0, s>>>(static_cast<const float *>(nf), tok_graph, tok_pos, in_deg, out_deg, Din,
                                                       Dout, pe, graph_token, ntok, D, static_cast<float *>(tok));
    else
        k4_sum_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16 *>(nf), tok_graph, tok_pos, in_deg,
                                                               out_deg, Din, Dout, pe, graph_token, ntok, D,
                                                               static_cast<__nv_bfloat16 *>(tok));
    MOBGT_LAUNCH_OK("k4_sum_fwd_kernel");
    return MOBGT_OK;
}

// table[key, 0:D] = sum over rows r with keys[r] == key of src[perm[r], col0:col0+D], summed in sorted order.
// perm / keys_sorted come from a STABLE sort of the key stream.  workspace: (2*nchunks*D floats) + 2*nchunks int32,
// nchunks = ceil(nrows/32).  Rows of `table` whose key never occurs are left untouched (zero them first).
extern "C" int32_t mobgt_segment_sum(const void *src, int32_t src_dtype, int64_t src_stride, int32_t col0, int32_t D,
                                     const int32_t *perm, const int32_t *keys_sorted, int32_t nrows, float *table,
                                     int32_t nkeys, void *workspace, int64_t workspace_bytes, void *stream) {
    MOBGT_REQUIRE(src && perm && keys_sorted && table && workspace, MOBGT_ERR_NULL, "mobgt_segment_sum: null pointer");
    MOBGT_REQUIRE(D >= 4 && D % 4 == 0 && D <= 1024 && col0 % 4 == 0 && src_stride % 4 == 0, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_segment_sum: D=%d col0=%d stride=%lld must be multiples of 4", D, col0, (long long)src_stride);
    MOBGT_REQUIRE(src_dtype == MOBGT_F32 || src_dtype == MOBGT_BF16, MOBGT_ERR_BAD_DTYPE, "mobgt_segment_sum: dtype");
    if (nrows <= 0) return MOBGT_OK;
    const int nchunks = ceil_div(nrows, kChunk);
    const int64_t need = (int64_t)2 * nchunks * D * 4 + (int64_t)2 * nchunks * 4;
    MOBGT_REQUIRE(workspace_bytes >= need, MOBGT_ERR_WORKSPACE_TOO_SMALL, "mobgt_segment_sum: workspace %lld < %lld",
                  (long long)workspace_bytes, (long long)need);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float *part = static_cast<float *>(workspace);
    int32_t *flags = reinterpret_cast<int32_t *>(part + (size_t)2 * nchunks * D);
    int32_t *tail_key = flags + nchunks;
    const int threads = round_up(D / 4, 32);
    if (src_dtype == MOBGT_F32)
        k4_segsum_a_kernel<float><<<nchunks, threads, 0, s>>>(static_cast<const float *>(src), src_stride, col0, D, perm,
                                                              keys_sorted, nrows, table, nkeys, part, flags, tail_key);
    else
        k4_segsum_a_kernel<__nv_bfloat16><<<nchunks, threads, 0, s>>>(static_cast<const __nv_bfloat16 *>(src), src_stride, col0,
                                                                      D, perm, keys_sorted, nrows, table, nkeys, part, flags,
                                                                      tail_key);
    MOBGT_LAUNCH_OK("k4_segsum_a_kernel");
    const int seg_rows = min(kSegRows, 1024 / threads);
    k4_segsum_b_kernel<<<nchunks, dim3(threads, seg_rows), (size_t)seg_rows * D * sizeof(float), s>>>(part, flags, tail_key, nchunks, D,
                                                                                                 table, nkeys);
    MOBGT_LAUNCH_OK("k4_segsum_b_kernel");
    return MOBGT_OK;
}
