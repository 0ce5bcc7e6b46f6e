// K5 — POI-logit head: tensor-core GEMM fused with per-row top-k and rank counting (sm_100a).
//
// Replaces out_proj + the evaluation metrics path, model_fqandtoyo.py:1396/1408/1421 (logits = z W^T + b) followed by
// get_acc (:48-90: topk(20) on the device, then a numpy loop) and MRR_metric (:122-131: a full descending argsort of every
// row on the CPU).  Here the [M, V] logits never reach HBM: each CTA keeps a 128-row tile of z resident, streams 128-row
// tiles of its vocabulary slice through tcgen05.mma (accumulator in TMEM), and the epilogue threads (one per row) keep
//   * a running sorted top-k list of (value, global index)            -> Acc@k / NDCG@k
//   * count(s > s_t) and count(s == s_t and idx < t)  (rank of target) -> MRR (ties towards the lower index)
// Two modes: mode 0 extracts s_t (the target's own logit, from the same MMA arithmetic, only on tiles that contain a target),
// mode 1 counts and selects.  Vocabulary sharding across GPUs = `vocab_offset` + merging the per-shard lists
// (mobgt_topk_merge) after an NCCL all-gather; counts are summed.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace mobgt {
using namespace sm100;

constexpr int kHeadTile = 128;
constexpr int kHeadMaxK = 320;     // 2*hidden + 64 (model_fqandtoyo.py:1059-1068)
constexpr int kHeadMaxTop = 32;

struct HeadParams {
    const float *bias;       // [V] or null
    const int32_t *target;   // [M] global vocabulary index (or < 0)
    float *st;               // [M]
    float *topk_val;         // [M, nsplit, k]
    int32_t *topk_idx;       // [M, nsplit, k]
    int32_t *cnt_gt, *cnt_eq;  // [M, nsplit]
    float *logits;           // optional [M, V]
    int M, V, K, k, nsplit, mode;
    int64_t vocab_offset;
};

__global__ void __launch_bounds__(128, 1)
k5_head_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW, const HeadParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_a, bar_b, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float sbias[kHeadTile];
    __shared__ int any_target;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int mt = blockIdx.x, sp = blockIdx.y;
    const int chunks = p.K / 8;
    const int tile_bytes = chunks * kHeadTile * 16;
    uint8_t *sA = smem;
    uint8_t *sB = sA + tile_bytes;
    float *lval = reinterpret_cast<float *>(sB + tile_bytes);            // [k][128]
    int32_t *lidx = reinterpret_cast<int32_t *>(lval + p.k * kHeadTile);  // [k][128]

    const int ntiles = ceil_div(p.V, kHeadTile);
    const int tps = ceil_div(ntiles, p.nsplit);
    const int n_begin = sp * tps, n_end = min(ntiles, n_begin + tps);

    if (tid == 0) {
        mbar_init(&bar_a, 1);
        mbar_init(&bar_b, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmZ);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 0) tmem_alloc<128>(&tmem_slot);
    for (int j = 0; j < p.k; ++j) {
        lval[j * kHeadTile + tid] = -INFINITY;
        lidx[j * kHeadTile + tid] = -1;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

    const int row = mt * kHeadTile + tid;
    const bool row_ok = row < p.M;
    const long long tgt = row_ok ? (long long)p.target[row] : -1;       // global index
    const float st = (p.mode == 1 && row_ok) ? p.st[row] : 0.f;
    const long long tloc = tgt - p.vocab_offset;                         // index inside this shard
    int cgt = 0, ceq = 0, filled = 0;
    float thr = -INFINITY;

    if (tid == 0) {
        mbar_expect_tx(&bar_a, (uint32_t)tile_bytes);
        tma_load_3d(sA, &tmZ, &bar_a, 0, mt * kHeadTile, 0);
    }
    __syncwarp();
    uint32_t ph_b = 0, ph_mma = 0;
    bool a_ready = false;

    for (int n = n_begin; n < n_end; ++n) {
        if (p.mode == 0) {   // only tiles that hold some row's target matter
            if (tid == 0) any_target = 0;
            __syncthreads();
            if (row_ok && tloc >= (long long)n * kHeadTile && tloc < (long long)(n + 1) * kHeadTile) any_target = 1;
            __syncthreads();
            if (!any_target) continue;
        }
        if (tid == 0) {
            mbar_expect_tx(&bar_b, (uint32_t)tile_bytes);
            tma_load_3d(sB, &tmW, &bar_b, 0, n * kHeadTile, 0);
            if (!a_ready) mbar_wait(&bar_a, 0);
            mbar_wait(&bar_b, ph_b);
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(kHeadTile, kHeadTile, 0, 0);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            for (int ks = 0; ks < p.K / 16; ++ks)
                umma_bf16(tmem, make_smem_desc(a0 + ks * 2 * kHeadTile * 16, kHeadTile * 16, 128),
                          make_smem_desc(b0 + ks * 2 * kHeadTile * 16, kHeadTile * 16, 128), idesc, ks > 0);
            umma_commit(&bar_mma);
        }
        a_ready = true;
        ph_b ^= 1;
        {
            const int col = n * kHeadTile + tid;
            sbias[tid] = (p.bias != nullptr && col < p.V) ? p.bias[col] : 0.f;
        }
        __syncthreads();
        mbar_wait(&bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        for (int c0 = 0; c0 < kHeadTile; c0 += 16) {
            uint32_t acc[16];
            tmem_ld16(tmem + lane_off + c0, acc);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const int col = n * kHeadTile + c0 + e;
                if (col >= p.V || !row_ok) break;
                const float v = __uint_as_float(acc[e]) + sbias[c0 + e];
                const long long gi = p.vocab_offset + col;
                if (p.logits) p.logits[(size_t)row * p.V + col] = v;
                if (p.mode == 0) {
                    if (gi == tgt) p.st[row] = v;
                } else {
                    if (gi != tgt) {
                        cgt += v > st;
                        ceq += (v == st) && (gi < tgt);
                    }
                    if (filled < p.k || v > thr) {      // ascending index scan: an equal value never displaces an earlier one
                        int pos = filled < p.k ? filled : p.k - 1;
                        while (pos > 0 && lval[(pos - 1) * kHeadTile + tid] < v) {
                            lval[pos * kHeadTile + tid] = lval[(pos - 1) * kHeadTile + tid];
                            lidx[pos * kHeadTile + tid] = lidx[(pos - 1) * kHeadTile + tid];
                            --pos;
                        }
                        lval[pos * kHeadTile + tid] = v;
                        lidx[pos * kHeadTile + tid] = (int32_t)gi;
                        if (filled < p.k) ++filled;
                        if (filled == p.k) thr = lval[(p.k - 1) * kHeadTile + tid];
                    }
                }
            }
            __syncwarp();     // tcgen05.ld is warp-collective: reconverge before the next chunk
        }
        tc_fence_before();
        __syncthreads();      // TMEM accumulator and sB are free for the next tile
    }
    if (p.mode == 1 && row_ok) {
        const size_t o = ((size_t)row * p.nsplit + sp);
        for (int j = 0; j < p.k; ++j) {
            p.topk_val[o * p.k + j] = lval[j * kHeadTile + tid];
            p.topk_idx[o * p.k + j] = lidx[j * kHeadTile + tid];
        }
        p.cnt_gt[o] = cgt;
        p.cnt_eq[o] = ceq;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

// Merge S sorted (descending, ties -> lower index first) candidate lists per row and add the rank counts.
__global__ void k5_topk_merge_kernel(const float *__restrict__ val, const int32_t *__restrict__ idx,
                                     const int32_t *__restrict__ cgt, const int32_t *__restrict__ ceq, int M, int S, int k,
                                     float *__restrict__ oval, int32_t *__restrict__ oidx, int32_t *__restrict__ rank) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= M) return;
    int head[64];
    for (int s = 0; s < S; ++s) head[s] = 0;
    for (int j = 0; j < k; ++j) {
        int best = -1;
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int s = 0; s < S; ++s) {
            if (head[s] >= k) continue;
            const size_t o = ((size_t)r * S + s) * k + head[s];
            const float v = val[o];
            const int i = idx[o];
            if (i < 0) continue;
            if (best < 0 || v > bv || (v == bv && i < bi)) { best = s; bv = v; bi = i; }
        }
        oval[(size_t)r * k + j] = best >= 0 ? bv : -INFINITY;
        oidx[(size_t)r * k + j] = best >= 0 ? bi : -1;
        if (best >= 0) ++head[best];
    }
    if (rank) {
        int rk = 0;
        for (int s = 0; s < S; ++s) rk += cgt[(size_t)r * S + s] + ceq[(size_t)r * S + s];
        rank[r] = rk;
    }
}

}  // namespace mobgt

using namespace mobgt;

extern "C" int32_t mobgt_head_topk(const void *z, const void *W, const float *bias, const int32_t *target, int32_t M,
                                   int32_t V, int32_t K, int64_t vocab_offset, int32_t k, int32_t nsplit, int32_t mode,
                                   float *st, float *topk_val, int32_t *topk_idx, int32_t *cnt_gt, int32_t *cnt_eq,
                                   float *logits_dump, void *stream) {
    MOBGT_REQUIRE(z && W && target && st, MOBGT_ERR_NULL, "mobgt_head_topk: null pointer");
    MOBGT_REQUIRE(mode == 0 || (topk_val && topk_idx && cnt_gt && cnt_eq), MOBGT_ERR_NULL, "mobgt_head_topk: null output");
    MOBGT_REQUIRE(K % 16 == 0 && K >= 16 && K <= kHeadMaxK, MOBGT_ERR_BAD_SHAPE, "mobgt_head_topk: K=%d", K);
    MOBGT_REQUIRE(k >= 1 && k <= kHeadMaxTop && nsplit >= 1 && nsplit <= 64, MOBGT_ERR_BAD_SHAPE, "mobgt_head_topk: k=%d nsplit=%d",
                  k, nsplit);
    if (M <= 0 || V <= 0) return MOBGT_OK;
    CUtensorMap tmZ, tmW;
    {
        uint64_t dims[3] = {8, (uint64_t)M, (uint64_t)K / 8};
        uint64_t str[2] = {(uint64_t)K * 2, 16};
        uint32_t box[3] = {8, kHeadTile, (uint32_t)K / 8};
        int32_t rc = encode_tmap_bf16(&tmZ, z, 3, dims, str, box, 0);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {8, (uint64_t)V, (uint64_t)K / 8};
        uint64_t str[2] = {(uint64_t)K * 2, 16};
        uint32_t box[3] = {8, kHeadTile, (uint32_t)K / 8};
        int32_t rc = encode_tmap_bf16(&tmW, W, 3, dims, str, box, 0);
        if (rc) return rc;
    }
    const size_t smem = (size_t)2 * (K / 8) * kHeadTile * 16 + (size_t)2 * k * kHeadTile * 4 + 1024;
    MOBGT_CUDA_OK(cudaFuncSetAttribute(k5_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HeadParams p{bias, target, st, topk_val, topk_idx, cnt_gt, cnt_eq, logits_dump, M, V, K, k, nsplit, mode, vocab_offset};
    dim3 grid((unsigned)ceil_div(M, kHeadTile), (unsigned)nsplit);
    k5_head_kernel<<<grid, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmZ, tmW, p);
    MOBGT_LAUNCH_OK("k5_head_kernel");
    return MOBGT_OK;
}

extern "C" int32_t mobgt_topk_merge(const float *val, const int32_t *idx, const int32_t *cnt_gt, const int32_t *cnt_eq,
                                    int32_t M, int32_t S, int32_t k, float *out_val, int32_t *out_idx, int32_t *rank,
                                    void *stream) {
    MOBGT_REQUIRE(val && idx && out_val && out_idx, MOBGT_ERR_NULL, "mobgt_topk_merge: null pointer");
    MOBGT_REQUIRE(!rank || (cnt_gt && cnt_eq), MOBGT_ERR_NULL, "mobgt_topk_merge: rank needs the counts");
    MOBGT_REQUIRE(S >= 1 && S <= 64 && k >= 1 && k <= kHeadMaxTop, MOBGT_ERR_BAD_SHAPE, "mobgt_topk_merge: S=%d k=%d", S, k);
    if (M <= 0) return MOBGT_OK;
    k5_topk_merge_kernel<<<ceil_div(M, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(val, idx, cnt_gt, cnt_eq, M, S, k, out_val,
                                                                                         out_idx, rank);
    MOBGT_LAUNCH_OK("k5_topk_merge_kernel");
    return MOBGT_OK;
}
