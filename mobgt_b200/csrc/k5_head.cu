// K5 — POI-logit head: tensor-core GEMM fused with per-row top-k and rank counting (sm_100a).
//
// Replaces out_proj + the evaluation metrics path, model_fqandtoyo.py:1396/1408/1421 (logits = z W^T + b) followed by
// get_acc (:48-90: topk(20) on the device, then a numpy loop) and MRR_metric (:122-131: a full descending argsort of every
// row on the CPU).  The [M, V] logits never reach HBM.
//
// mode 1 (k5_head_kernel) — warp-specialised, one CTA per (128-row tile of z, vocabulary split):
//   warp 0      TMA producer: z tile once (resident, K/64 blocks of [128 rows][128 B]); W tiles stream through a ring of
//               32 KB stages, one stage = 256 vocabulary rows x 64 K-columns, 2-D boxes with the 128-byte swizzle
//   warp 1      MMA issuer: tcgen05.mma M=128 N=256 K=16, K/16 steps per tile, accumulators double-buffered in TMEM
//               (2 x 256 columns); a stage is released by tcgen05.commit, a finished accumulator is published the same way
//   warp 2      TMEM allocation;  warp 3 idle
//   warps 4-11  two epilogue warpgroups, one per accumulator buffer (tile t goes to group t & 1).  Thread = row.  Per element:
//               + bias, one compare for the rank count, a 4-wide max against the running k-th value.  The rank
//                   rank = #(s > s_t) + #(s == s_t and idx < t)       (ties towards the lower index)
//               needs ONE compare per element: columns before the target are compared against prev_float(s_t) (>= s_t),
//               columns after it against s_t; only the tile that holds the target takes a per-element path.
//               Top-k candidates (rare after the first tiles) are appended to a small per-row buffer and inserted into
//               the row's sorted list (64-bit keys: order-preserving value bits | ~index) in a lane-aligned drain loop, so
//               the warp does not serialise 32 rows' insertions one after the other.
//   Each (row, split, group) emits a sorted top-k list + its partial rank count; mobgt_topk_merge merges them.
// mode 0 (k5_target_logit_kernel) — s_t, the target's own logit, from the SAME MMA arithmetic: per 128-row tile the W rows
//   of the rows' targets are gathered into the swizzled B tile, one 128x128 MMA group runs, and s_t is the diagonal.
// Vocabulary sharding across GPUs = `vocab_offset` + all-reduce MAX of s_t + merging the per-shard lists after an
// all-gather (parallel.sharded_head_topk); counts are summed.
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace mobgt {
using namespace sm100;

constexpr int kHeadTile = 128;
constexpr int kHeadMaxK = 320;     // 2*hidden + 64 (model_fqandtoyo.py:1059-1068)
constexpr int kHeadMaxTop = 32;
constexpr int kKB = 64;                          // K columns per stage = one 128-byte swizzle row
constexpr int kBlkBytes = kHeadTile * 128;       // 16 KB: [128 rows][128 B] (one K block of the resident z tile)
constexpr int kNT = 256;                         // vocabulary rows per tile = MMA N.  With both operands in shared memory an
                                                 // M = 128, N = 128, K = 16 MMA reads 8 KB per 64 tensor cycles = the whole
                                                 // 128 B/clk of the SM's shared memory; N = 256 reads 12 KB per 128 cycles.
constexpr int kStageBytes = kNT * 128;           // 32 KB: [256 vocabulary rows][128 B]
constexpr int kMaxRing = 4;
constexpr int kCandCap = 16;
constexpr int kHeadThreads = 384;                // 4 service warps + 2 epilogue warpgroups
bool g_head_no_cluster = false;                  // mobgt_debug_head_cluster(0): measurement switch (scripts/k5bench.py)

struct HeadParams {
    const float *bias;       // [V] or null
    const int32_t *target;   // [M] global vocabulary index (or < 0)
    float *st;               // [M]
    float *topk_val;         // [M, nsplit, k]
    int32_t *topk_idx;       // [M, nsplit, k]
    int32_t *cnt_gt, *cnt_eq;  // [M, nsplit]
    float *logits;           // optional [M, V]
    int M, V, K, k, nsplit, mode;
    int ring;                // W stages in shared memory
    long long *timeline;     // debug (mobgt_debug_set_timeline) or NULL
    uint32_t *thr_share;     // [M] best k-th value seen by ANY list of the row (order-preserving bits; 0 = none) or NULL
    int64_t vocab_offset;
};

__device__ __forceinline__ uint32_t f2ord(float v) {
    const uint32_t u = __float_as_uint(v);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
// largest float strictly below x (x finite or +inf); -inf stays -inf:  v >= x  <=>  v > prev_float(x)
__device__ __forceinline__ float prev_float(float x) {
    const uint32_t u = __float_as_uint(x);
    if ((u << 1) == 0u) return __uint_as_float(0x80000001u);   // +-0 -> -denorm_min
    if (u == 0xFF800000u) return x;
    return __uint_as_float((u & 0x80000000u) ? u + 1u : u - 1u);
}

__device__ __forceinline__ unsigned long long lds64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, unsigned long long v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
constexpr uint32_t kEnt = kHeadTile * 8;   // byte distance between consecutive entries of one row's list / buffer

// ---- per-row selection state ---------------------------------------------------------------------------------------
// The sorted top-k list of a row lives in the REGISTERS of the row's thread (LK 64-bit keys: order-preserving value bits in
// the high word, ~global index in the low word, so a larger key is a better entry and ties go to the lower index).  An
// insertion is a fully unrolled compare-exchange chain from the bottom of the list — ALU only, no dependent shared-memory
// round trips.  Candidates (rare after the first tiles) are appended to a small per-row buffer in shared memory and
// inserted at the end of the tile, when all 32 lanes of the warp drain together; a buffer that fills up inside a tile is
// drained at ONE place (the re-run loop of harvest()) so the unrolled insertion exists twice in the kernel, not per group.
template <int LK>
__device__ __forceinline__ void list_insert(unsigned long long (&list)[LK], unsigned long long key) {
    list[LK - 1] = key;      // precondition: key > list[LK-1]
#pragma unroll
    for (int j = LK - 1; j > 0; --j) {
        const unsigned long long a = list[j - 1], b = list[j];
        const bool sw = b > a;
        list[j - 1] = sw ? b : a;
        list[j] = sw ? a : b;
    }
}
// buffer entries are RAW (high word = float bits, low word = column inside this shard); the order-preserving key is built
// here, off the hot path
template <int LK>
__device__ __forceinline__ float list_drain(unsigned long long (&list)[LK], uint32_t cd, int &ncand, long long vocab_offset) {
#pragma unroll 1
    while (ncand > 0) {
        --ncand;
        const unsigned long long raw = lds64(cd + (uint32_t)ncand * kEnt);
        const uint32_t gi = (uint32_t)(vocab_offset + (long long)(uint32_t)raw);
        const unsigned long long key = ((unsigned long long)f2ord(__uint_as_float((uint32_t)(raw >> 32))) << 32) | (uint32_t)(~gi);
        if (key > list[LK - 1]) list_insert<LK>(list, key);
    }
    const unsigned long long kth = list[LK - 1];
    return kth ? ord2f((uint32_t)(kth >> 32)) : -INFINITY;
}
// append one raw candidate entry {column, value bits}; Q = column offset inside the 32-column chunk (immediate)
template <int Q>
__device__ __forceinline__ void cand_append(uint32_t addr, uint32_t cb, float v) {
    asm volatile(
        "{\n\t.reg .u32 c;\n\t"
        "add.u32 c, %1, %2;\n\t"
        "st.shared.v2.u32 [%0], {c, %3};\n\t}"
        ::"r"(addr), "r"(cb), "n"(Q), "r"(__float_as_uint(v))
        : "memory");
}
// Harvest the top-k candidates of a 32-column chunk, group of 4 columns by group.  The caller (all 32 lanes together) has made
// room for 8 entries in every lane's buffer; a lane that finds more in one chunk (a list without a useful bound yet) drains on
// the spot.  The column and the 64-bit entry are formed INSIDE the store's asm block, so nothing of the rare path is hoisted
// into the common one.
template <int LK, int G>
__device__ __forceinline__ void harvest_groups(const float (&v)[32], const float (&m4)[8], unsigned long long (&list)[LK],
                                               float &thr, int &ncand, uint32_t cd, uint32_t cb, long long vocab_offset) {
    if constexpr (G < 8) {
        if (m4[G] > thr) {
            if (ncand > kCandCap - 4) thr = fmaxf(thr, list_drain<LK>(list, cd, ncand, vocab_offset));
            if (v[4 * G] > thr) { cand_append<4 * G>(cd + (uint32_t)ncand * kEnt, cb, v[4 * G]); ++ncand; }
            if (v[4 * G + 1] > thr) { cand_append<4 * G + 1>(cd + (uint32_t)ncand * kEnt, cb, v[4 * G + 1]); ++ncand; }
            if (v[4 * G + 2] > thr) { cand_append<4 * G + 2>(cd + (uint32_t)ncand * kEnt, cb, v[4 * G + 2]); ++ncand; }
            if (v[4 * G + 3] > thr) { cand_append<4 * G + 3>(cd + (uint32_t)ncand * kEnt, cb, v[4 * G + 3]); ++ncand; }
        }
        harvest_groups<LK, G + 1>(v, m4, list, thr, ncand, cd, cb, vocab_offset);
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// spin with back-off: the TMA / MMA warps share their schedulers with epilogue warps and must not eat their issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t *bar, uint32_t parity, unsigned ns = 24) {
    while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}

// kCluster = 2: the CTAs of two adjacent row tiles (same vocabulary split, so the same sequence of W tiles) form a cluster; each
// loads HALF of every W stage and multicasts it into both CTAs' rings, which halves the L2 -> SM traffic of the operand every
// row tile of a split streams.  A ring slot may be refilled only when BOTH CTAs' MMAs have consumed it: bar_empty counts two
// arrivals, delivered by a multicast tcgen05.commit.
template <int LK, int kCluster>
__global__ void __launch_bounds__(kHeadThreads, 1)
k5_head_kernel(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmW, const HeadParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a, bar_full[kMaxRing], bar_empty[kMaxRing], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int warp = warp_index_uniform();
    const int lane = tid & 31;
    const int mt = blockIdx.x, sp = blockIdx.y;
    const int kblocks = ceil_div(p.K, kKB);
    // timing experiments only (scripts/k5bench.py --dbg): mode bits 8.. = skip harvest (1) / skip count + harvest (2); bits 12.. = spin back-off / 8 ns
    const int dbg = (p.mode >> 8) & 15;
    const unsigned spin_ns = ((p.mode >> 12) & 15) ? 8u * ((p.mode >> 12) & 15) : 24u;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + (size_t)kblocks * kBlkBytes;
    unsigned long long *cand = reinterpret_cast<unsigned long long *>(sB + (size_t)p.ring * kStageBytes);   // [2][cap][128]

    const int ntiles = ceil_div(p.V, kNT);
    const int gs = gridDim.y;
    const int tps = ceil_div(ntiles, gs);
    const int n_begin = sp * tps, n_end = min(ntiles, n_begin + tps);
    const int T = max(0, n_end - n_begin);

    if (tid == 0) {
        mbar_init(&bar_a, 1);
        for (int s = 0; s < kMaxRing; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], kCluster);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&acc_full[a], 1);
            mbar_init(&acc_empty[a], 4);
        }
        fence_barrier_init();
        tma_prefetch_desc(&tmZ);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 2) tmem_alloc<2 * kNT>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();      // the peer's barriers exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << kCluster) - 1u);

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            mbar_expect_tx(&bar_a, (uint32_t)(kblocks * kBlkBytes));
            for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(sA + (size_t)kb * kBlkBytes, &tmZ, &bar_a, kb * kKB, mt * kHeadTile);
            int it = 0;
            for (int t = 0; t < T; ++t) {
                const int n = n_begin + t;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % p.ring;
                    const uint32_t ph = (uint32_t)(it / p.ring) & 1u;
                    mbar_wait_relaxed(&bar_empty[s], ph ^ 1u, spin_ns);
                    mbar_expect_tx(&bar_full[s], (uint32_t)kStageBytes);     // own half + the peer's half
                    if (kCluster == 1)
                        tma_load_2d(sB + (size_t)s * kStageBytes, &tmW, &bar_full[s], kb * kKB, n * kNT);
                    else
                        tma_load_2d_mc(sB + (size_t)s * kStageBytes + crank * (kStageBytes / kCluster), &tmW, &bar_full[s], kb * kKB,
                                       n * kNT + (int)crank * (kNT / kCluster), kMask);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(kHeadTile, kNT, 0, 0);
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            mbar_wait(&bar_a, 0);
            int it = 0;
            for (int t = 0; t < T; ++t) {
                const int a = t & 1;
                const bool stamp = p.timeline != nullptr && blockIdx.x == 1 && blockIdx.y == 1 && t < 14;
                if (stamp) p.timeline[16 + 4 * t] = clock64();
                mbar_wait_relaxed(&acc_empty[a], ((uint32_t)(t >> 1) & 1u) ^ 1u, spin_ns);
                tc_fence_after();
                if (stamp) p.timeline[16 + 4 * t + 1] = clock64();
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % p.ring;
                    mbar_wait_relaxed(&bar_full[s], (uint32_t)(it / p.ring) & 1u, spin_ns);
                    tc_fence_after();
                    const int ksteps = min(kKB, p.K - kb * kKB) / 16;
                    for (int j = 0; j < ksteps; ++j)
                        umma_bf16(tmem + (uint32_t)(a * kNT), make_smem_desc_sw128(a0 + kb * kBlkBytes + j * 32),
                                  make_smem_desc_sw128(b0 + s * kStageBytes + j * 32), idesc, (kb | j) != 0);
                    if (kCluster == 1) umma_commit(&bar_empty[s]);
                    else umma_commit_mc(&bar_empty[s], kMask);
                    if (stamp && kb == kblocks - 1) p.timeline[16 + 4 * t + 2] = clock64();
                }
                umma_commit(&acc_full[a]);
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue warpgroups
        const int e = (warp - 4) >> 2;                      // accumulator buffer / tile parity
        const int r = ((warp & 3) << 5) | lane;             // row inside the tile == TMEM lane
        const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
        const int row = mt * kHeadTile + r;
        const bool row_ok = row < p.M;
        const long long tgt = row_ok ? (long long)p.target[row] : -1;
        const long long tloc64 = tgt - p.vocab_offset;      // column of the target inside this shard (any sign / size)
        const int tloc = tloc64 < 0 ? -1 : (tloc64 > 0x7fffff00ll ? 0x7fffff00 : (int)tloc64);
        const float st = row_ok ? p.st[row] : 0.f;
        const float st_prev = prev_float(st);
        const uint32_t cd = smem_u32(cand + (size_t)e * kCandCap * kHeadTile + r);
        const uint32_t sbias = smem_u32(cand + (size_t)2 * kCandCap * kHeadTile) + (uint32_t)e * 2 * kNT * 4;   // [2][256] f32
        unsigned long long list[LK];
#pragma unroll
        for (int j = 0; j < LK; ++j) list[j] = 0ull;
        float thr = -INFINITY;
        int ncand = 0, cnt = 0;
        uint32_t last_pub = 0u;
        // bias of columns col_base + r and col_base + 128 + r of the tile (two elements per thread, coalesced): -inf past the
        // vocabulary, so the zero accumulators of the TMA-zero-filled tail rows neither count nor qualify
        auto tile_bias = [&](int t, int half) -> float {
            const int col = (n_begin + t) * kNT + half * kHeadTile + r;
            return col < p.V ? (p.bias ? __ldg(p.bias + col) : 0.f) : -INFINITY;
        };
        float bnext0 = e < T ? tile_bias(e, 0) : 0.f, bnext1 = e < T ? tile_bias(e, 1) : 0.f;

        int it = 0;
        for (int t = e; t < T; t += 2, ++it) {
            const int n = n_begin + t;
            const int col_base = n * kNT;
            // the best k-th value any list of this row has published so far: a lower bound
            // of the row's final k-th value, so every list may use it as its threshold (>= : ties must survive, hence prev_float)
            const uint32_t gshare = (p.thr_share != nullptr && row_ok) ? __ldcg(p.thr_share + row) : 0u;
            const uint32_t sb = sbias + (uint32_t)(it & 1) * kNT * 4;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb + (uint32_t)r * 4), "f"(bnext0) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(sb + (uint32_t)(kHeadTile + r) * 4), "f"(bnext1) : "memory");
            if (t + 2 < T) {
                bnext0 = tile_bias(t + 2, 0);
                bnext1 = tile_bias(t + 2, 1);
            }
            named_bar_sync(1 + e, kHeadTile);
            mbar_wait(&acc_full[e], (uint32_t)(t >> 1) & 1u);
            tc_fence_after();
            if (gshare != 0u) thr = fmaxf(thr, prev_float(ord2f(gshare)));
            const bool estamp = p.timeline != nullptr && blockIdx.x == 1 && blockIdx.y == 1 && t < 14 && r == 0;
            if (estamp) p.timeline[80 + 4 * t] = clock64();
#pragma unroll 1
            for (int c0 = 0; c0 < kNT; c0 += 32) {
                uint32_t acc[32];
                float v[32];
                const int cb = col_base + c0;
                tmem_ld32(tmem + lane_off + (uint32_t)(e * kNT + c0), acc);
#pragma unroll
                for (int q = 0; q < 8; ++q) {          // same address in every lane: broadcast reads
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(v[4 * q]), "=f"(v[4 * q + 1]), "=f"(v[4 * q + 2]), "=f"(v[4 * q + 3])
                                 : "r"(sb + (uint32_t)(c0 + 4 * q) * 4));
                }
                tmem_ld_wait();
                if (c0 == kNT - 32) {   // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[e]);
                    if (estamp) p.timeline[80 + 4 * t + 1] = clock64();
                }
#pragma unroll
                for (int q = 0; q < 32; ++q) v[q] += __uint_as_float(acc[q]);
                if (p.logits != nullptr && row_ok) {     // logits dump (tests)
#pragma unroll
                    for (int q = 0; q < 32; ++q)
                        if (cb + q < p.V) p.logits[(size_t)row * p.V + cb + q] = v[q];
                }
                // rank count: ONE subtract + one shifted add per element (sign bit of cmp - v  <=>  v > cmp).  Chunks entirely
                // before the target compare against prev_float(s_t) (v >= s_t), the others against s_t; the chunk that holds
                // the target fixes up its own leading columns.
                const float cmp = (tloc >= cb + 32) ? st_prev : st;
                if (tloc >= cb && tloc < cb + 32) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        if (cb + q < tloc) cnt += (int)(v[q] == st);
                        if (cb + q == tloc) cnt -= (int)(v[q] > st);      // the target itself never counts
                    }
                }
                float m4[8];
                if (dbg & 2) {
                    cnt += (int)(__float_as_uint(v[0] + v[31]) >> 31);
                    __syncwarp();
                    continue;
                }
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const float v0 = v[4 * g], v1 = v[4 * g + 1], v2 = v[4 * g + 2], v3 = v[4 * g + 3];
                    cnt += (int)(__float_as_uint(cmp - v0) >> 31) + (int)(__float_as_uint(cmp - v1) >> 31) +
                           (int)(__float_as_uint(cmp - v2) >> 31) + (int)(__float_as_uint(cmp - v3) >> 31);
                    m4[g] = fmaxf(fmaxf(v0, v1), fmaxf(v2, v3));
                }
                // harvest the top-k candidates of the chunk, group by group (rare after the first tiles).  The column and the
                // 64-bit entry are formed INSIDE the store's asm block, so nothing of the rare path is hoisted into the
                // common one; a buffer without room for 4 more entries is drained on the spot.
                if (!(dbg & 1)) {
                    // one vote per 32-column chunk: in steady state no lane holds a candidate in 9 chunks out of 10
                    const float mx = fmaxf(fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])), fmaxf(fmaxf(m4[4], m4[5]), fmaxf(m4[6], m4[7])));
                    if (__any_sync(0xffffffffu, mx > thr)) {
                        // all 32 lanes: whoever is short of room makes EVERY lane drain now — the unrolled insertion chain then
                        // runs once per warp, each lane on its own list, instead of once per lane at 32 different moments
                        if (__any_sync(0xffffffffu, ncand > kCandCap - 8)) thr = fmaxf(thr, list_drain<LK>(list, cd, ncand, p.vocab_offset));
                        harvest_groups<LK, 0>(v, m4, list, thr, ncand, cd, (uint32_t)cb, p.vocab_offset);
                    }
                }
                __syncwarp();     // tcgen05.ld is warp-collective: reconverge before the next chunk
            }
            {   // all lanes together; fresh threshold for the next tile (never below a bound taken from thr_share)
                const float own = list_drain<LK>(list, cd, ncand, p.vocab_offset);
                thr = fmaxf(thr, own);
                const uint32_t kth = (uint32_t)(list[LK - 1] >> 32);
                if (p.thr_share != nullptr && row_ok && kth > last_pub && kth > gshare) {
                    atomicMax(p.thr_share + row, kth);
                    last_pub = kth;
                }
            }
            __syncwarp();
            if (estamp) p.timeline[80 + 4 * t + 2] = clock64();
        }
        if (row_ok) {
            const size_t o = (size_t)row * p.nsplit + (size_t)sp * 2 + e;
#pragma unroll
            for (int j = 0; j < LK; ++j) {
                if (j < p.k) {
                    const unsigned long long key = list[j];
                    p.topk_val[o * p.k + j] = key ? ord2f((uint32_t)(key >> 32)) : -INFINITY;
                    p.topk_idx[o * p.k + j] = key ? (int32_t)(~(uint32_t)key) : -1;
                }
            }
            p.cnt_gt[o] = cnt;
            p.cnt_eq[o] = 0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kCluster > 1) cluster_sync_all();      // no CTA leaves while its peer may still multicast into it / arrive on its barriers
    if (warp == 2) tmem_dealloc<2 * kNT>(tmem);
}

// mode 0: st[row] = z[row] . W[target[row]] + bias[target[row]] through the same tcgen05 arithmetic as mode 1
// (same operand layout, same K-step sequence), for the rows whose target lies in this shard.
struct TargetParams {
    const __nv_bfloat16 *W;
    const float *bias;
    const int32_t *target;
    float *st;
    int M, V, K;
    int64_t vocab_offset;
};

__global__ void __launch_bounds__(128, 1)
k5_target_logit_kernel(const __grid_constant__ CUtensorMap tmZ, const TargetParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_a, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
    const int mt = blockIdx.x;
    const int kblocks = ceil_div(p.K, kKB);
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem;
    uint8_t *sB = sA + (size_t)kblocks * kBlkBytes;

    if (tid == 0) {
        mbar_init(&bar_a, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmZ);
    }
    if (warp == 0) tmem_alloc<128>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (warp == 0 && elect_one()) {
        mbar_expect_tx(&bar_a, (uint32_t)(kblocks * kBlkBytes));
        for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(sA + (size_t)kb * kBlkBytes, &tmZ, &bar_a, kb * kKB, mt * kHeadTile);
    }
    __syncwarp();
    // gather: row r of the B tile = W[target[row] - vocab_offset], written in the 128-byte-swizzle layout
    const int row = mt * kHeadTile + tid;
    const long long tloc = row < p.M ? (long long)p.target[row] - p.vocab_offset : -1;
    const bool own = tloc >= 0 && tloc < p.V;
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.W + (size_t)(own ? tloc : 0) * p.K);
        const int nchunk = p.K / 8;
        for (int c = 0; c < kblocks * 8; ++c) {
            const uint4 v = (own && c < nchunk) ? __ldg(src + c) : make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4 *>(sB + (size_t)(c >> 3) * kBlkBytes + tid * 128 + (((c & 7) ^ (tid & 7)) << 4)) = v;
        }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = make_idesc_bf16(kHeadTile, kHeadTile, 0, 0);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        mbar_wait(&bar_a, 0);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
            const int ksteps = min(kKB, p.K - kb * kKB) / 16;
            for (int j = 0; j < ksteps; ++j)
                umma_bf16(tmem, make_smem_desc_sw128(a0 + kb * kBlkBytes + j * 32), make_smem_desc_sw128(b0 + kb * kBlkBytes + j * 32),
                          idesc, (kb | j) != 0);
        }
        umma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    uint32_t acc[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(warp * 32), acc);    // the diagonal 32x32 block of this warp
    tmem_ld_wait();
    float d = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) d = (q == lane) ? __uint_as_float(acc[q]) : d;
    if (own) p.st[row] = d + (p.bias ? p.bias[tloc] : 0.f);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

// Merge S sorted candidate lists per row (descending, ties -> lower index first) and add the rank counts.
// One WARP per row: the S*k candidates are packed into 64-bit keys (order-preserving value bits | ~index) in shared memory;
// k rounds of "every lane scans its strided share for its best key, warp arg-max, the winner's slot is cleared".
constexpr int kMergeWarps = 4;
__global__ void __launch_bounds__(kMergeWarps * 32)
k5_topk_merge_kernel(const float *__restrict__ val, const int32_t *__restrict__ idx, const int32_t *__restrict__ cgt,
                     const int32_t *__restrict__ ceq, int M, int S, int k, float *__restrict__ oval,
                     int32_t *__restrict__ oidx, int32_t *__restrict__ rank) {
    extern __shared__ __align__(8) unsigned long long mkeys[];      // [kMergeWarps][S*k]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * kMergeWarps + warp;
    if (r >= M) return;
    const int n = S * k;
    unsigned long long *keys = mkeys + (size_t)warp * n;
    for (int i = lane; i < n; i += 32) {
        const size_t o = (size_t)r * n + i;
        const int id = idx[o];
        keys[i] = id < 0 ? 0ull : (((unsigned long long)f2ord(val[o]) << 32) | (uint32_t)(~(uint32_t)id));
    }
    __syncwarp();
    for (int j = 0; j < k; ++j) {
        unsigned long long best = 0ull;
        int bi = -1;
        for (int i = lane; i < n; i += 32) {
            const unsigned long long x = keys[i];
            if (x > best) { best = x; bi = i; }
        }
        unsigned long long wb = best;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long y = __shfl_xor_sync(0xffffffffu, wb, o);
            wb = y > wb ? y : wb;
        }
        if (wb != 0ull && best == wb) keys[bi] = 0ull;       // keys are unique (distinct vocabulary indices)
        if (lane == 0) {
            oval[(size_t)r * k + j] = wb ? ord2f((uint32_t)(wb >> 32)) : -INFINITY;
            oidx[(size_t)r * k + j] = wb ? (int32_t)(~(uint32_t)wb) : -1;
        }
        __syncwarp();
    }
    if (rank) {
        int rk = 0;
        for (int s = lane; s < S; s += 32) rk += cgt[(size_t)r * S + s] + ceq[(size_t)r * S + s];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rk += __shfl_xor_sync(0xffffffffu, rk, o);
        if (lane == 0) rank[r] = rk;
    }
}

static int32_t encode_rows_sw128(CUtensorMap *tm, const void *base, int rows, int K, int box_rows = kHeadTile) {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
    uint64_t str[1] = {(uint64_t)K * 2};
    uint32_t box[2] = {kKB, (uint32_t)box_rows};
    return encode_tmap_bf16(tm, base, 2, dims, str, box, 1);
}

}  // namespace mobgt

using namespace mobgt;

extern "C" int32_t mobgt_head_topk(const void *z, const void *W, const float *bias, const int32_t *target, int32_t M,
                                   int32_t V, int32_t K, int64_t vocab_offset, int32_t k, int32_t nsplit, int32_t mode,
                                   float *st, float *topk_val, int32_t *topk_idx, int32_t *cnt_gt, int32_t *cnt_eq,
                                   float *logits_dump, void *thr_share, void *stream) {
    MOBGT_REQUIRE(z && W && target && st, MOBGT_ERR_NULL, "mobgt_head_topk: null pointer");
    MOBGT_REQUIRE(mode == 0 || (topk_val && topk_idx && cnt_gt && cnt_eq), MOBGT_ERR_NULL, "mobgt_head_topk: null output");
    MOBGT_REQUIRE(K % 16 == 0 && K >= 16 && K <= kHeadMaxK, MOBGT_ERR_BAD_SHAPE, "mobgt_head_topk: K=%d", K);
    MOBGT_REQUIRE(k >= 1 && k <= kHeadMaxTop, MOBGT_ERR_BAD_SHAPE, "mobgt_head_topk: k=%d", k);
    MOBGT_REQUIRE(mode == 0 || (nsplit >= 2 && nsplit <= 160 && nsplit % 2 == 0), MOBGT_ERR_BAD_SHAPE,
                  "mobgt_head_topk: nsplit=%d must be even, in [2,160] (two epilogue groups per vocabulary split)", nsplit);
    MOBGT_REQUIRE(((uintptr_t)z & 15) == 0 && ((uintptr_t)W & 15) == 0 && (!bias || ((uintptr_t)bias & 15) == 0), MOBGT_ERR_BAD_SHAPE,
                  "mobgt_head_topk: z, W and bias must be 16-byte aligned");
    if (M <= 0 || V <= 0) return MOBGT_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    CUtensorMap tmZ, tmW;
    int32_t rc = encode_rows_sw128(&tmZ, z, M, K);
    if (rc) return rc;
    const int kblocks = ceil_div(K, kKB);
    if (mode == 0) {
        const size_t smem = (size_t)2 * kblocks * kBlkBytes + 1024;
        MOBGT_CUDA_OK(cudaFuncSetAttribute(k5_target_logit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TargetParams tp{static_cast<const __nv_bfloat16 *>(W), bias, target, st, M, V, K, vocab_offset};
        k5_target_logit_kernel<<<ceil_div(M, kHeadTile), 128, smem, s>>>(tmZ, tp);
        MOBGT_LAUNCH_OK("k5_target_logit_kernel");
        return MOBGT_OK;
    }
    // two adjacent row tiles share their W stream through a 2-CTA cluster (multicast halves) whenever the tile count is even
    const int cluster = (ceil_div(M, kHeadTile) % 2 == 0 && !g_head_no_cluster) ? 2 : 1;
    rc = encode_rows_sw128(&tmW, W, V, K, kNT / cluster);
    if (rc) return rc;
    const size_t fixed = (size_t)kblocks * kBlkBytes + (size_t)2 * kCandCap * kHeadTile * 8 + (size_t)4 * kNT * 4 + 1024;
    int ring = (int)((227 * 1024 - 512 - (long long)fixed) / kStageBytes);
    ring = ring > kMaxRing ? kMaxRing : ring;
    MOBGT_REQUIRE(ring >= 2, MOBGT_ERR_UNSUPPORTED, "mobgt_head_topk: no shared-memory plan for K=%d k=%d", K, k);
    const size_t smem = fixed + (size_t)ring * kStageBytes;
    HeadParams p{bias, target, st, topk_val, topk_idx, cnt_gt, cnt_eq, logits_dump, M, V, K, k, nsplit, mode, ring, g_timeline_dev,
                 static_cast<uint32_t *>(thr_share), vocab_offset};
    dim3 grid((unsigned)ceil_div(M, kHeadTile), (unsigned)(nsplit / 2));
    // the list length is a compile-time constant (register-resident list): the smallest built size >= k
    auto launch = [&](auto kern, const HeadParams &hp, dim3 grid) -> int32_t {
        MOBGT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid;
        cfg.blockDim = dim3(kHeadThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)cluster;       // two adjacent row tiles of the same vocabulary split
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        MOBGT_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tmZ, tmW, hp));
        return MOBGT_OK;
    };
    auto run = [&](const HeadParams &hp, dim3 grid) -> int32_t {
        if (cluster == 2) {
            if (k <= 10) return launch(k5_head_kernel<10, 2>, hp, grid);
            if (k <= 20) return launch(k5_head_kernel<20, 2>, hp, grid);
            return launch(k5_head_kernel<32, 2>, hp, grid);
        }
        if (k <= 10) return launch(k5_head_kernel<10, 1>, hp, grid);
        if (k <= 20) return launch(k5_head_kernel<20, 1>, hp, grid);
        return launch(k5_head_kernel<32, 1>, hp, grid);
    };
    rc = run(p, grid);
    if (rc) return rc;
    MOBGT_LAUNCH_OK("k5_head_kernel");
    return MOBGT_OK;
}

// Debug / measurement switches of mobgt_head_topk; 1 = defaults.
extern "C" int32_t mobgt_debug_head_cluster(int32_t on) {
    mobgt::g_head_no_cluster = (on & 1) == 0;      // bit 0: pair row tiles into clusters (1 = default)
    return MOBGT_OK;
}

extern "C" int32_t mobgt_topk_merge(const float *val, const int32_t *idx, const int32_t *cnt_gt, const int32_t *cnt_eq,
                                    int32_t M, int32_t S, int32_t k, float *out_val, int32_t *out_idx, int32_t *rank,
                                    void *stream) {
    MOBGT_REQUIRE(val && idx && out_val && out_idx, MOBGT_ERR_NULL, "mobgt_topk_merge: null pointer");
    MOBGT_REQUIRE(!rank || (cnt_gt && cnt_eq), MOBGT_ERR_NULL, "mobgt_topk_merge: rank needs the counts");
    MOBGT_REQUIRE(S >= 1 && S <= 160 && k >= 1 && k <= kHeadMaxTop, MOBGT_ERR_BAD_SHAPE, "mobgt_topk_merge: S=%d k=%d", S, k);
    if (M <= 0) return MOBGT_OK;
    const size_t smem = (size_t)kMergeWarps * S * k * sizeof(unsigned long long);
    MOBGT_CUDA_OK(cudaFuncSetAttribute(k5_topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k5_topk_merge_kernel<<<ceil_div(M, kMergeWarps), kMergeWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
        val, idx, cnt_gt, cnt_eq, M, S, k, out_val, out_idx, rank);
    MOBGT_LAUNCH_OK("k5_topk_merge_kernel");
    return MOBGT_OK;
}
