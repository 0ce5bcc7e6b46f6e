// K3 (backward) — biased multi-head attention backward on tcgen05 / TMEM, fed by TMA (sm_100a).
//
// Gradient of model_fqandtoyo.py:1693-1706 for packed var-len graphs (see k3_attn_fwd.cu):
//   P  = exp(scale * Q K^T + bias - lse)            recomputed per tile, never stored in HBM
//   dP = dO V^T ;  D = rowsum(dO * O) ;  dS = P * (dP - D)          (dS is also d(bias))
//   dV = P^T dO ;  dK = scale * dS^T Q ;  dQ = scale * dS K
// One CTA per (graph, head).  Five MMAs per (kv block j, query tile i):
//   S, dP    M=128 (queries) N<=128 K=32           -> TMEM [0,128), [128,256)
//   dV, dK   M=128 (keys)    N=32   K=128 queries  -> TMEM [256,288), [288,320)   (accumulate over i)
//   dQ_i     M=128 (queries) N=32   K<=128 keys    -> TMEM [320+32 i, ...)        (accumulate over j)
// P and dS are written once to shared memory as bf16 in the [chunk][row][8] layout; that single image is the
// K-major A operand of dQ = dS.K and the MN-major (transposed) A operand of dV = P^T.dO / dK = dS^T.Q.
// dS is streamed to the fp32 dBias buffer [B,H,T,Tp] (overwrite or accumulate across layers); the table gradients
// are reduced from it by mobgt_bias_bwd.
#include <cuda_bf16.h>

#include "common.cuh"
#include "k3_small.cuh"
#include "umma.cuh"

namespace mobgt {
using namespace sm100;

namespace bwd {
constexpr int kAttD = 24;
constexpr int kAttChunks = 3;
constexpr int kTile = 128;
constexpr int kBoxBytes = 4 * kTile * 16;
constexpr int kBoxTxBytes = kAttChunks * kTile * 16;
constexpr int kBiasTileBytes = 2 * kTile * 128;
constexpr int kPBytes = 16 * kTile * 16;
constexpr int kMaxTiles = 5;   // T <= 513
}  // namespace bwd
using namespace bwd;

struct AttnBwdParams {
    const int32_t *tok_off;
    const __nv_bfloat16 *o;    // [ntok, H*24]
    const __nv_bfloat16 *dout; // [ntok, H*24]
    const float *lse;          // [ntok, H]
    __nv_bfloat16 *dq, *dk, *dv;  // [ntok, *] with row stride dqkv_stride (elements)
    int64_t dqkv_stride;
    float *dbias;              // [B,H,T,Tp] f32 (modes 0, 1); mode 2 writes bf16 through the tmDS tensor map
    int H, T, Tp;
    float scale;
    int max_boxes;
    int accumulate;            // 0: f32 overwrite, 1: f32 dbias += dS, 2: bf16 overwrite (per-layer plane, TMA store)
    int bias_bufs;             // 1 or 2 bias tiles in shared memory
    long long *timeline;       // debug (mobgt_debug_set_timeline) or NULL
    AttnDrop drop;             // the forward's attention dropout (same seed): th16 == 0 when it was off
    const __nv_bfloat16 *q, *k, *v;   // raw views of the operands (row stride qkv_stride) for the single-token tail
    int64_t qkv_stride;
    const __nv_bfloat16 *bias;        // [B,H,T,Tp]
    int small_t;                      // graphs of at most this many tokens belong to the SIMT kernel (k3_attn_small.cu); 0: none
    const int32_t *order;             // [B] graph ids in launch order (descending size) or NULL
};

// 24 bf16 (three 16-byte words) -> fp32
__device__ __forceinline__ void unpack24(const uint4 &a, const uint4 &b, const uint4 &c, float (&f)[24]) {
    const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
#pragma unroll
    for (int e = 0; e < 12; ++e) {
        f[2 * e] = __uint_as_float(w[e] << 16);
        f[2 * e + 1] = __uint_as_float(w[e] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ float bf16_at(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

__device__ __forceinline__ float bwd_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// d(bias) += v at 4 consecutive floats: one fire-and-forget vector reduction (REDG.E.ADD.F32x4) — no load round trip
// in the tile loop.  Every dbias element is touched by exactly one thread per launch, launches are stream-ordered, so
// the layer sum is still deterministic.
__device__ __forceinline__ void red_add_f32x4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// 256 threads = two warpgroups.  Thread (wg, t128) owns query row t128 of the current tile (TMEM lane t128; warps w and
// w+4 may both access lanes 32*(w%4)..+31) and the 16-column chunks c0 = 16*wg, 16*wg + 32, ... of the score tile.
// Q / dO (and, when it fits, the bias tile) are double-buffered: the TMA loads of iteration t+1 are issued before the
// S / dP MMAs of iteration t, so their latency hides behind the softmax-gradient math.
// kDrop (training-mode attention dropout, forward: O = (M o P / (1-p)) V with keep mask M regenerated here):
//   dV = (M o P / (1-p))^T dO ;  dP = M o (dO V^T) / (1-p) ;  dS = P o (dP - D) ;  D = rowsum(dO o O) still holds.
//
// Single-token tail ("fold"): a graph whose token count is 128 m + 1 (every graph at a node cap of 128 m: n nodes + the graph
// token) would spill ONE query row and ONE key column into a second tile in each dimension, i.e. three extra near-empty
// tile iterations per (j, i) border that cost as much latency as full ones.  For such graphs the MMA loop runs over the m x m
// full tiles only and the last token sp = 128 m is handled by plain SIMT math (lse is known, so every element is independent):
//   row sp against all keys, all rows against key sp -> dS / P scalars in shared memory, their rank-1 contributions are added
//   to the dK / dV / dQ accumulators in the epilogues, and dQ[sp], dK[sp], dV[sp] are three small mat-vec reductions.
template <bool kDrop>
__global__ void __launch_bounds__(256, 1)
k3_attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmBias, const __grid_constant__ CUtensorMap tmDS,
                   const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_qdo[2], bar_bias[2], bar_kv, bar_s, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float sLse[kMaxTiles * kTile], sDelta[kMaxTiles * kTile];
    __shared__ float sA_ds[kMaxTiles * kTile], sA_pd[kMaxTiles * kTile];   // tail: dS / dropped P of (row sp, key c)
    __shared__ float sB_ds[kMaxTiles * kTile], sB_pd[kMaxTiles * kTile];   // tail: dS / dropped P of (row r, key sp)
    __shared__ float sSp[4][kAttD];                                        // q, k, v, dO of token sp
    __shared__ float sRed[3][10][kAttD];

    MOBGT_STAMP(p.timeline, 0);
    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, t128 = tid & 127;
    const bool warp0 = warp_index_uniform() == 0;   // the issuing warp (one elected lane issues TMA / MMA)
    const int gi = blockIdx.x / p.H, h = blockIdx.x - gi * p.H;
    const int g = p.order ? p.order[gi] : gi;       // largest graphs first: the CTAs that return at once are dispatched behind them
    const int t0 = p.tok_off[g];
    const int Tg = p.tok_off[g + 1] - t0;
    if (Tg <= p.small_t) return;                           // the whole CTA: nothing has been set up yet
    const bool fold = Tg > kTile && (Tg % kTile) == 1;     // single-token tail handled by SIMT (see above)
    const int NB = fold ? Tg / kTile : ceil_div(Tg, kTile);
    const int sp = Tg - 1;                                 // the tail token (fold only)
    const int NT = NB * NB;
    const int HD = p.H * kAttD;
    const int nbias = p.bias_bufs;

    // Global loads issued at the very top, consumed after the barrier / TMA set-up (their latency hides behind it):
    //   this thread's first row of the lse / delta prologue, and (fold) the bias row / column of the tail token.
    const int rows_pro = NB * kTile + (fold ? 1 : 0);
    float lse_pre = 0.f;
    uint4 o_pre[3], d_pre[3];
    if (tid < Tg && tid < rows_pro) {
        lse_pre = p.lse[(size_t)(t0 + tid) * p.H + h];
        const uint4 *po = reinterpret_cast<const uint4 *>(p.o + (size_t)(t0 + tid) * HD + h * kAttD);
        const uint4 *pd = reinterpret_cast<const uint4 *>(p.dout + (size_t)(t0 + tid) * HD + h * kAttD);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            o_pre[q] = po[q];
            d_pre[q] = pd[q];
        }
    }
    float b_row[3] = {0.f, 0.f, 0.f}, b_col[2] = {0.f, 0.f};
    if (fold) {
        const __nv_bfloat16 *bias_pl0 = p.bias + (size_t)(g * p.H + h) * p.T * p.Tp;
#pragma unroll
        for (int u = 0; u < 3; ++u)
            if (tid + u * 256 < Tg) b_row[u] = bf16_at(bias_pl0 + (size_t)sp * p.Tp + tid + u * 256);
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (tid + u * 256 < sp) b_col[u] = bf16_at(bias_pl0 + (size_t)(tid + u * 256) * p.Tp + sp);
    }
    uint8_t *sBias = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // nbias x 32 KB, 1024-aligned (swizzle atom)
    uint8_t *sP = sBias + (size_t)nbias * kBiasTileBytes;
    uint8_t *sdS = sP + kPBytes;
    uint8_t *sQ = sdS + kPBytes;            // 2 x 8 KB
    uint8_t *sdO = sQ + 2 * kBoxBytes;      // 2 x 8 KB
    uint8_t *sK = sdO + 2 * kBoxBytes;
    uint8_t *sV = sK + (size_t)p.max_boxes * kBoxBytes;

    if (tid == 0) {
        mbar_init(&bar_qdo[0], 1);
        mbar_init(&bar_qdo[1], 1);
        mbar_init(&bar_bias[0], 1);
        mbar_init(&bar_bias[1], 1);
        mbar_init(&bar_kv, 1);
        mbar_init(&bar_s, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmdO);
        tma_prefetch_desc(&tmBias);
        if (p.accumulate == 2) tma_prefetch_desc(&tmDS);
        // first loads right away (this thread initialised the barriers): they fly during the lse / delta prologue below
        const int plane0 = g * p.H + h;
        mbar_expect_tx(&bar_kv, (uint32_t)(2 * NB * kBoxTxBytes));
        for (int b = 0; b < NB; ++b) {
            tma_load_3d(sK + (size_t)b * kBoxBytes, &tmK, &bar_kv, 0, t0 + b * kTile, h * kAttChunks);
            tma_load_3d(sV + (size_t)b * kBoxBytes, &tmV, &bar_kv, 0, t0 + b * kTile, h * kAttChunks);
        }
        mbar_expect_tx(&bar_qdo[0], 2 * kBoxTxBytes);
        tma_load_3d(sQ, &tmQ, &bar_qdo[0], 0, t0, h * kAttChunks);
        tma_load_3d(sdO, &tmdO, &bar_qdo[0], 0, t0, h * kAttChunks);
        mbar_expect_tx(&bar_bias[0], kBiasTileBytes);
        tma_load_3d(sBias, &tmBias, &bar_bias[0], 0, 0, plane0);
        tma_load_3d(sBias + kTile * 128, &tmBias, &bar_bias[0], 64, 0, plane0);
    }
    {   // zero the K-padding chunk (d = 24..31) of every operand box: 128 rows x 16 B each
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint8_t *qd = wg == 0 ? sQ : sdO;
        *reinterpret_cast<uint4 *>(qd + 3 * kTile * 16 + t128 * 16) = z;
        *reinterpret_cast<uint4 *>(qd + kBoxBytes + 3 * kTile * 16 + t128 * 16) = z;
        uint8_t *kv = wg == 0 ? sK : sV;
        for (int b = 0; b < NB; ++b) *reinterpret_cast<uint4 *>(kv + (size_t)b * kBoxBytes + 3 * kTile * 16 + t128 * 16) = z;
    }
    // lse (log2 units) and D = rowsum(dO * O) of every query row of this (graph, head).  Rows past the graph get
    // lse = +inf, so that p = 2^(s - lse) = 0 and dS = 0 there without any per-element select.
    float sp_val = 0.f;
    if (fold && tid < 4 * kAttD) {   // q, k, v, dO of the tail token as fp32 (load issued here, stored after the loop below)
        const int which = tid / kAttD, e = tid - which * kAttD;
        const __nv_bfloat16 *src = which == 0 ? p.q : which == 1 ? p.k : which == 2 ? p.v : p.dout;
        const int64_t stride = which == 3 ? (int64_t)HD : p.qkv_stride;
        sp_val = bf16_at(src + (size_t)(t0 + sp) * stride + h * kAttD + e);
    }
    for (int r = tid; r < rows_pro; r += 256) {
        float lse2 = INFINITY, dl = 0.f;
        if (r < Tg) {
            const bool pre = r == tid;      // first row: loaded at the top of the kernel
            lse2 = (pre ? lse_pre : p.lse[(size_t)(t0 + r) * p.H + h]) * 1.4426950408889634f;
            const uint4 *po = reinterpret_cast<const uint4 *>(p.o + (size_t)(t0 + r) * HD + h * kAttD);
            const uint4 *pd = reinterpret_cast<const uint4 *>(p.dout + (size_t)(t0 + r) * HD + h * kAttD);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const uint4 a = pre ? o_pre[q] : po[q], b = pre ? d_pre[q] : pd[q];
                const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    dl += __uint_as_float(aw[e] << 16) * __uint_as_float(bw[e] << 16);
                    dl += __uint_as_float(aw[e] & 0xFFFF0000u) * __uint_as_float(bw[e] & 0xFFFF0000u);
                }
            }
        }
        sLse[r] = lse2;
        sDelta[r] = dl;
    }
    if (fold && tid < 4 * kAttD) sSp[tid / kAttD][tid % kAttD] = sp_val;
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    MOBGT_STAMP(p.timeline, 1);
    const uint32_t tmem = tmem_slot;
    const uint32_t tS = tmem, tdP = tmem + 128, tdK = tmem + 256, tdV = tmem + 288, tdQ = tmem + 320;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int plane = g * p.H + h;

    // TMA loads of flattened iteration t = j * NB + i (thread 0 only)
    auto load_qdo = [&](int t) {
        const int i = t % NB, b = t & 1;
        mbar_expect_tx(&bar_qdo[b], 2 * kBoxTxBytes);
        tma_load_3d(sQ + b * kBoxBytes, &tmQ, &bar_qdo[b], 0, t0 + i * kTile, h * kAttChunks);
        tma_load_3d(sdO + b * kBoxBytes, &tmdO, &bar_qdo[b], 0, t0 + i * kTile, h * kAttChunks);
    };
    auto load_bias = [&](int t) {
        const int j = t / NB, i = t % NB, b = (nbias == 2) ? (t & 1) : 0;
        uint8_t *dst = sBias + (size_t)b * kBiasTileBytes;
        mbar_expect_tx(&bar_bias[b], kBiasTileBytes);
        tma_load_3d(dst, &tmBias, &bar_bias[b], j * kTile, i * kTile, plane);
        tma_load_3d(dst + kTile * 128, &tmBias, &bar_bias[b], j * kTile + 64, i * kTile, plane);
    };
    // S = Q_i K_j^T and dP = dO_i V_j^T of flattened iteration t into TMEM (thread 0 only)
    auto issue_sdp = [&](int t) {
        const int j = t / NB, b = t & 1;
        mbar_wait(&bar_qdo[b], (t >> 1) & 1);
        tc_fence_after();
        const int nbj = round_up(min(kTile, Tg - j * kTile), 16);
        const uint32_t idesc = make_idesc_bf16(kTile, nbj, 0, 0);
        const uint32_t aq = smem_u32(sQ + b * kBoxBytes), ado = smem_u32(sdO + b * kBoxBytes);
        const uint32_t bk = smem_u32(sK + (size_t)j * kBoxBytes), bv = smem_u32(sV + (size_t)j * kBoxBytes);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            umma_bf16(tS, make_smem_desc(aq + ks * 2 * kTile * 16, kTile * 16, 128),
                      make_smem_desc(bk + ks * 2 * kTile * 16, kTile * 16, 128), idesc, ks > 0);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            umma_bf16(tdP, make_smem_desc(ado + ks * 2 * kTile * 16, kTile * 16, 128),
                      make_smem_desc(bv + ks * 2 * kTile * 16, kTile * 16, 128), idesc, ks > 0);
        umma_commit(&bar_s);
    };
    if (warp0 && elect_one()) {
        mbar_wait(&bar_kv, 0);
        MOBGT_STAMP(p.timeline, 2);
        issue_sdp(0);
        MOBGT_STAMP(p.timeline, 3);
    }
    __syncwarp();
    uint32_t ph_s = 0, ph_mma = 0;
    const float sl2 = p.scale * 1.4426950408889634f;
    constexpr float kL2e = 1.4426950408889634f;
    bool mma_pending = false;
    uint32_t seed_lo = 0, seed_hi = 0;
    if (kDrop) attn_drop_fold_seed(p.drop, seed_lo, seed_hi);
    const uint32_t th_hi = p.drop.th16 << 16;

    if (fold) {   // ---- the single-token tail, SIMT (overlaps the first S / dP MMAs)
        const float ik = p.drop.inv_keep;
        const __nv_bfloat16 *bias_pl = p.bias + (size_t)plane * p.T * p.Tp;
        __nv_bfloat16 *ds16 = reinterpret_cast<__nv_bfloat16 *>(p.dbias) + (size_t)plane * p.T * p.Tp;   // mode 2
        float *ds32 = p.dbias + (size_t)plane * p.T * p.Tp;                                              // modes 0 / 1
        auto emit_ds = [&](int r, int c, float v) {
            const size_t o = (size_t)r * p.Tp + c;
            if (p.accumulate == 2) ds16[o] = __float2bfloat16_rn(v);
            else if (p.accumulate == 1) atomicAdd(ds32 + o, v);
            else ds32[o] = v;
        };
        // 24 bf16 of row `row` of a [chunk][row][8] operand box
        auto box_row = [&](const uint8_t *box, int row, float (&f)[24]) {
            const uint8_t *b0 = box + row * 16;
            unpack24(*reinterpret_cast<const uint4 *>(b0), *reinterpret_cast<const uint4 *>(b0 + kTile * 16),
                     *reinterpret_cast<const uint4 *>(b0 + 2 * kTile * 16), f);
        };
        mbar_wait(&bar_kv, 0);                      // K / V boxes and the first Q / dO tile have landed
        mbar_wait(&bar_qdo[0], 0);                  // (every thread observes the phases; tile 0 stays put until iteration 1)
        {   // part A: query row sp against every key c (itself included)
            const float lse_s = sLse[sp], delta_s = sDelta[sp];
            const uint32_t rk = kDrop ? attn_drop_rowkey((uint32_t)plane, (uint32_t)sp, seed_lo, seed_hi) : 0u;
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int c = tid + u * 256;
                if (c >= Tg) break;
                float dot = 0.f, dp = 0.f;
                if (c < sp) {
                    float kf[24], vf[24];
                    box_row(sK + (size_t)(c >> 7) * kBoxBytes, c & 127, kf);
                    box_row(sV + (size_t)(c >> 7) * kBoxBytes, c & 127, vf);
#pragma unroll
                    for (int e = 0; e < kAttD; ++e) {
                        dot = fmaf(sSp[0][e], kf[e], dot);
                        dp = fmaf(sSp[3][e], vf[e], dp);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < kAttD; ++e) {
                        dot = fmaf(sSp[0][e], sSp[1][e], dot);
                        dp = fmaf(sSp[3][e], sSp[2][e], dp);
                    }
                }
                const float pr = bwd_exp2(fmaf(dot, sl2, b_row[u] * kL2e) - lse_s);
                bool kp = true;
                if (kDrop) kp = (attn_drop_keep8(rk, (uint32_t)(c >> 3), p.drop.th16) >> (c & 7)) & 1u;
                const float ds = pr * ((kDrop ? (kp ? dp * ik : 0.f) : dp) - delta_s);
                sA_ds[c] = ds;
                sA_pd[c] = kDrop ? (kp ? pr * ik : 0.f) : pr;
                emit_ds(sp, c, ds);
            }
        }
        {   // part B: every full-tile query row r against key sp (rows of tile 0 from shared memory, the others from global)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int r = tid + u * 256;
                if (r >= sp) break;
                float qf[24], df[24];
                if (r < kTile) {
                    box_row(sQ, r, qf);
                    box_row(sdO, r, df);
                } else {
                    const uint4 *qg = reinterpret_cast<const uint4 *>(p.q + (size_t)(t0 + r) * p.qkv_stride + h * kAttD);
                    const uint4 *dg = reinterpret_cast<const uint4 *>(p.dout + (size_t)(t0 + r) * HD + h * kAttD);
                    unpack24(qg[0], qg[1], qg[2], qf);
                    unpack24(dg[0], dg[1], dg[2], df);
                }
                float dot = 0.f, dp = 0.f;
#pragma unroll
                for (int e = 0; e < kAttD; ++e) {
                    dot = fmaf(qf[e], sSp[1][e], dot);
                    dp = fmaf(df[e], sSp[2][e], dp);
                }
                const float pr = bwd_exp2(fmaf(dot, sl2, b_col[u] * kL2e) - sLse[r]);
                bool kp = true;
                if (kDrop) {
                    const uint32_t rk = attn_drop_rowkey((uint32_t)plane, (uint32_t)r, seed_lo, seed_hi);
                    kp = (attn_drop_keep8(rk, (uint32_t)(sp >> 3), p.drop.th16) >> (sp & 7)) & 1u;
                }
                const float ds = pr * ((kDrop ? (kp ? dp * ik : 0.f) : dp) - sDelta[r]);
                sB_ds[r] = ds;
                sB_pd[r] = kDrop ? (kp ? pr * ik : 0.f) : pr;
                emit_ds(r, sp, ds);
            }
        }
        __syncthreads();
        if (tid < 240) {   // dQ[sp] = sum_c dS[sp][c] k_c ; dK[sp] = sum_r dS[r][sp] q_r ; dV[sp] = sum_r P[r][sp] dO_r : thread = (e, seg)
            const int seg = tid / kAttD, e = tid - seg * kAttD;
            const int eo = (e >> 3) * (kTile * 16) + (e & 7) * 2;      // byte offset of element e inside a box row
            float aq = 0.f, ak = 0.f, av = 0.f;
            for (int c = seg; c < sp; c += 10)                         // keys: every box is resident
                aq = fmaf(sA_ds[c], bf16_at(reinterpret_cast<const __nv_bfloat16 *>(
                                        sK + (size_t)(c >> 7) * kBoxBytes + (c & 127) * 16 + eo)), aq);
            for (int r = seg; r < kTile; r += 10) {                    // query rows of tile 0: shared memory
                ak = fmaf(sB_ds[r], bf16_at(reinterpret_cast<const __nv_bfloat16 *>(sQ + r * 16 + eo)), ak);
                av = fmaf(sB_pd[r], bf16_at(reinterpret_cast<const __nv_bfloat16 *>(sdO + r * 16 + eo)), av);
            }
            for (int r0 = kTile + seg; r0 < sp; r0 += 80) {            // further tiles: global, 8 rows in flight
                float qv8[8], dv8[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int r = r0 + 10 * u;
                    qv8[u] = r < sp ? bf16_at(p.q + (size_t)(t0 + r) * p.qkv_stride + h * kAttD + e) : 0.f;
                    dv8[u] = r < sp ? bf16_at(p.dout + (size_t)(t0 + r) * HD + h * kAttD + e) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int r = r0 + 10 * u;
                    if (r < sp) {
                        ak = fmaf(sB_ds[r], qv8[u], ak);
                        av = fmaf(sB_pd[r], dv8[u], av);
                    }
                }
            }
            sRed[0][seg][e] = aq;
            sRed[1][seg][e] = ak;
            sRed[2][seg][e] = av;
        }
        __syncthreads();
        if (tid < 3 * kAttD) {
            const int which = tid / kAttD, e = tid - which * kAttD;
            float a = 0.f;
#pragma unroll
            for (int sg = 0; sg < 10; ++sg) a += sRed[which][sg][e];
            // the (sp, sp) element: dQ += dS k_sp ; dK += dS q_sp ; dV += P dO_sp
            a += which == 0 ? sA_ds[sp] * sSp[1][e] : which == 1 ? sA_ds[sp] * sSp[0][e] : sA_pd[sp] * sSp[3][e];
            if (which < 2) a *= p.scale;
            __nv_bfloat16 *dst = (which == 0 ? p.dq : which == 1 ? p.dk : p.dv) + (size_t)(t0 + sp) * p.dqkv_stride + h * kAttD + e;
            *dst = __float2bfloat16_rn(a);
        }
    }

    for (int j = 0; j < NB; ++j) {
        const int kv_valid = min(kTile, Tg - j * kTile);
        const int nb = round_up(kv_valid, 16);
        for (int i = 0; i < NB; ++i) {
            const int t = j * NB + i, buf = t & 1;
            const int q_valid = min(kTile, Tg - i * kTile);
            const int qk = round_up(q_valid, 16);          // K extent (query rows) of the dV / dK MMAs
            const int row = i * kTile + t128;
            const bool row_ok = row < Tg;
            const bool warp_live = i * kTile + (warp & 3) * 32 < Tg;   // any valid query row in this warp's 32 lanes?
            if (mma_pending) {   // the previous iteration's dV/dK/dQ MMAs read sP, sdS and the other Q / dO buffer
                mbar_wait(&bar_mma, ph_mma);
                ph_mma ^= 1;
                tc_fence_after();
            }
            MOBGT_STAMP(p.timeline, 8 + 8 * t + 0);   // after bar_mma wait
            if (warp0 && elect_one()) {
                if (t + 1 < NT) {
                    load_qdo(t + 1);
                    if (nbias == 2) load_bias(t + 1);
                }
            }
            __syncwarp();
            mbar_wait(&bar_s, ph_s);
            ph_s ^= 1;
            MOBGT_STAMP(p.timeline, 8 + 8 * t + 1);   // S / dP ready
            const int bb = (nbias == 2) ? buf : 0;
            mbar_wait(&bar_bias[bb], (nbias == 2) ? ((t >> 1) & 1) : (t & 1));
            tc_fence_after();
            MOBGT_STAMP(p.timeline, 8 + 8 * t + 2);   // bias tile ready

            if (warp_live) {
                const uint8_t *sB = sBias + (size_t)bb * kBiasTileBytes;
                const float lse2 = sLse[row];
                const float delta = sDelta[row];
                const uint32_t rowkey = kDrop ? attn_drop_rowkey((uint32_t)plane, (uint32_t)row, seed_lo, seed_hi) : 0u;
                const float ik = p.drop.inv_keep;
                float *db_row = p.dbias + ((size_t)plane * p.T + row) * p.Tp + j * kTile;
                uint32_t svv[4][16], dpvv[4][16];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)     // every TMEM load of this thread's columns in flight before the one wait
                    if (wg * 16 + cc * 32 < nb) {
                        tmem_ld16(tS + lane_off + wg * 16 + cc * 32, svv[cc]);
                        tmem_ld16(tdP + lane_off + wg * 16 + cc * 32, dpvv[cc]);
                    }
                tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c0 = wg * 16 + cc * 32;
                    if (c0 >= nb) continue;
                    const uint32_t (&sv)[16] = svv[cc];
                    const uint32_t (&dpv)[16] = dpvv[cc];
#pragma unroll
                    for (int q8 = 0; q8 < 2; ++q8) {
                        const int c8 = (c0 >> 3) + q8;
                        const int colb = c8 * 8;
                        const uint8_t *bp = sB + (c8 >> 3) * (kTile * 128) + t128 * 128 + (((c8 & 7) ^ (t128 & 7)) << 4);
                        uint4 bvv = *reinterpret_cast<const uint4 *>(bp);
                        if (!row_ok) bvv = make_uint4(0, 0, 0, 0);   // bias rows past the graph are never written: not numbers
                        const uint32_t bw[4] = {bvv.x, bvv.y, bvv.z, bvv.w};
                        float dsv[8];
                        uint4 pk, dk;
                        AttnDropWords dw;
                        if (kDrop) dw = attn_drop_words(rowkey, (uint32_t)(j * (kTile / 8) + c8));
                        if (colb + 8 <= kv_valid) {   // all 8 key columns valid: no per-element masking
                            float pv[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float bias = __uint_as_float((e & 1) ? (bw[e >> 1] & 0xFFFF0000u) : (bw[e >> 1] << 16));
                                const float s = fmaf(__uint_as_float(sv[q8 * 8 + e]), sl2, bias * kL2e);
                                pv[e] = bwd_exp2(s - lse2);
                                if (kDrop) {
                                    const bool kp = attn_drop_keep(dw, e, th_hi);
                                    dsv[e] = pv[e] * ((kp ? __uint_as_float(dpv[q8 * 8 + e]) * ik : 0.f) - delta);
                                    pv[e] = kp ? pv[e] * ik : 0.f;      // the dV operand is the dropped-out P
                                } else {
                                    dsv[e] = pv[e] * (__uint_as_float(dpv[q8 * 8 + e]) - delta);
                                }
                            }
                            pk.x = pack_bf16(pv[0], pv[1]); pk.y = pack_bf16(pv[2], pv[3]);
                            pk.z = pack_bf16(pv[4], pv[5]); pk.w = pack_bf16(pv[6], pv[7]);
                            if (row_ok && p.accumulate != 2) {   // d(bias) = dS, fp32, this thread's row
                                if (p.accumulate) {
                                    red_add_f32x4(db_row + colb, dsv[0], dsv[1], dsv[2], dsv[3]);
                                    red_add_f32x4(db_row + colb + 4, dsv[4], dsv[5], dsv[6], dsv[7]);
                                } else {
                                    float4 *d4 = reinterpret_cast<float4 *>(db_row + colb);
                                    d4[0] = make_float4(dsv[0], dsv[1], dsv[2], dsv[3]);
                                    d4[1] = make_float4(dsv[4], dsv[5], dsv[6], dsv[7]);
                                }
                            }
                        } else {                      // ragged edge of the graph (bias padding columns are never read as numbers)
                            float pv[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const bool ok = row_ok && (colb + e < kv_valid);
                                const float bias = __uint_as_float((e & 1) ? (bw[e >> 1] & 0xFFFF0000u) : (bw[e >> 1] << 16));
                                const float s = fmaf(__uint_as_float(sv[q8 * 8 + e]), sl2, bias * kL2e);
                                pv[e] = ok ? bwd_exp2(s - lse2) : 0.f;
                                if (kDrop) {
                                    const bool kp = attn_drop_keep(dw, e, th_hi);
                                    dsv[e] = ok ? pv[e] * ((kp ? __uint_as_float(dpv[q8 * 8 + e]) * ik : 0.f) - delta) : 0.f;
                                    pv[e] = kp ? pv[e] * ik : 0.f;
                                } else {
                                    dsv[e] = ok ? pv[e] * (__uint_as_float(dpv[q8 * 8 + e]) - delta) : 0.f;
                                }
                            }
                            pk.x = pack_bf16(pv[0], pv[1]); pk.y = pack_bf16(pv[2], pv[3]);
                            pk.z = pack_bf16(pv[4], pv[5]); pk.w = pack_bf16(pv[6], pv[7]);
                            if (row_ok && p.accumulate != 2) {
#pragma unroll
                                for (int e = 0; e < 8; ++e)
                                    if (colb + e < kv_valid) {
                                        if (p.accumulate) atomicAdd(db_row + colb + e, dsv[e]);
                                        else db_row[colb + e] = dsv[e];
                                    }
                            }
                        }
                        dk.x = pack_bf16(dsv[0], dsv[1]); dk.y = pack_bf16(dsv[2], dsv[3]);
                        dk.z = pack_bf16(dsv[4], dsv[5]); dk.w = pack_bf16(dsv[6], dsv[7]);
                        *reinterpret_cast<uint4 *>(sP + c8 * (kTile * 16) + t128 * 16) = pk;
                        *reinterpret_cast<uint4 *>(sdS + c8 * (kTile * 16) + t128 * 16) = dk;
                    }
                }
            }
            MOBGT_STAMP(p.timeline, 8 + 8 * t + 3);   // math done (thread 0)
            fence_proxy_async_smem();
            tc_fence_before();
            __syncthreads();
            MOBGT_STAMP(p.timeline, 8 + 8 * t + 4);   // barrier passed
            if (warp0 && elect_one()) {
                tc_fence_after();
                if (nbias == 1 && t + 1 < NT) load_bias(t + 1);   // single bias buffer: free only now
                if (p.accumulate == 2) {   // this layer's dS plane (bf16): one TMA store straight from the MMA operand image
                    tma_store_4d(&tmDS, sdS, 0, i * kTile, j * (kTile / 8), plane);
                    tma_store_commit();
                }
                const uint32_t aP = smem_u32(sP), aS = smem_u32(sdS);
                const uint32_t bQ = smem_u32(sQ + buf * kBoxBytes), bdO = smem_u32(sdO + buf * kBoxBytes);
                const uint32_t bK = smem_u32(sK + (size_t)j * kBoxBytes);
                const uint32_t id_t = make_idesc_bf16(kTile, 32, 1, 1);   // A = P^T / dS^T (MN-major), B MN-major
                const uint32_t id_q = make_idesc_bf16(kTile, 32, 0, 1);   // A = dS (K-major),        B MN-major
                for (int ks = 0; ks < qk / 16; ++ks)      // dV_j += P^T dO_i   (K = query rows)
                    umma_bf16(tdV, make_smem_desc(aP + ks * 256, 128, kTile * 16),
                              make_smem_desc(bdO + ks * 256, 128, kTile * 16), id_t, (i > 0 || ks > 0));
                for (int ks = 0; ks < qk / 16; ++ks)      // dK_j += dS^T Q_i
                    umma_bf16(tdK, make_smem_desc(aS + ks * 256, 128, kTile * 16),
                              make_smem_desc(bQ + ks * 256, 128, kTile * 16), id_t, (i > 0 || ks > 0));
                for (int ks = 0; ks < nb / 16; ++ks)      // dQ_i += dS K_j     (K = key rows)
                    umma_bf16(tdQ + 32 * i, make_smem_desc(aS + ks * 2 * kTile * 16, kTile * 16, 128),
                              make_smem_desc(bK + ks * 256, 128, kTile * 16), id_q, (j > 0 || ks > 0));
                umma_commit(&bar_mma);
                // S / dP of the next iteration right behind: every thread has finished reading this one's from TMEM
                MOBGT_STAMP(p.timeline, 8 + 8 * t + 5);   // dV/dK/dQ issued
                if (t + 1 < NT) issue_sdp(t + 1);
                MOBGT_STAMP(p.timeline, 8 + 8 * t + 6);   // next S/dP issued (includes the Q/dO TMA wait)
            }
            __syncwarp();
            mma_pending = true;
        }
        // ---- dK_j (warpgroup 0) and dV_j (warpgroup 1) complete: TMEM -> bf16 -> global (key row = j*128 + t128)
        mbar_wait(&bar_mma, ph_mma);
        ph_mma ^= 1;
        tc_fence_after();
        mma_pending = false;   // the wait above already covered the last iteration's MMAs
        MOBGT_STAMP(p.timeline, 8 + 8 * (j * NB + NB - 1) + 7);   // block's dK/dV complete
        {
            const int krow = j * kTile + t128;
            if (j * kTile + (warp & 3) * 32 < Tg) {
                uint32_t kvv[32];
                tmem_ld32((wg == 0 ? tdK : tdV) + lane_off, kvv);
                tmem_ld_wait();
                if (krow < Tg) {
                    const float sc = wg == 0 ? p.scale : 1.0f;
                    if (fold) {   // + dS[sp][krow] q_sp (dK) / P[sp][krow] dO_sp (dV)
                        const float a = wg == 0 ? sA_ds[krow] : sA_pd[krow];
                        const float *vec = wg == 0 ? sSp[0] : sSp[3];
#pragma unroll
                        for (int e = 0; e < kAttD; ++e) kvv[e] = __float_as_uint(fmaf(a, vec[e], __uint_as_float(kvv[e])));
                    }
                    uint32_t w[12];
#pragma unroll
                    for (int e = 0; e < 12; ++e)
                        w[e] = pack_bf16(__uint_as_float(kvv[2 * e]) * sc, __uint_as_float(kvv[2 * e + 1]) * sc);
                    uint4 *dst = reinterpret_cast<uint4 *>((wg == 0 ? p.dk : p.dv) + (size_t)(t0 + krow) * p.dqkv_stride + h * kAttD);
                    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                    dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    // ---- dQ_i for every query tile (tiles alternate between the warpgroups)
    tc_fence_after();
    for (int i = wg; i < NB; i += 2) {
        const int row = i * kTile + t128;
        if (i * kTile + (warp & 3) * 32 >= Tg) continue;
        uint32_t qv[32];
        tmem_ld32(tdQ + 32 * i + lane_off, qv);
        tmem_ld_wait();
        if (row < Tg) {
            if (fold) {   // + dS[row][sp] k_sp
                const float a = sB_ds[row];
#pragma unroll
                for (int e = 0; e < kAttD; ++e) qv[e] = __float_as_uint(fmaf(a, sSp[1][e], __uint_as_float(qv[e])));
            }
            uint32_t w[12];
#pragma unroll
            for (int e = 0; e < 12; ++e)
                w[e] = pack_bf16(__uint_as_float(qv[2 * e]) * p.scale, __uint_as_float(qv[2 * e + 1]) * p.scale);
            uint4 *dst = reinterpret_cast<uint4 *>(p.dq + (size_t)(t0 + row) * p.dqkv_stride + h * kAttD);
            dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
            dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
            dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
        }
    }
    if (warp0 && p.accumulate == 2) tma_store_wait_all();   // whichever lane of warp 0 was the elected issuer
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
    MOBGT_STAMP(p.timeline, 4);
}


// =====================================================================================================================
// Single-box variant: every graph of the launch has at most 128 + 1 tokens (BASELINE configs[1]: graphs <= 128 nodes; also
// the natural node-count law, median 4 nodes) and dS goes to the layer's bf16 plane (mode 2, the training path).
//
// The general kernel above keeps S and dP in separate TMEM columns (352 columns with the three accumulators), both score
// tiles in registers (246 registers / thread) and double-buffered operand tiles (177 KB): ONE CTA per SM, whose life is a
// serial chain prologue -> tail -> tile -> epilogue with nothing to overlap it (profiles/r04_k3_timeline.txt).  With a single
// (query tile, key block) per CTA none of that buffering buys anything, so this variant is sized for TWO CTAs per SM — one
// CTA's TMA / MMA / TMEM round trips hide behind the other's softmax-gradient math:
//   * S and dP are produced one after the other into the SAME 128 TMEM columns (P is taken to registers in between):
//     128 + 3 x 32 = 224 columns -> a 256-column allocation;
//   * the bias tile is consumed into P before dS exists, so the dS operand image reuses the bias tile's 32 KB;
//     no double buffers: 97 KB of dynamic shared memory;
//   * one score tile (64 fp32 P values) in registers at a time: <= 128 registers / thread.
// Arithmetic, operand layouts, the dropout mask and the single-token tail are those of the general kernel.
template <bool kDrop>
__global__ void __launch_bounds__(256, 2)
k3_attn_bwd1_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ CUtensorMap tmBias, const __grid_constant__ CUtensorMap tmDS,
                    const AttnBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int kRows = kTile + 4;
    __shared__ __align__(8) uint64_t bar_ld, bar_bias, bar_s, bar_mma;
    __shared__ uint32_t tmem_slot;
    __shared__ float sLse[kRows], sDelta[kRows];
    __shared__ float sA_ds[kRows], sA_pd[kRows];   // tail: dS / dropped P of (row sp, key c)
    __shared__ float sB_ds[kRows], sB_pd[kRows];   // tail: dS / dropped P of (row r, key sp)
    __shared__ float sSp[4][kAttD];                // q, k, v, dO of token sp
    __shared__ float sRed[3][10][kAttD];

    const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, t128 = tid & 127;
    const bool warp0 = warp_index_uniform() == 0;
    const int gi = blockIdx.x / p.H, h = blockIdx.x - gi * p.H;
    const int g = p.order ? p.order[gi] : gi;       // largest graphs first: the CTAs that return at once are dispatched behind them
    const int t0 = p.tok_off[g];
    const int Tg = p.tok_off[g + 1] - t0;            // <= 129
    if (Tg <= p.small_t) return;                     // the whole CTA: nothing has been set up yet
    const bool fold = Tg == kTile + 1;
    const int nv = fold ? kTile : Tg;                // query rows == key columns covered by the MMA tile
    const int sp = Tg - 1;                           // the tail token (fold only)
    const int HD = p.H * kAttD;
    const int plane = g * p.H + h;

    // ---- global loads issued at the very top (consumed after the barrier / TMA set-up): one row per thread (Tg <= 129 < 256)
    float lse_pre = 0.f;
    uint4 o_pre[3], d_pre[3];
    if (tid < Tg) {
        lse_pre = p.lse[(size_t)(t0 + tid) * p.H + h];
        const uint4 *po = reinterpret_cast<const uint4 *>(p.o + (size_t)(t0 + tid) * HD + h * kAttD);
        const uint4 *pd = reinterpret_cast<const uint4 *>(p.dout + (size_t)(t0 + tid) * HD + h * kAttD);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            o_pre[q] = po[q];
            d_pre[q] = pd[q];
        }
    }
    float b_row = 0.f, b_col = 0.f;
    const __nv_bfloat16 *bias_pl = p.bias + (size_t)plane * p.T * p.Tp;
    if (fold) {
        if (tid < Tg) b_row = bf16_at(bias_pl + (size_t)sp * p.Tp + tid);
        if (tid < sp) b_col = bf16_at(bias_pl + (size_t)tid * p.Tp + sp);
    }
    float sp_val = 0.f;
    if (fold && tid < 4 * kAttD) {
        const int which = tid / kAttD, e = tid - which * kAttD;
        const __nv_bfloat16 *src = which == 0 ? p.q : which == 1 ? p.k : which == 2 ? p.v : p.dout;
        const int64_t stride = which == 3 ? (int64_t)HD : p.qkv_stride;
        sp_val = bf16_at(src + (size_t)(t0 + sp) * stride + h * kAttD + e);
    }
    uint8_t *sBD = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);   // bias tile (128-byte swizzle), later the dS operand image
    uint8_t *sP = sBD + kBiasTileBytes;
    uint8_t *sQ = sP + kPBytes;
    uint8_t *sdO = sQ + kBoxBytes;
    uint8_t *sK = sdO + kBoxBytes;
    uint8_t *sV = sK + kBoxBytes;

    if (tid == 0) {
        mbar_init(&bar_ld, 1);
        mbar_init(&bar_bias, 1);
        mbar_init(&bar_s, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmdO);
        tma_prefetch_desc(&tmBias);
        tma_prefetch_desc(&tmDS);
        mbar_expect_tx(&bar_ld, 4 * kBoxTxBytes);
        tma_load_3d(sK, &tmK, &bar_ld, 0, t0, h * kAttChunks);
        tma_load_3d(sV, &tmV, &bar_ld, 0, t0, h * kAttChunks);
        tma_load_3d(sQ, &tmQ, &bar_ld, 0, t0, h * kAttChunks);
        tma_load_3d(sdO, &tmdO, &bar_ld, 0, t0, h * kAttChunks);
        mbar_expect_tx(&bar_bias, kBiasTileBytes);
        tma_load_3d(sBD, &tmBias, &bar_bias, 0, 0, plane);
        tma_load_3d(sBD + kTile * 128, &tmBias, &bar_bias, 64, 0, plane);
    }
    {   // zero the K-padding chunk (d = 24..31) of the four operand boxes
        const uint4 z = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4 *>((wg == 0 ? sQ : sdO) + 3 * kTile * 16 + t128 * 16) = z;
        *reinterpret_cast<uint4 *>((wg == 0 ? sK : sV) + 3 * kTile * 16 + t128 * 16) = z;
    }
    if (tid < kRows) {   // lse (log2 units; +inf past the graph so that p = 0 there) and D = rowsum(dO * O)
        float lse2 = INFINITY, dl = 0.f;
        if (tid < Tg) {
            lse2 = lse_pre * 1.4426950408889634f;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const uint32_t aw[4] = {o_pre[q].x, o_pre[q].y, o_pre[q].z, o_pre[q].w};
                const uint32_t bw[4] = {d_pre[q].x, d_pre[q].y, d_pre[q].z, d_pre[q].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    dl += __uint_as_float(aw[e] << 16) * __uint_as_float(bw[e] << 16);
                    dl += __uint_as_float(aw[e] & 0xFFFF0000u) * __uint_as_float(bw[e] & 0xFFFF0000u);
                }
            }
        }
        sLse[tid] = lse2;
        sDelta[tid] = dl;
    }
    if (fold && tid < 4 * kAttD) sSp[tid / kAttD][tid % kAttD] = sp_val;
    if (warp == 0) tmem_alloc<256>(&tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tS = tmem, tdK = tmem + 128, tdV = tmem + 160, tdQ = tmem + 192;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int nb = round_up(nv, 16);                 // MMA N / K extent along the keys (and along the queries)
    const uint32_t id_sn = make_idesc_bf16(kTile, nb, 0, 0);

    if (warp0 && elect_one()) {   // S = Q K^T
        mbar_wait(&bar_ld, 0);
        tc_fence_after();
        const uint32_t aq = smem_u32(sQ), bk = smem_u32(sK);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            umma_bf16(tS, make_smem_desc(aq + ks * 2 * kTile * 16, kTile * 16, 128),
                      make_smem_desc(bk + ks * 2 * kTile * 16, kTile * 16, 128), id_sn, ks > 0);
        umma_commit(&bar_s);
    }
    __syncwarp();
    const float sl2 = p.scale * 1.4426950408889634f;
    constexpr float kL2e = 1.4426950408889634f;
    uint32_t seed_lo = 0, seed_hi = 0;
    if (kDrop) attn_drop_fold_seed(p.drop, seed_lo, seed_hi);
    const uint32_t th_hi = p.drop.th16 << 16;
    const float ik = p.drop.inv_keep;

    if (fold) {   // ---- the single-token tail, SIMT (overlaps the S MMA)
        __nv_bfloat16 *ds16 = reinterpret_cast<__nv_bfloat16 *>(p.dbias) + (size_t)plane * p.T * p.Tp;
        auto box_row = [&](const uint8_t *box, int row, float (&f)[24]) {
            const uint8_t *b0 = box + row * 16;
            unpack24(*reinterpret_cast<const uint4 *>(b0), *reinterpret_cast<const uint4 *>(b0 + kTile * 16),
                     *reinterpret_cast<const uint4 *>(b0 + 2 * kTile * 16), f);
        };
        mbar_wait(&bar_ld, 0);
        if (tid < Tg) {   // part A: query row sp against key c = tid (itself included)
            const int c = tid;
            const uint32_t rk = kDrop ? attn_drop_rowkey((uint32_t)plane, (uint32_t)sp, seed_lo, seed_hi) : 0u;
            float dot = 0.f, dp = 0.f;
            if (c < sp) {
                float kf[24], vf[24];
                box_row(sK, c, kf);
                box_row(sV, c, vf);
#pragma unroll
                for (int e = 0; e < kAttD; ++e) {
                    dot = fmaf(sSp[0][e], kf[e], dot);
                    dp = fmaf(sSp[3][e], vf[e], dp);
                }
            } else {
#pragma unroll
                for (int e = 0; e < kAttD; ++e) {
                    dot = fmaf(sSp[0][e], sSp[1][e], dot);
                    dp = fmaf(sSp[3][e], sSp[2][e], dp);
                }
            }
            const float pr = bwd_exp2(fmaf(dot, sl2, b_row * kL2e) - sLse[sp]);
            bool kp = true;
            if (kDrop) kp = (attn_drop_keep8(rk, (uint32_t)(c >> 3), p.drop.th16) >> (c & 7)) & 1u;
            const float ds = pr * ((kDrop ? (kp ? dp * ik : 0.f) : dp) - sDelta[sp]);
            sA_ds[c] = ds;
            sA_pd[c] = kDrop ? (kp ? pr * ik : 0.f) : pr;
            ds16[(size_t)sp * p.Tp + c] = __float2bfloat16_rn(ds);
        }
        if (tid < sp) {   // part B: query row r = tid against key sp
            const int r = tid;
            float qf[24], df[24];
            box_row(sQ, r, qf);
            box_row(sdO, r, df);
            float dot = 0.f, dp = 0.f;
#pragma unroll
            for (int e = 0; e < kAttD; ++e) {
                dot = fmaf(qf[e], sSp[1][e], dot);
                dp = fmaf(df[e], sSp[2][e], dp);
            }
            const float pr = bwd_exp2(fmaf(dot, sl2, b_col * kL2e) - sLse[r]);
            bool kp = true;
            if (kDrop) {
                const uint32_t rk = attn_drop_rowkey((uint32_t)plane, (uint32_t)r, seed_lo, seed_hi);
                kp = (attn_drop_keep8(rk, (uint32_t)(sp >> 3), p.drop.th16) >> (sp & 7)) & 1u;
            }
            const float ds = pr * ((kDrop ? (kp ? dp * ik : 0.f) : dp) - sDelta[r]);
            sB_ds[r] = ds;
            sB_pd[r] = kDrop ? (kp ? pr * ik : 0.f) : pr;
            ds16[(size_t)r * p.Tp + sp] = __float2bfloat16_rn(ds);
        }
        __syncthreads();
        if (tid < 240) {   // dQ[sp] = sum_c dS[sp][c] k_c ; dK[sp] = sum_r dS[r][sp] q_r ; dV[sp] = sum_r P[r][sp] dO_r
            const int seg = tid / kAttD, e = tid - seg * kAttD;
            const int eo = (e >> 3) * (kTile * 16) + (e & 7) * 2;
            float aq = 0.f, ak = 0.f, av = 0.f;
            for (int c = seg; c < sp; c += 10) {
                aq = fmaf(sA_ds[c], bf16_at(reinterpret_cast<const __nv_bfloat16 *>(sK + c * 16 + eo)), aq);
                ak = fmaf(sB_ds[c], bf16_at(reinterpret_cast<const __nv_bfloat16 *>(sQ + c * 16 + eo)), ak);
                av = fmaf(sB_pd[c], bf16_at(reinterpret_cast<const __nv_bfloat16 *>(sdO + c * 16 + eo)), av);
            }
            sRed[0][seg][e] = aq;
            sRed[1][seg][e] = ak;
            sRed[2][seg][e] = av;
        }
        __syncthreads();
        if (tid < 3 * kAttD) {
            const int which = tid / kAttD, e = tid - which * kAttD;
            float a = 0.f;
#pragma unroll
            for (int sg = 0; sg < 10; ++sg) a += sRed[which][sg][e];
            a += which == 0 ? sA_ds[sp] * sSp[1][e] : which == 1 ? sA_ds[sp] * sSp[0][e] : sA_pd[sp] * sSp[3][e];
            if (which < 2) a *= p.scale;
            __nv_bfloat16 *dst = (which == 0 ? p.dq : which == 1 ? p.dk : p.dv) + (size_t)(t0 + sp) * p.dqkv_stride + h * kAttD + e;
            *dst = __float2bfloat16_rn(a);
        }
    }

    // ---- the tile: thread (wg, t128) owns query row t128 and the 16-column chunks c0 = 16 wg + 32 cc
    const int row = t128;
    const bool row_ok = row < nv;
    const bool warp_live = (warp & 3) * 32 < nv;
    const float lse2 = sLse[row];
    const float delta = sDelta[row];
    uint32_t pv[4][16];        // raw S, then (in place) the fp32 probabilities P: the only per-tile register array
    uint32_t keep[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};   // dropout keep bits, bit 16 (cc & 1) + 8 q8 + e of word cc >> 1
    mbar_wait(&bar_s, 0);
    mbar_wait(&bar_bias, 0);
    tc_fence_after();
    if (warp_live) {
        const uint32_t rowkey = kDrop ? attn_drop_rowkey((uint32_t)plane, (uint32_t)row, seed_lo, seed_hi) : 0u;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc)
            if (wg * 16 + cc * 32 < nb) tmem_ld16(tS + lane_off + wg * 16 + cc * 32, pv[cc]);
        tmem_ld_wait();
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = wg * 16 + cc * 32;
            if (c0 >= nb) continue;
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {
                const int c8 = (c0 >> 3) + q8;
                const int colb = c8 * 8;
                const uint8_t *bp = sBD + (c8 >> 3) * (kTile * 128) + t128 * 128 + (((c8 & 7) ^ (t128 & 7)) << 4);
                uint4 bvv = *reinterpret_cast<const uint4 *>(bp);
                if (!row_ok) bvv = make_uint4(0, 0, 0, 0);       // bias rows past the graph are never written: not numbers
                const uint32_t bw[4] = {bvv.x, bvv.y, bvv.z, bvv.w};
                const bool full = colb + 8 <= nv;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float bias = __uint_as_float((e & 1) ? (bw[e >> 1] & 0xFFFF0000u) : (bw[e >> 1] << 16));
                    const float s = fmaf(__uint_as_float(pv[cc][q8 * 8 + e]), sl2, bias * kL2e);
                    float pr = bwd_exp2(s - lse2);               // rows past the graph: lse = +inf -> 0
                    if (!full && colb + e >= nv) pr = 0.f;       // key columns past the graph
                    pv[cc][q8 * 8 + e] = __float_as_uint(pr);
                }
                if (kDrop) {
                    const AttnDropWords dw = attn_drop_words(rowkey, (uint32_t)c8);
                    uint32_t m = 0;
#pragma unroll
                    for (int e = 0; e < 8; ++e) m |= attn_drop_keep(dw, e, th_hi) ? (1u << e) : 0u;
                    const int sh = 16 * (cc & 1) + 8 * q8;
                    keep[cc >> 1] = (keep[cc >> 1] & ~(0xFFu << sh)) | (m << sh);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();              // every thread has S (as P) in registers and is done with the bias tile
    if (warp0 && elect_one()) {   // dP = dO V^T into the columns S occupied
        tc_fence_after();
        const uint32_t ado = smem_u32(sdO), bv = smem_u32(sV);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
            umma_bf16(tS, make_smem_desc(ado + ks * 2 * kTile * 16, kTile * 16, 128),
                      make_smem_desc(bv + ks * 2 * kTile * 16, kTile * 16, 128), id_sn, ks > 0);
        umma_commit(&bar_s);
    }
    __syncwarp();
    if (warp_live) {              // the (dropped-out) P operand of dV, written while the dP MMA runs
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
            const int c0 = wg * 16 + cc * 32;
            if (c0 >= nb) continue;
#pragma unroll
            for (int q8 = 0; q8 < 2; ++q8) {
                const int c8 = (c0 >> 3) + q8;
                float f[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float pr = __uint_as_float(pv[cc][q8 * 8 + e]);
                    if (kDrop) f[e] = ((keep[cc >> 1] >> (16 * (cc & 1) + 8 * q8 + e)) & 1u) ? pr * ik : 0.f;
                    else f[e] = pr;
                }
                uint4 pk;
                pk.x = pack_bf16(f[0], f[1]); pk.y = pack_bf16(f[2], f[3]);
                pk.z = pack_bf16(f[4], f[5]); pk.w = pack_bf16(f[6], f[7]);
                *reinterpret_cast<uint4 *>(sP + c8 * (kTile * 16) + t128 * 16) = pk;
            }
        }
    }
    mbar_wait(&bar_s, 1);
    tc_fence_after();
    if (warp_live) {              // dS = P o (dP - D) -> bf16 operand image over the bias tile
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t dpv[2][16];
#pragma unroll
            for (int u = 0; u < 2; ++u)
                if (wg * 16 + (2 * half + u) * 32 < nb) tmem_ld16(tS + lane_off + wg * 16 + (2 * half + u) * 32, dpv[u]);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int cc = 2 * half + u;
                const int c0 = wg * 16 + cc * 32;
                if (c0 >= nb) continue;
#pragma unroll
                for (int q8 = 0; q8 < 2; ++q8) {
                    const int c8 = (c0 >> 3) + q8;
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float pr = __uint_as_float(pv[cc][q8 * 8 + e]);
                        float dp = __uint_as_float(dpv[u][q8 * 8 + e]);
                        if (kDrop) dp = ((keep[cc >> 1] >> (16 * (cc & 1) + 8 * q8 + e)) & 1u) ? dp * ik : 0.f;
                        f[e] = pr * (dp - delta);
                    }
                    uint4 dk;
                    dk.x = pack_bf16(f[0], f[1]); dk.y = pack_bf16(f[2], f[3]);
                    dk.z = pack_bf16(f[4], f[5]); dk.w = pack_bf16(f[6], f[7]);
                    *reinterpret_cast<uint4 *>(sBD + c8 * (kTile * 16) + t128 * 16) = dk;
                }
            }
        }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp0 && elect_one()) {
        tc_fence_after();
        tma_store_4d(&tmDS, sBD, 0, 0, 0, plane);      // this layer's dS plane straight from the MMA operand image
        tma_store_commit();
        const uint32_t aP = smem_u32(sP), aS = smem_u32(sBD);
        const uint32_t bQ = smem_u32(sQ), bdO = smem_u32(sdO), bK = smem_u32(sK);
        const uint32_t id_t = make_idesc_bf16(kTile, 32, 1, 1);   // A = P^T / dS^T (MN-major), B MN-major
        const uint32_t id_q = make_idesc_bf16(kTile, 32, 0, 1);   // A = dS (K-major),        B MN-major
        for (int ks = 0; ks < nb / 16; ++ks)      // dV = P^T dO   (K = query rows)
            umma_bf16(tdV, make_smem_desc(aP + ks * 256, 128, kTile * 16), make_smem_desc(bdO + ks * 256, 128, kTile * 16), id_t, ks > 0);
        for (int ks = 0; ks < nb / 16; ++ks)      // dK = dS^T Q
            umma_bf16(tdK, make_smem_desc(aS + ks * 256, 128, kTile * 16), make_smem_desc(bQ + ks * 256, 128, kTile * 16), id_t, ks > 0);
        for (int ks = 0; ks < nb / 16; ++ks)      // dQ = dS K     (K = key rows)
            umma_bf16(tdQ, make_smem_desc(aS + ks * 2 * kTile * 16, kTile * 16, 128), make_smem_desc(bK + ks * 256, 128, kTile * 16), id_q, ks > 0);
        umma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    if (warp_live) {
        {   // dK (warpgroup 0) / dV (warpgroup 1): key row = t128
            uint32_t kvv[32];
            tmem_ld32((wg == 0 ? tdK : tdV) + lane_off, kvv);
            tmem_ld_wait();
            if (row_ok) {
                const float sc = wg == 0 ? p.scale : 1.0f;
                if (fold) {   // + dS[sp][krow] q_sp (dK) / P[sp][krow] dO_sp (dV)
                    const float a = wg == 0 ? sA_ds[row] : sA_pd[row];
                    const float *vec = wg == 0 ? sSp[0] : sSp[3];
#pragma unroll
                    for (int e = 0; e < kAttD; ++e) kvv[e] = __float_as_uint(fmaf(a, vec[e], __uint_as_float(kvv[e])));
                }
                uint32_t w[12];
#pragma unroll
                for (int e = 0; e < 12; ++e)
                    w[e] = pack_bf16(__uint_as_float(kvv[2 * e]) * sc, __uint_as_float(kvv[2 * e + 1]) * sc);
                uint4 *dst = reinterpret_cast<uint4 *>((wg == 0 ? p.dk : p.dv) + (size_t)(t0 + row) * p.dqkv_stride + h * kAttD);
                dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                dst[2] = make_uint4(w[8], w[9], w[10], w[11]);
            }
        }
        {   // dQ: columns [0,16) by warpgroup 0, [16,24) by warpgroup 1
            uint32_t qv[16];
            tmem_ld16(tdQ + lane_off + wg * 16, qv);
            tmem_ld_wait();
            if (row_ok) {
                if (fold) {   // + dS[row][sp] k_sp
                    const float a = sB_ds[row];
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (wg == 0 || e < 8) qv[e] = __float_as_uint(fmaf(a, sSp[1][wg * 16 + e], __uint_as_float(qv[e])));
                }
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    w[e] = pack_bf16(__uint_as_float(qv[2 * e]) * p.scale, __uint_as_float(qv[2 * e + 1]) * p.scale);
                uint4 *dst = reinterpret_cast<uint4 *>(p.dq + (size_t)(t0 + row) * p.dqkv_stride + h * kAttD);
                if (wg == 0) {
                    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                } else {
                    dst[2] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    }
    if (warp0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace mobgt

using namespace mobgt;

// Inputs as mobgt_attn_fwd plus o, dout (bf16 [ntok, H*24], contiguous) and lse.  Outputs dq, dk, dv: bf16 with a common
// row stride (usually slices of one [ntok, 3*H*24] buffer, so the fused projection's backward gets a single tensor);
// dbias f32 [B,H,T,Tp]: overwritten (accumulate = 0) or accumulated into (accumulate = 1).
extern "C" int32_t mobgt_attn_bwd(const void *q, const void *k, const void *v, int64_t qkv_row_stride, const void *bias,
                                  const void *o, const void *dout, const float *lse, const int32_t *tok_off,
                                  const int32_t *graph_order, int32_t B, int32_t H, int32_t ntok, int32_t T, int32_t Tp, int32_t t_max_host, int32_t t_min_host,
                                  float scale, void *dq,
                                  void *dk, void *dv, int64_t dqkv_row_stride, void *dbias, int32_t accumulate, float drop_p,
                                  uint64_t seed, const void *seed_dev, void *stream) {
    MOBGT_REQUIRE(q && k && v && bias && o && dout && lse && tok_off && dq && dk && dv && dbias, MOBGT_ERR_NULL,
                  "mobgt_attn_bwd: null pointer");
    MOBGT_REQUIRE(qkv_row_stride % 8 == 0 && dqkv_row_stride % 8 == 0 && Tp % 8 == 0 && Tp >= T, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_attn_bwd: strides / Tp must be multiples of 8");
    MOBGT_REQUIRE(t_max_host >= 1 && t_max_host <= T && T <= MOBGT_MAX_NODES + 1, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_attn_bwd: t_max=%d T=%d", t_max_host, T);
    if (B == 0 || ntok == 0) return MOBGT_OK;
    MOBGT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, MOBGT_ERR_BAD_SHAPE, "mobgt_attn_bwd: drop_p=%f must be in [0, 1)", drop_p);
    // the same device-side split by graph size as the forward (k3_attn_small.cu)
    const bool any_small = (t_min_host <= 0 || t_min_host <= kSmallT) && H % 4 == 0;
    const bool any_large = t_max_host > kSmallT;
    ForkJoin fj{};
    bool forked = false;
    if (any_small) {
        SmallAttnParams sp{};
        sp.tok_off = tok_off;
        sp.order = graph_order;
        sp.q = static_cast<const __nv_bfloat16 *>(q); sp.k = static_cast<const __nv_bfloat16 *>(k); sp.v = static_cast<const __nv_bfloat16 *>(v);
        sp.qkv_stride = qkv_row_stride;
        sp.bias = static_cast<const __nv_bfloat16 *>(bias);
        sp.H = H; sp.T = T; sp.Tp = Tp; sp.scale = scale; sp.small_t = kSmallT;
        sp.drop = make_attn_drop(drop_p, seed, seed_dev);
        sp.lse = const_cast<float *>(lse);
        sp.o = static_cast<const __nv_bfloat16 *>(o); sp.dout = static_cast<const __nv_bfloat16 *>(dout);
        sp.dq = static_cast<__nv_bfloat16 *>(dq); sp.dk = static_cast<__nv_bfloat16 *>(dk); sp.dv = static_cast<__nv_bfloat16 *>(dv);
        sp.dqkv_stride = dqkv_row_stride;
        sp.dbias = dbias; sp.accumulate = accumulate;
        if (!any_large) return launch_small_attn_bwd(sp, B, static_cast<cudaStream_t>(stream));
        // both kernels: they work on disjoint graphs, so the SIMT one runs on a side stream next to the tensor-core one
        int32_t rc = get_fork_join(1, &fj);
        if (rc) return rc;
        MOBGT_CUDA_OK(cudaEventRecord(fj.fork, static_cast<cudaStream_t>(stream)));
        MOBGT_CUDA_OK(cudaStreamWaitEvent(fj.side, fj.fork, 0));
        rc = launch_small_attn_bwd(sp, B, fj.side);
        MOBGT_CUDA_OK(cudaEventRecord(fj.join, fj.side));
        forked = true;
        if (rc) {
            cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), fj.join, 0);
            return rc;
        }
    }
    auto run_tensor_core = [&]() -> int32_t {
    CUtensorMap tmQ, tmK, tmV, tmdO, tmB;
    const void *ptrs[4] = {q, k, v, dout};
    CUtensorMap *maps[4] = {&tmQ, &tmK, &tmV, &tmdO};
    for (int i = 0; i < 4; ++i) {
        uint64_t dims[3] = {8, (uint64_t)ntok, (uint64_t)H * kAttChunks};
        uint64_t str[2] = {(uint64_t)(i < 3 ? qkv_row_stride : H * kAttD) * 2, 16};
        uint32_t box[3] = {8, kTile, kAttChunks};
        int32_t rc = encode_tmap_bf16(maps[i], ptrs[i], 3, dims, str, box, 0);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)Tp, (uint64_t)T, (uint64_t)B * H};
        uint64_t str[2] = {(uint64_t)Tp * 2, (uint64_t)T * Tp * 2};
        uint32_t box[3] = {64, kTile, 1};
        int32_t rc = encode_tmap_bf16(&tmB, bias, 3, dims, str, box, 1);
        if (rc) return rc;
    }
    CUtensorMap tmDS = tmB;   // only used in mode 2
    if (accumulate == 2) {    // dbias is a bf16 [B,H,T,Tp] plane of this layer, written in the [chunk][row][8] box order
        uint64_t dims[4] = {8, (uint64_t)T, (uint64_t)Tp / 8, (uint64_t)B * H};
        uint64_t str[3] = {(uint64_t)Tp * 2, 16, (uint64_t)T * Tp * 2};
        uint32_t box[4] = {8, kTile, kTile / 8, 1};
        int32_t rc = encode_tmap_bf16(&tmDS, dbias, 4, dims, str, box, 0);
        if (rc) return rc;
    }
    // a graph of 128 m + 1 tokens keeps only its m full boxes in shared memory (single-token tail), and T <= 513
    const int max_boxes = t_max_host > kTile && t_max_host % kTile == 1 ? t_max_host / kTile : ceil_div(t_max_host, kTile);
    const size_t smem_base = 2 * kPBytes + 4 * kBoxBytes + (size_t)2 * max_boxes * kBoxBytes + 1024;
    const int bias_bufs = (smem_base + 2 * kBiasTileBytes <= 204 * 1024) ? 2 : 1;    // + 20 KB of static shared memory
    const size_t smem = smem_base + (size_t)bias_bufs * kBiasTileBytes;
    MOBGT_REQUIRE(drop_p >= 0.f && drop_p < 1.f, MOBGT_ERR_BAD_SHAPE, "mobgt_attn_bwd: drop_p=%f must be in [0, 1)", drop_p);
    const AttnDrop drop = make_attn_drop(drop_p, seed, seed_dev);
    auto kern = drop.th16 ? k3_attn_bwd_kernel<true> : k3_attn_bwd_kernel<false>;
    MOBGT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AttnBwdParams p{tok_off,
                    static_cast<const __nv_bfloat16 *>(o),
                    static_cast<const __nv_bfloat16 *>(dout),
                    lse,
                    static_cast<__nv_bfloat16 *>(dq),
                    static_cast<__nv_bfloat16 *>(dk),
                    static_cast<__nv_bfloat16 *>(dv),
                    dqkv_row_stride,
                    static_cast<float *>(dbias),
                    H,
                    T,
                    Tp,
                    scale,
                    max_boxes,
                    accumulate,
                    bias_bufs,
                    g_timeline_dev,
                    drop,
                    static_cast<const __nv_bfloat16 *>(q),
                    static_cast<const __nv_bfloat16 *>(k),
                    static_cast<const __nv_bfloat16 *>(v),
                    qkv_row_stride,
                    static_cast<const __nv_bfloat16 *>(bias),
                    any_small ? kSmallT : 0, graph_order};
    if (accumulate == 2 && t_max_host <= kTile + 1) {   // every graph fits one (query tile, key block): two CTAs per SM
        const size_t smem1 = (size_t)kBiasTileBytes + kPBytes + 4 * kBoxBytes + 1024;
        auto kern1 = drop.th16 ? k3_attn_bwd1_kernel<true> : k3_attn_bwd1_kernel<false>;
        MOBGT_CUDA_OK(cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        MOBGT_CUDA_OK(cudaFuncSetAttribute(kern1, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        kern1<<<B * H, 256, smem1, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, tmdO, tmB, tmDS, p);
        MOBGT_LAUNCH_OK("k3_attn_bwd1_kernel");
        return MOBGT_OK;
    }
    kern<<<B * H, 256, smem, static_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, tmdO, tmB, tmDS, p);
    MOBGT_LAUNCH_OK("k3_attn_bwd_kernel");
    return MOBGT_OK;
    };
    const int32_t rc = run_tensor_core();
    if (forked) cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), fj.join, 0);   // join on every path
    return rc;
}
