// Self-test of the tcgen05 / TMA building blocks (test hook; exercised by tests/test_umma_selftest.py).
// D[128,N] = A * B with A, B loaded by 3-D TMA into the no-swizzle [chunk][row][8] layout, in each
// combination of K-major / MN-major operands that the attention kernels rely on.
#include "common.cuh"
#include "umma.cuh"

namespace mobgt {
using namespace sm100;

struct SelfParams {
    int N, K, a_mn, b_mn;
    float *out;  // [128, N]
};

__global__ void __launch_bounds__(128) selftest_umma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB, SelfParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int M = 128, N = p.N, K = p.K;
    uint8_t *sA = smem;                    // M*K*2 bytes
    uint8_t *sB = smem + (size_t)M * K * 2;  // N*K*2 bytes

    if (tid == 0) {
        mbar_init(&bar_load, 1);
        mbar_init(&bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<256>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (tid == 0) {
        mbar_expect_tx(&bar_load, (uint32_t)((M + N) * K * 2));
        // box = {8, rows, chunks}: rows = M (K-major) or K (MN-major)
        if (p.a_mn == 2) {   // SWIZZLE_128B K-major: one 2-D box {64, rows} per 64-wide K block
            for (int kb = 0; kb < K / 64; ++kb) {
                tma_load_2d(sA + (size_t)kb * M * 128, &tmA, &bar_load, kb * 64, 0);
                tma_load_2d(sB + (size_t)kb * N * 128, &tmB, &bar_load, kb * 64, 0);
            }
        } else {
            tma_load_3d(sA, &tmA, &bar_load, 0, 0, 0);
            tma_load_3d(sB, &tmB, &bar_load, 0, 0, 0);
        }
        mbar_wait(&bar_load, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(M, N, p.a_mn == 1, p.b_mn == 1);
        for (int ks = 0; ks < K / 16; ++ks) {
            uint64_t ad, bd;
            if (p.a_mn == 2) {
                ad = make_smem_desc_sw128(smem_u32(sA) + (ks / 4) * M * 128 + (ks % 4) * 32);
                bd = make_smem_desc_sw128(smem_u32(sB) + (ks / 4) * N * 128 + (ks % 4) * 32);
                umma_bf16(tmem, ad, bd, idesc, ks > 0);
                continue;
            }
            if (!p.a_mn) ad = make_smem_desc(smem_u32(sA) + ks * 2 * (M * 16), M * 16, 128);
            else         ad = make_smem_desc(smem_u32(sA) + ks * 256, 128, K * 16);
            if (!p.b_mn) bd = make_smem_desc(smem_u32(sB) + ks * 2 * (N * 16), N * 16, 128);
            else         bd = make_smem_desc(smem_u32(sB) + ks * 256, 128, K * 16);
            umma_bf16(tmem, ad, bd, idesc, ks > 0);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after();
    const int row = tid;
    for (int c = 0; c < N; c += 16) {
        uint32_t r[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
        tmem_ld_wait();
        for (int e = 0; e < 16; ++e) p.out[(size_t)row * N + c + e] = __uint_as_float(r[e]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace mobgt

using namespace mobgt;

// A: a_mn ? [K,128] : [128,K] row-major bf16 ;  B: b_mn ? [K,N] : [N,K] ;  out f32 [128,N]
extern "C" int32_t mobgt_selftest_umma(const void *A, const void *B, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                                       float *out, void *stream) {
    MOBGT_REQUIRE(A && B && out, MOBGT_ERR_NULL, "mobgt_selftest_umma: null pointer");
    MOBGT_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_selftest_umma: N=%d K=%d", N, K);
    const int M = 128;
    CUtensorMap tmA, tmB;
    if (a_mn == 2 || b_mn == 2) {   // both operands K-major, SWIZZLE_128B
        MOBGT_REQUIRE(a_mn == 2 && b_mn == 2 && K % 64 == 0 && N % 8 == 0, MOBGT_ERR_BAD_SHAPE, "mobgt_selftest_umma: sw128 needs K %% 64 == 0");
        uint64_t dimA[2] = {(uint64_t)K, (uint64_t)M}, dimB[2] = {(uint64_t)K, (uint64_t)N};
        uint64_t str[1] = {(uint64_t)K * 2};
        uint32_t boxA[2] = {64, (uint32_t)M}, boxB[2] = {64, (uint32_t)N};
        int32_t rc = encode_tmap_bf16(&tmA, A, 2, dimA, str, boxA, 1);
        if (rc) return rc;
        rc = encode_tmap_bf16(&tmB, B, 2, dimB, str, boxB, 1);
        if (rc) return rc;
    } else {
    {
        const int rows = a_mn ? K : M, cols = a_mn ? M : K;  // row-major [rows, cols]
        uint64_t dims[3] = {8, (uint64_t)rows, (uint64_t)cols / 8};
        uint64_t str[2] = {(uint64_t)cols * 2, 16};
        uint32_t box[3] = {8, (uint32_t)rows, (uint32_t)cols / 8};
        int32_t rc = encode_tmap_bf16(&tmA, A, 3, dims, str, box, 0);
        if (rc) return rc;
    }
    {
        const int rows = b_mn ? K : N, cols = b_mn ? N : K;
        uint64_t dims[3] = {8, (uint64_t)rows, (uint64_t)cols / 8};
        uint64_t str[2] = {(uint64_t)cols * 2, 16};
        uint32_t box[3] = {8, (uint32_t)rows, (uint32_t)cols / 8};
        int32_t rc = encode_tmap_bf16(&tmB, B, 3, dims, str, box, 0);
        if (rc) return rc;
    }
    }
    const size_t smem = (size_t)(M + N) * K * 2 + 1024;
    MOBGT_CUDA_OK(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SelfParams p{N, K, a_mn, b_mn, out};
    selftest_umma_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tmA, tmB, p);
    MOBGT_LAUNCH_OK("selftest_umma_kernel");
    return MOBGT_OK;
}
