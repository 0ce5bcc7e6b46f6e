/*
 * libmobgt — C-ABI of the B200-native (sm_100a) MobGT hot path.
 *
 * This is the drop-in boundary.  The reference (Yukayo/MobGT) has no C plugin
 * API: its hot path sits behind Python call sites.  Every entry point below
 * names the reference interface it replaces (file:line under
 * /root/reference/graphormer/).  Host code (mobgt_b200/*.py) binds these with
 * ctypes; see INTEGRATION.md for the stub a reference maintainer would add.
 *
 * Conventions
 *   - every function returns int32 status: 0 = OK, <0 = MOBGT_ERR_*; never throws
 *   - mobgt_last_error() returns a thread-local message for the last failure
 *   - pointers are DEVICE pointers unless the name ends in _host
 *   - no ownership transfer: the caller allocates inputs, outputs and workspaces
 *   - kernels are enqueued on `stream` (a cudaStream_t passed as void*) and the
 *     call never synchronises, so every call is CUDA-graph capturable
 *   - there is no CPU fallback anywhere in this library
 */
#ifndef MOBGT_H_
#define MOBGT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOBGT_OK 0
#define MOBGT_ERR_BAD_SHAPE (-1)
#define MOBGT_ERR_BAD_DTYPE (-2)
#define MOBGT_ERR_UNSUPPORTED (-3)
#define MOBGT_ERR_CUDA (-4)
#define MOBGT_ERR_WORKSPACE_TOO_SMALL (-5)
#define MOBGT_ERR_NULL (-6)

#define MOBGT_UNREACHABLE 510 /* algos.pyx:32 — the reference's "infinity" */
#define MOBGT_MAX_NODES 512   /* rel_pos_encoder has 512 rows (model_fqandtoyo.py:788) */
#define MOBGT_MAX_HOPS 32

/* dtype tags for float tensors crossing the ABI */
#define MOBGT_F32 0
#define MOBGT_BF16 1

int32_t mobgt_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf. */
int32_t mobgt_last_error(char *buf, size_t buflen);
/* 0 if a CUDA device of compute capability 10.x is current, else MOBGT_ERR_CUDA. */
int32_t mobgt_device_check(void);

/* ------------------------------------------------------------------------------------------
 * K1 — preprocessing: batched all-pairs shortest paths + path-edge extraction.
 * Replaces algos.floyd_warshall (algos.pyx:9-54), algos.get_all_edges (:57-62) and
 * algos.gen_edge_input (:65-96) as driven by preprocess_item (wrapper.py:42-61,99) and
 * the hop-axis slice of the collators (collator.py:323).
 *
 * Graphs are packed: graph g has n[g] nodes (1..512) and owns n[g]^2 cells starting at
 * cell offset sq_off[g] of every per-cell array.
 *   feat      u8  [sum n^2]       attn_edge_type (wrapper.py:49-53): 0 = no edge, else
 *                                 edge feature (count+2).  adjacency = (feat != 0).
 *   dist      i16 [sum n^2]       M + shift        (M in 0..509 | 510)          (out)
 *   path      i16 [sum n^2]       FW `path` matrix (0..n-1 | 510), may be NULL  (out)
 *   edge_in   u8  [sum n^2, hops] e + shift; "no hop" (-1 in the reference) is stored
 *                                 as (uint8)(-1 + shift): 255 raw / 0 shifted   (out)
 *   maxdist   i32 [G]             max(M) per graph (wrapper.py:58)              (out)
 * shift = 0 gives the raw algos.pyx values, shift = 1 the collated ones
 * (pad_rel_pos_unsqueeze / pad_3d_unsqueeze "+1", collator.py:76-93).
 * gids (i32 [G_launch], may be NULL = identity) selects which graphs this launch covers;
 * n_max_host is the largest n among them (chooses the cluster size / shared-memory plan).
 * ------------------------------------------------------------------------------------------ */
int32_t mobgt_apsp_edge_input(const uint8_t *feat, const int32_t *n, const int64_t *sq_off,
                              const int32_t *gids, int32_t G_launch, int32_t n_max_host,
                              int32_t hops, int32_t shift,
                              int16_t *dist, int16_t *path, uint8_t *edge_in, int32_t *maxdist,
                              void *stream);

/* Stand-alone mirror of algos.gen_edge_input(max_dist, path, edge_feat) (algos.pyx:65-96) for a
 * caller that already holds a floyd_warshall `path` matrix (i16, packed like above).
 * Writes the first `hops` hop slots; same encoding as edge_in above. */
int32_t mobgt_gen_edge_input(const int16_t *path, const uint8_t *feat, const int32_t *n,
                             const int64_t *sq_off, int32_t G, int32_t n_max_host, int32_t hops,
                             int32_t shift, uint8_t *edge_in, void *stream);

/* in_degree / out_degree exactly as wrapper.py:97-98 names them (row sums / column sums of the
 * bool adjacency), + shift, for packed nodes: node_off[g] = sum_{g'<g} n[g'].  i16 out. */
int32_t mobgt_degrees(const uint8_t *feat, const int32_t *n, const int64_t *sq_off,
                      const int64_t *node_off, int32_t G, int32_t shift,
                      int16_t *in_degree, int16_t *out_degree, void *stream);

/* ------------------------------------------------------------------------------------------
 * Test hooks (tests/test_umma_selftest.py): exercise the tcgen05 / TMA building blocks in isolation.
 * out[128,N] (f32) = A * B, bf16 operands; a_mn / b_mn select MN-major operands
 * (A: a_mn ? [K,128] : [128,K];  B: b_mn ? [K,N] : [N,K], all row-major).
 * ------------------------------------------------------------------------------------------ */
int32_t mobgt_selftest_umma(const void *A, const void *B, int32_t N, int32_t K, int32_t a_mn, int32_t b_mn,
                            float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MOBGT_H_ */
