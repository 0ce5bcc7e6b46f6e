// Interface between the attention entry points (k3_attn_fwd.cu / k3_attn_bwd.cu) and the small-graph SIMT kernels
// (k3_attn_small.cu).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace mobgt {

constexpr int kSmallT = 16;   // graphs of at most this many tokens (n + 1) take the SIMT path (9 trajectory graphs in 10)

struct SmallAttnParams {
    const int32_t *tok_off;
    const int32_t *order;             // [B] graph ids in launch order (descending size) or NULL
    const __nv_bfloat16 *q, *k, *v;   // [ntok, qkv_stride]
    int64_t qkv_stride;
    const __nv_bfloat16 *bias;        // [B,H,T,Tp]
    int H, T, Tp;
    float scale;
    int small_t;                      // graphs with more tokens are left to the tensor-core kernels
    AttnDrop drop;
    // forward
    __nv_bfloat16 *out;               // [ntok, H*24]
    float *lse;                       // [ntok, H]   (backward: input)
    // backward
    const __nv_bfloat16 *o, *dout;    // [ntok, H*24]
    __nv_bfloat16 *dq, *dk, *dv;      // [ntok, dqkv_stride]
    int64_t dqkv_stride;
    void *dbias;                      // f32 (accumulate 0 / 1) or bf16 (accumulate 2) [B,H,T,Tp]
    int accumulate;
};

int32_t launch_small_attn_fwd(const SmallAttnParams &p, int B, cudaStream_t s);      // H % 4 == 0
int32_t launch_small_attn_bwd(const SmallAttnParams &p, int B, cudaStream_t s);

}  // namespace mobgt
