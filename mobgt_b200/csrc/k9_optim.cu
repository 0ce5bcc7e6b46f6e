// K9 — AdamW over the flat parameter / gradient buffers (sm_100a, HBM-bound).
//
// Replaces `torch.optim.AdamW(self.parameters(), lr=peak_lr, weight_decay=weight_decay)` of configure_optimizers
// (model_fqandtoyo.py:1599-1616) for the training loop of the path.  The trainer keeps every parameter and every gradient
// as views of ONE flat fp32 buffer each (the gradient buffer is also what the data-parallel all-reduce exchanges), so the
// optimizer step is a single pass: read p, g, m, v — write p, m, v (28 B / parameter), instead of a multi-tensor launch per
// 30-odd parameter chunks.  Arithmetic = torch's AdamW (decoupled weight decay, bias-corrected moments, no amsgrad):
//     p *= 1 - lr * wd ;  m = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2 ;
//     p -= (lr / (1 - b1^t)) * m / ( sqrt(v) / sqrt(1 - b2^t) + eps )
#include "common.cuh"

namespace mobgt {

__global__ void __launch_bounds__(256) k9_adamw_kernel(float4 *__restrict__ p, const float4 *__restrict__ g, float4 *__restrict__ m,
                                                       float4 *__restrict__ v, int64_t n4, float decay, float b1, float b2,
                                                       float step_size, float inv_sqrt_bc2, float eps) {
    const float ob1 = 1.f - b1, ob2 = 1.f - b2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = p[i], mm = m[i], vv = v[i];
        const float4 gg = __ldg(g + i);
#define MOBGT_ADAMW_LANE(c)                                                  \
        mm.c = fmaf(ob1, gg.c - mm.c, mm.c);                                 \
        vv.c = fmaf(ob2, gg.c * gg.c, b2 * vv.c);                            \
        pp.c = fmaf(-step_size, mm.c / fmaf(sqrtf(vv.c), inv_sqrt_bc2, eps), pp.c * decay);
        MOBGT_ADAMW_LANE(x) MOBGT_ADAMW_LANE(y) MOBGT_ADAMW_LANE(z) MOBGT_ADAMW_LANE(w)
#undef MOBGT_ADAMW_LANE
        p[i] = pp;
        m[i] = mm;
        v[i] = vv;
    }
}

}  // namespace mobgt

using namespace mobgt;

// n must be a multiple of 4 (the caller pads its flat buffers); all four buffers 16-byte aligned.
extern "C" int32_t mobgt_adamw_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                                    float beta1, float beta2, float eps, float weight_decay, int64_t step, void *stream) {
    MOBGT_REQUIRE(param && grad && exp_avg && exp_avg_sq, MOBGT_ERR_NULL, "mobgt_adamw_step: null pointer");
    MOBGT_REQUIRE(n >= 0 && n % 4 == 0 && step >= 1, MOBGT_ERR_BAD_SHAPE, "mobgt_adamw_step: n=%lld (multiple of 4) step=%lld (>= 1)",
                  (long long)n, (long long)step);
    MOBGT_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) == 0, MOBGT_ERR_BAD_SHAPE,
                  "mobgt_adamw_step: buffers must be 16-byte aligned");
    if (n == 0) return MOBGT_OK;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const int64_t n4 = n / 4;
    const int blocks = (int)((n4 + 255) / 256 < (int64_t)kNumSMs * 8 ? (n4 + 255) / 256 : (int64_t)kNumSMs * 8);
    k9_adamw_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<float4 *>(param), reinterpret_cast<const float4 *>(grad), reinterpret_cast<float4 *>(exp_avg),
        reinterpret_cast<float4 *>(exp_avg_sq), n4, 1.0f - lr * weight_decay, beta1, beta2, (float)((double)lr / bc1),
        (float)(1.0 / sqrt(bc2)), eps);
    MOBGT_LAUNCH_OK("k9_adamw_kernel");
    return MOBGT_OK;
}
