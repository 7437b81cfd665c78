// Multi-tensor Adam (train.py:55,76-80: torch.optim.Adam defaults; lr supplied per step by the
// reference's update_learning_rate schedule).  One launch for all parameter tensors; bandwidth-bound:
// reads p,g,m,v and writes p,m,v (+ optional bf16 shadow of p for the tensor-core arm).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
adam_multi_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads,
                  float* const* __restrict__ exp_avg, float* const* __restrict__ exp_avg_sq,
                  void* const* __restrict__ bf16_copy, const int64_t* __restrict__ sizes,
                  float lr_over_bc1, float beta1, float beta2, float eps, float inv_sqrt_bc2, float grad_scale) {
    pdl_trigger();
    pdl_wait();
    const int t = blockIdx.y;
    const int64_t n = sizes[t];
    float* p = params[t];
    const float* g = grads[t];
    float* m = exp_avg[t];
    float* v = exp_avg_sq[t];
    bf16* sh = bf16_copy ? (bf16*)bf16_copy[t] : nullptr;
    auto upd = [&](float gi, float& mi, float& vi, float& pi) {
        gi *= grad_scale;
        mi = beta1 * mi + (1.f - beta1) * gi;
        vi = beta2 * vi + (1.f - beta2) * gi * gi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        pi = pi - lr_over_bc1 * (mi / denom);
    };
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (!sh || (reinterpret_cast<uintptr_t>(sh) & 7) == 0);
    const int64_t n4 = vec ? n >> 2 : 0;
    for (int64_t i = tid; i < n4; i += nthr) {                 // 128-bit accesses: 4 elements per thread and iteration
        const float4 g4 = __ldcs(reinterpret_cast<const float4*>(g) + i);
        float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
        upd(g4.x, m4.x, v4.x, p4.x); upd(g4.y, m4.y, v4.y, p4.y); upd(g4.z, m4.z, v4.z, p4.z); upd(g4.w, m4.w, v4.w, p4.w);
        reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4; reinterpret_cast<float4*>(p)[i] = p4;
        if (sh) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
            uint2 u; u.x = *reinterpret_cast<const uint32_t*>(&lo); u.y = *reinterpret_cast<const uint32_t*>(&hi);
            reinterpret_cast<uint2*>(sh)[i] = u;
        }
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nthr) {         // tail / unaligned tensors
        float mi = m[i], vi = v[i], pi = p[i];
        upd(g[i], mi, vi, pi);
        m[i] = mi; v[i] = vi; p[i] = pi;
        if (sh) sh[i] = __float2bfloat16_rn(pi);
    }
}

}  // namespace

extern "C" int vqa_adam_multi(float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, void* const* bf16_copy, const int64_t* sizes, int n,
                              int64_t max_size, float lr, float beta1, float beta2, float eps, int step,
                              float grad_scale, void* stream) {
    VQA_REQUIRE(n > 0 && step >= 1 && max_size > 0, "adam: bad arguments n=%d step=%d", n, step);
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    int64_t gx = ceil_div64(max_size, 256 * 4);
    if (gx > 148 * 8) gx = 148 * 8;
    dim3 grid((unsigned)gx, (unsigned)n);
    VQA_CUDA(vqa_launch_pdl(adam_multi_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, bf16_copy, sizes,
                                                             (float)(lr / bc1), beta1, beta2, eps,
                                                             (float)(1.0 / sqrt(bc2)), grad_scale));
    VQA_CHECK_LAUNCH("adam_multi");
    return 0;
}
