// Multi-tensor Adam (train.py:55,76-80: torch.optim.Adam defaults; lr supplied per step by the
// reference's update_learning_rate schedule).  One launch for all parameter tensors; bandwidth-bound:
// reads p,g,m,v and writes p,m,v (+ optional bf16 shadow of p for the tensor-core arm).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256)
adam_multi_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads,
                  float* const* __restrict__ exp_avg, float* const* __restrict__ exp_avg_sq,
                  void* const* __restrict__ bf16_copy, const int64_t* __restrict__ sizes,
                  float lr_over_bc1, float beta1, float beta2, float eps, float inv_sqrt_bc2, float grad_scale,
                  const VqaStepState* __restrict__ state) {
    pdl_trigger();
    pdl_wait();
    if (state) { lr_over_bc1 = state->lr_over_bc1; inv_sqrt_bc2 = state->inv_sqrt_bc2; }     // published by vqa_step_tick
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
                                                             (float)(1.0 / sqrt(bc2)), grad_scale, (const VqaStepState*)nullptr));
    VQA_CHECK_LAUNCH("adam_multi");
    return 0;
}

extern "C" int vqa_adam_multi_dev(float* const* params, const float* const* grads, float* const* exp_avg,
                                  float* const* exp_avg_sq, void* const* bf16_copy, const int64_t* sizes, int n,
                                  int64_t max_size, const VqaStepState* state, float beta1, float beta2, float eps,
                                  float grad_scale, void* stream) {
    VQA_REQUIRE(n > 0 && max_size > 0 && state, "adam_dev: bad arguments n=%d", n);
    int64_t gx = ceil_div64(max_size, 256 * 4);
    if (gx > 148 * 8) gx = 148 * 8;
    dim3 grid((unsigned)gx, (unsigned)n);
    VQA_CUDA(vqa_launch_pdl(adam_multi_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq, bf16_copy, sizes,
                            0.f, beta1, beta2, eps, 0.f, grad_scale, state));
    VQA_CHECK_LAUNCH("adam_multi_dev");
    return 0;
}

// ---- device-resident step state (include/vqa_b200.h: VqaStepState) ------------------------------------------------
namespace {
__global__ void step_tick_kernel(VqaStepState* st, double lr0, double half_life, double beta1, double beta2) {
    pdl_trigger();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int64_t it = st->iteration;
    const int64_t step = st->adam_step + 1;
    const double lr = lr0 * exp2(-(double)it / half_life);                    // train.py:31-35
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    st->lr = (float)lr;
    st->lr_over_bc1 = (float)(lr / bc1);
    st->inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    st->adam_step = step;
    st->iteration = it + 1;
    uint64_t z = st->seed + 0x9E3779B97F4A7C15ull;                            // splitmix64
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    st->seed = (z ^ (z >> 31)) & ~VQA_SEED_ON_DEVICE;
}

__global__ void scale_by_scalar_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ scalar, int64_t n) {
    pdl_trigger();
    pdl_wait();
    const float k = __ldg(scalar);
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 ? n >> 2 : 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = tid; i < n4; i += nthr) {
        float4 v = reinterpret_cast<const float4*>(src)[i];
        v.x *= k; v.y *= k; v.z *= k; v.w *= k;
        reinterpret_cast<float4*>(dst)[i] = v;
    }
    for (int64_t i = 4 * n4 + tid; i < n; i += nthr) dst[i] = src[i] * k;
}
}  // namespace

extern "C" int vqa_step_tick(VqaStepState* state, double lr0, double half_life, double beta1, double beta2, void* stream) {
    VQA_REQUIRE(state && half_life > 0.0, "step_tick: bad arguments");
    VQA_CUDA(vqa_launch_pdl(step_tick_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, state, lr0, half_life, beta1, beta2));
    VQA_CHECK_LAUNCH("step_tick");
    return 0;
}

extern "C" int vqa_scale_by_device_scalar(const float* src, float* dst, const float* scalar, int64_t n, void* stream) {
    VQA_REQUIRE(src && dst && scalar && n >= 0, "scale_by_device_scalar: bad arguments");
    if (n == 0) return 0;
    int64_t grid = ceil_div64(n, 256 * 4);
    if (grid > 148 * 8) grid = 148 * 8;
    VQA_CUDA(vqa_launch_pdl(scale_by_scalar_kernel, dim3((unsigned)grid), dim3(256), 0, (cudaStream_t)stream, src, dst, scalar, n));
    VQA_CHECK_LAUNCH("scale_by_device_scalar");
    return 0;
}

extern "C" int vqa_zero(void* ptr, int64_t bytes, void* stream) {
    VQA_REQUIRE(bytes >= 0 && (ptr || bytes == 0), "zero: bad arguments");
    if (bytes == 0) return 0;
    VQA_CUDA(cudaMemsetAsync(ptr, 0, (size_t)bytes, (cudaStream_t)stream));
    return 0;
}


extern "C" int vqa_copy(void* dst, const void* src, int64_t bytes, void* stream) {
    VQA_REQUIRE(bytes >= 0 && ((dst && src) || bytes == 0), "copy: bad arguments");
    if (bytes == 0) return 0;
    VQA_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}
