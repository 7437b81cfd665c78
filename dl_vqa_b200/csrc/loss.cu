// Fused log-softmax + soft-target VQA loss + gradient + VQA accuracy score in one pass over the logits
// (train.py:190-206 and utils/train_utils.py:12-25).  One CTA per sample; the row of logits is read
// once from HBM (N*4 bytes), dlogits written once; no host round trips (the reference does
// 2 D2H syncs for index building and B .item() syncs for the score).
#include "common.cuh"

namespace {

constexpr int LT = 256;

__global__ void __launch_bounds__(LT)
softloss_rows_kernel(const float* __restrict__ logits, const int64_t* __restrict__ a_idx,
                     const int64_t* __restrict__ a_val, float* __restrict__ dlogits, float* __restrict__ loss_rows,
                     float* __restrict__ score_rows, int B, int N, int A) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    __shared__ int redi[32];
    __shared__ float s_bcast[2];
    __shared__ int s_arg;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* y = logits + (int64_t)b * N;

    // max + first argmax
    float best = -INFINITY; int arg = 0x7fffffff;
    for (int n = tid; n < N; n += LT) { const float v = y[n]; if (v > best) { best = v; arg = n; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
    }
    if (lane == 0) { red[warp] = best; redi[warp] = arg; }
    __syncthreads();
    if (warp == 0) {
        best = lane < LT / 32 ? red[lane] : -INFINITY;
        arg = lane < LT / 32 ? redi[lane] : 0x7fffffff;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ov > best || (ov == best && oa < arg)) { best = ov; arg = oa; }
        }
        if (lane == 0) { s_bcast[0] = best; s_arg = arg; }
    }
    __syncthreads();
    const float mx = s_bcast[0];
    const int amax = s_arg;

    float se = 0.f;
    for (int n = tid; n < N; n += LT) se += __expf(y[n] - mx);
    se = block_sum(se, red);
    const float lse = mx + __logf(se);
    const float inv_se = 1.f / se;

    // sparse soft targets: slots with a_idx != 0 (train.py:197-203); weights a_val/10
    float wsum = 0.f, lrow = 0.f, cnt = 0.f;
    for (int j = 0; j < A; ++j) {
        const int64_t id = a_idx[(int64_t)b * A + j];
        if (id >= 1 && id <= N) {                    // 0 = padding; ids outside 1..N (the reference would raise) are ignored
            const float w = (float)a_val[(int64_t)b * A + j] / 10.0f;
            wsum += w;
            lrow += w * (lse - y[id - 1]);
            if ((int)(id - 1) == amax) cnt = (float)a_val[(int64_t)b * A + j];
        }
    }
    if (tid == 0) {
        loss_rows[b] = lrow;
        score_rows[b] = fminf(0.3f * cnt, 1.0f);
    }
    if (dlogits) {
        const float invB = 1.f / (float)B;
        float* d = dlogits + (int64_t)b * N;
        const float k = wsum * inv_se * invB;
        for (int n = tid; n < N; n += LT) d[n] = __expf(y[n] - mx) * k;
        __syncthreads();
        if (tid == 0)
            for (int j = 0; j < A; ++j) {
                const int64_t id = a_idx[(int64_t)b * A + j];
                if (id >= 1 && id <= N) d[id - 1] -= ((float)a_val[(int64_t)b * A + j] / 10.0f) * invB;
            }
    }
}

__global__ void __launch_bounds__(LT)
softloss_reduce_kernel(const float* __restrict__ loss_rows, const float* __restrict__ score_rows,
                       float* __restrict__ loss_out, float* __restrict__ score_out, int B) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[32];
    float l = 0.f, s = 0.f;
    for (int i = threadIdx.x; i < B; i += LT) { l += loss_rows[i]; s += score_rows[i]; }
    l = block_sum(l, red);
    s = block_sum(s, red);
    if (threadIdx.x == 0) { *loss_out = l / (float)B; *score_out = s; }
}

}  // namespace

extern "C" int vqa_softloss_fwd_bwd(const float* logits, const int64_t* a_idx, const int64_t* a_val, float* dlogits,
                                    float* loss_rows, float* score_rows, float* loss_out, float* score_out,
                                    int B, int N, int A, void* stream) {
    VQA_REQUIRE(B > 0 && N > 0 && A >= 0, "softloss: bad dims B=%d N=%d A=%d", B, N, A);
    VQA_REQUIRE(logits && a_idx && a_val && loss_rows && score_rows && loss_out && score_out, "softloss: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    VQA_CUDA(vqa_launch_pdl(softloss_rows_kernel, dim3(B), dim3(LT), 0, st, logits, a_idx, a_val, dlogits, loss_rows, score_rows, B, N, A));
    VQA_CHECK_LAUNCH("softloss_rows");
    VQA_CUDA(vqa_launch_pdl(softloss_reduce_kernel, dim3(1), dim3(LT), 0, st, loss_rows, score_rows, loss_out, score_out, B));
    VQA_CHECK_LAUNCH("softloss_reduce");
    return 0;
}
