// Bandwidth-bound helpers of the VQA step: dropout + L2 normalisation, embedding + tanh, LSTM backward
// pointwise, column sums (bias gradients), casts, dropout application.
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// image.drop + channel L2 norm (models/model.py:84, :56).  One warp per spatial row of C channels; a lane owns
// groups of 8 consecutive channels (128-bit loads / stores), the row stays in registers between the reduction
// and the scaling pass, and both dropout sites use the 8-element vector flags of common.cuh (one hash per group).
// ------------------------------------------------------------------------------------------
namespace {
constexpr int DN_MAXK = 4;            // up to 4 groups of 8 channels per lane: C <= 1024

}  // namespace

template <typename T, int NK>
__global__ void dropnorm_fwd_kernel(const T* __restrict__ x, T* __restrict__ vn, T* __restrict__ vnd,
                                    float* __restrict__ nrm, int64_t R, int C, Dropout d_img, Dropout d_att) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const Dropout8 di = make_dropout8(d_img, SITE_IMAGE), da = make_dropout8(d_att, SITE_ATT_V);
    const int c8n = C >> 3;
    float v[NK][8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int c8 = lane + 32 * k;
        if (c8 < c8n) {
            float m[8];
            ld8(x + r * C + c8 * 8, v[k]);
            dropout_mult8(di, (uint32_t)(r * c8n + c8), m);
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[k][i] *= m[i]; ss = fmaf(v[k][i], v[k][i], ss); }
        }
    }
    ss = warp_sum(ss);
    const float n = sqrtf(ss);
    const float inv = 1.f / (n + 1e-12f);
    if (lane == 0) nrm[r] = n;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int c8 = lane + 32 * k;
        if (c8 < c8n) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[k][i] *= inv;
            st8(vn + r * C + c8 * 8, v[k]);
            if (vnd) {
                float m[8];
                dropout_mult8(da, (uint32_t)(r * c8n + c8), m);
#pragma unroll
                for (int i = 0; i < 8; ++i) v[k][i] *= m[i];
                st8(vnd + r * C + c8 * 8, v[k]);
            }
        }
    }
}

template <typename T, int NK>
__global__ void dropnorm_bwd_kernel(const T* __restrict__ dvn, const T* __restrict__ dvnd, const T* __restrict__ vn,
                                    const float* __restrict__ nrm, T* __restrict__ dx, int64_t R, int C,
                                    Dropout d_img, Dropout d_att) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const Dropout8 di = make_dropout8(d_img, SITE_IMAGE), da = make_dropout8(d_att, SITE_ATT_V);
    const int c8n = C >> 3;
    float dy[NK][8], y[NK][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int c8 = lane + 32 * k;
        if (c8 < c8n) {
            const int64_t off = r * C + c8 * 8;
            if (dvn) ld8(dvn + off, dy[k]);
            else {
#pragma unroll
                for (int i = 0; i < 8; ++i) dy[k][i] = 0.f;
            }
            if (dvnd) {
                float t[8], m[8];
                ld8(dvnd + off, t);
                dropout_mult8(da, (uint32_t)(r * c8n + c8), m);
#pragma unroll
                for (int i = 0; i < 8; ++i) dy[k][i] = fmaf(t[i], m[i], dy[k][i]);
            }
            ld8(vn + off, y[k]);
#pragma unroll
            for (int i = 0; i < 8; ++i) s = fmaf(dy[k][i], y[k][i], s);
        }
    }
    s = warp_sum(s);
    const float n = nrm[r];
    const float inv = 1.f / (n + 1e-12f);
    const float kk = n > 0.f ? s / n : 0.f;
#pragma unroll
    for (int k = 0; k < NK; ++k) {
        const int c8 = lane + 32 * k;
        if (c8 < c8n) {
            float m[8], g[8];
            dropout_mult8(di, (uint32_t)(r * c8n + c8), m);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = (inv * dy[k][i] - kk * y[k][i]) * m[i];
            st8(dx + r * C + c8 * 8, g);
        }
    }
}

extern "C" int vqa_dropnorm_fwd(const void* x, void* vn, void* vnd, float* nrm, int act_dtype, int64_t R, int C,
                                float p_img, float p_att, uint64_t seed, void* stream) {
    VQA_REQUIRE(R > 0 && C > 0 && x && vn && nrm, "dropnorm_fwd: bad arguments");
    VQA_REQUIRE(C % 8 == 0 && C <= 256 * DN_MAXK, "dropnorm_fwd: channel count %d must be a multiple of 8 and <= %d", C, 256 * DN_MAXK);
    VQA_REQUIRE(R * (C / 8) < (1ll << 32), "dropnorm_fwd: tensor too large for the 32-bit dropout counter");
    const Dropout di = make_dropout(seed, p_img), da = make_dropout(seed, p_att);
    const int wpb = 8;
    const unsigned grid = (unsigned)ceil_div64(R, wpb);
    // NK = groups of 8 channels per lane: 1 covers C <= 256 (the config.yaml shape) with a third of the registers
    if (act_dtype == VQA_F32) {
        if (C <= 256) VQA_CUDA(vqa_launch_pdl(dropnorm_fwd_kernel<float, 1>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, (const float*)x, (float*)vn, (float*)vnd, nrm, R, C, di, da));
        else VQA_CUDA(vqa_launch_pdl(dropnorm_fwd_kernel<float, DN_MAXK>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, (const float*)x, (float*)vn, (float*)vnd, nrm, R, C, di, da));
    } else if (act_dtype == VQA_BF16) {
        if (C <= 256) VQA_CUDA(vqa_launch_pdl(dropnorm_fwd_kernel<bf16, 1>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)vn, (bf16*)vnd, nrm, R, C, di, da));
        else VQA_CUDA(vqa_launch_pdl(dropnorm_fwd_kernel<bf16, DN_MAXK>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)vn, (bf16*)vnd, nrm, R, C, di, da));
    } else VQA_REQUIRE(false, "dropnorm_fwd: bad dtype");
    VQA_CHECK_LAUNCH("dropnorm_fwd");
    return 0;
}

extern "C" int vqa_dropnorm_bwd(const void* dvn, const void* dvnd, const void* vn, const float* nrm, void* dx,
                                int act_dtype, int64_t R, int C, float p_img, float p_att, uint64_t seed, void* stream) {
    VQA_REQUIRE(R > 0 && C > 0 && vn && nrm && dx && (dvn || dvnd), "dropnorm_bwd: bad arguments");
    VQA_REQUIRE(C % 8 == 0 && C <= 256 * DN_MAXK, "dropnorm_bwd: channel count %d must be a multiple of 8 and <= %d", C, 256 * DN_MAXK);
    const Dropout di = make_dropout(seed, p_img), da = make_dropout(seed, p_att);
    const int wpb = 8;
    const unsigned grid = (unsigned)ceil_div64(R, wpb);
    if (act_dtype == VQA_F32) {
        if (C <= 256) VQA_CUDA(vqa_launch_pdl(dropnorm_bwd_kernel<float, 1>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, (const float*)dvn, (const float*)dvnd, (const float*)vn, nrm, (float*)dx, R, C, di, da));
        else VQA_CUDA(vqa_launch_pdl(dropnorm_bwd_kernel<float, DN_MAXK>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, (const float*)dvn, (const float*)dvnd, (const float*)vn, nrm, (float*)dx, R, C, di, da));
    } else if (act_dtype == VQA_BF16) {
        if (C <= 256) VQA_CUDA(vqa_launch_pdl(dropnorm_bwd_kernel<bf16, 1>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, (const bf16*)dvn, (const bf16*)dvnd, (const bf16*)vn, nrm, (bf16*)dx, R, C, di, da));
        else VQA_CUDA(vqa_launch_pdl(dropnorm_bwd_kernel<bf16, DN_MAXK>, dim3(grid), dim3(wpb * 32), 0, (cudaStream_t)stream, (const bf16*)dvn, (const bf16*)dvnd, (const bf16*)vn, nrm, (bf16*)dx, R, C, di, da));
    } else VQA_REQUIRE(false, "dropnorm_bwd: bad dtype");
    VQA_CHECK_LAUNCH("dropnorm_bwd");
    return 0;
}

// Same backward, fused with the max-pool backward of the last conv layer and its bias gradient (tensor-core arm):
// instead of the compact gradient w.r.t. the pooled activation it writes the UN-POOLED gradient
//   dy[b, 2ph+dy, 2pw+dx, c] = mask[b,ph,pw,c] == dy*2+dx ? g : 0
// that the conv dgrad / wgrad kernels consume, and db[c] = sum of g where the ReLU was alive (mask < 4).
// One pass instead of dropnorm_bwd + unpool (saves writing and re-reading the compact gradient and re-reading the mask).
__global__ void __launch_bounds__(256, 3)
dropnorm_bwd_unpool_kernel(const bf16* __restrict__ dvn, const bf16* __restrict__ dvnd, const bf16* __restrict__ vn,
                           const float* __restrict__ nrm, const uint8_t* __restrict__ mask, bf16* __restrict__ dy,
                           float* __restrict__ db, int64_t R, int C, int PH, int PW, Dropout d_img, Dropout d_att) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[8][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Dropout8 di = make_dropout8(d_img, SITE_IMAGE), da = make_dropout8(d_att, SITE_ATT_V);
    const int c8n = C >> 3;
    const bool act = lane < c8n;
    float bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
    // software pipeline: the five loads of the NEXT row are issued before the current row's reduction / stores
    struct Row { uint4 g, gd, y; uint2 mk; float n; };
    auto fetch = [&](int64_t r, Row& x) {
        x.g = x.gd = x.y = make_uint4(0u, 0u, 0u, 0u);
        x.mk = make_uint2(0x04040404u, 0x04040404u);
        x.n = 0.f;
        if (act && r < R) {
            const int64_t off = r * C + lane * 8;
            if (dvn) x.g = __ldcs(reinterpret_cast<const uint4*>(dvn + off));
            if (dvnd) x.gd = __ldcs(reinterpret_cast<const uint4*>(dvnd + off));
            x.y = __ldcs(reinterpret_cast<const uint4*>(vn + off));
            x.mk = __ldcs(reinterpret_cast<const uint2*>(mask + off));
            x.n = nrm[r];
        }
    };
    const int64_t stride = (int64_t)gridDim.x * 8;
    Row nxt;
    fetch((int64_t)blockIdx.x * 8 + warp, nxt);
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < R; r += stride) {
        const Row cur = nxt;
        fetch(r + stride, nxt);
        float dyv[8], y[8], t[8];
        unpack8(cur.g, dyv);
        unpack8(cur.gd, t);
        unpack8(cur.y, y);
        if (dvnd) {
            float m[8];
            dropout_mult8(da, (uint32_t)(r * c8n + lane), m);
#pragma unroll
            for (int i = 0; i < 8; ++i) dyv[i] = fmaf(t[i], m[i], dyv[i]);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(dyv[i], y[i], s);
        s = warp_sum(s);
        if (act) {
            const float n = cur.n;
            const float inv = 1.f / (n + 1e-12f);
            const float kk = n > 0.f ? s / n : 0.f;
            float m[8];
            dropout_mult8(di, (uint32_t)(r * c8n + lane), m);
            uint32_t gq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn((inv * dyv[2 * i] - kk * y[2 * i]) * m[2 * i],
                                                                (inv * dyv[2 * i + 1] - kk * y[2 * i + 1]) * m[2 * i + 1]);
                gq[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            {   // the bias sees what the conv kernels see (bf16), where the ReLU was alive (mask byte != 4)
                const uint32_t z0 = cur.mk.x << 5, z1 = cur.mk.y << 5;      // bit 2 -> sign bit of each byte
                const uint32_t a0 = gq[0] & ~sign_mask16_lo(z0), a1 = gq[1] & ~sign_mask16_hi(z0);
                const uint32_t a2 = gq[2] & ~sign_mask16_lo(z1), a3 = gq[3] & ~sign_mask16_hi(z1);
                bsum[0] += __uint_as_float(a0 << 16); bsum[1] += __uint_as_float(a0 & 0xffff0000u);
                bsum[2] += __uint_as_float(a1 << 16); bsum[3] += __uint_as_float(a1 & 0xffff0000u);
                bsum[4] += __uint_as_float(a2 << 16); bsum[5] += __uint_as_float(a2 & 0xffff0000u);
                bsum[6] += __uint_as_float(a3 << 16); bsum[7] += __uint_as_float(a3 & 0xffff0000u);
            }
            const int pw = (int)(r % PW);
            const int64_t t2 = r / PW;
            const int ph = (int)(t2 % PH);
            const int64_t b = t2 / PH;
            bf16* obase = dy + ((b * 2 * PH + 2 * ph) * (2 * PW) + 2 * pw) * C + lane * 8;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // bytes equal to e -> 0x08 -> sign bit; prmt replicates the sign over each bf16
                const uint32_t z0 = (0x08080808u - (cur.mk.x ^ (0x01010101u * e))) << 4;
                const uint32_t z1 = (0x08080808u - (cur.mk.y ^ (0x01010101u * e))) << 4;
                uint4 o;
                o.x = gq[0] & sign_mask16_lo(z0); o.y = gq[1] & sign_mask16_hi(z0);
                o.z = gq[2] & sign_mask16_lo(z1); o.w = gq[3] & sign_mask16_hi(z1);
                __stcs(reinterpret_cast<uint4*>(obase + ((int64_t)(e >> 1) * (2 * PW) + (e & 1)) * C), o);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = bsum[i];
    __syncthreads();
    if (threadIdx.x < C) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(db + threadIdx.x, t);
    }
}

extern "C" int vqa_dropnorm_bwd_unpool(const void* dvn, const void* dvnd, const void* vn, const float* nrm, const uint8_t* mask,
                                       void* dy, float* db, int B, int PH, int PW, int C, float p_img, float p_att,
                                       uint64_t seed, void* stream) {
    VQA_REQUIRE(B > 0 && PH > 0 && PW > 0 && vn && nrm && mask && dy && db && (dvn || dvnd), "dropnorm_bwd_unpool: bad arguments");
    VQA_REQUIRE(C % 8 == 0 && C <= 256, "dropnorm_bwd_unpool: channel count %d must be a multiple of 8 and <= 256", C);
    const int64_t R = (int64_t)B * PH * PW;
    VQA_REQUIRE(R * (C / 8) < (1ll << 32), "dropnorm_bwd_unpool: tensor too large for the 32-bit dropout counter");
    const Dropout di = make_dropout(seed, p_img), da = make_dropout(seed, p_att);
    cudaStream_t st = (cudaStream_t)stream;
    VQA_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, st));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = ceil_div64(R, 8);
    const unsigned grid = (unsigned)(want < (int64_t)sms * 3 ? want : (int64_t)sms * 3);     // one resident wave
    VQA_CUDA(vqa_launch_pdl(dropnorm_bwd_unpool_kernel, dim3(grid), dim3(256), 0, st, (const bf16*)dvn, (const bf16*)dvnd, (const bf16*)vn, nrm, mask, (bf16*)dy, db,
                                                     R, C, PH, PW, di, da));
    VQA_CHECK_LAUNCH("dropnorm_bwd_unpool");
    return 0;
}

// ------------------------------------------------------------------------------------------
// embedding -> dropout -> tanh, written step-indexed for both LSTM directions (models/model.py:155-162)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int step_token_pos(int dir, int s, int len) { return dir == 0 ? s : len - 1 - s; }

// `order` (optional, vqa_length_order): row b of the step-indexed output holds sample order[b]; the dropout mask stays a
// function of the SAMPLE index, so ordering the batch does not change which elements are dropped
template <typename T>
__global__ void embed_tanh_fwd_kernel(const int64_t* __restrict__ q, const int64_t* __restrict__ q_len,
                                      const int* __restrict__ order,
                                      const float* __restrict__ emb, T* __restrict__ xs,
                                      int B, int T_, int E, int ldx, int dirs, Dropout d) {
    pdl_trigger();
    pdl_wait();
    const int64_t row = blockIdx.x;                 // (dir, s, b)
    const int b = order ? order[(int)(row % B)] : (int)(row % B);
    const int s = (int)((row / B) % T_);
    const int dir = (int)(row / ((int64_t)B * T_));
    const int len = clamp_len(q_len[b], T_);
    T* o = xs + row * ldx;
    const bool active = s < len;
    int t = 0; int64_t tok = 0;
    if (active) { t = step_token_pos(dir, s, len); tok = q[(int64_t)b * T_ + t]; }
    for (int e = threadIdx.x; e < ldx; e += blockDim.x) {
        float v = 0.f;
        if (active && e < E) {
            const float m = dropout_mult(d, SITE_EMBED, ((uint64_t)b * T_ + t) * E + e);
            v = tanhf(emb[tok * E + e] * m);
        }
        o[e] = from_f32<T>(v);
    }
}

template <typename T>
__global__ void embed_tanh_bwd_kernel(const int64_t* __restrict__ q, const int64_t* __restrict__ q_len,
                                      const int* __restrict__ order,
                                      const T* __restrict__ xs, const T* __restrict__ dxs, float* __restrict__ demb,
                                      int B, int T_, int E, int ldx, int dirs, Dropout d) {
    pdl_trigger();
    pdl_wait();
    const int64_t row = blockIdx.x;
    const int b = order ? order[(int)(row % B)] : (int)(row % B);
    const int s = (int)((row / B) % T_);
    const int dir = (int)(row / ((int64_t)B * T_));
    const int len = clamp_len(q_len[b], T_);
    if (s >= len) return;
    const int t = step_token_pos(dir, s, len);
    const int64_t tok = q[(int64_t)b * T_ + t];
    if (tok == 0) return;                            // padding_idx: no gradient
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const float y = to_f32(xs[row * ldx + e]);
        const float m = dropout_mult(d, SITE_EMBED, ((uint64_t)b * T_ + t) * E + e);
        const float g = to_f32(dxs[row * ldx + e]) * (1.f - y * y) * m;
        atomicAdd(demb + tok * E + e, g);
    }
}

extern "C" int vqa_embed_tanh_fwd_ordered(const int64_t* q, const int64_t* q_len, const int* order, const float* emb, void* xs,
                                          int act_dtype, int B, int T, int E, int ldx, int dirs, float p, uint64_t seed, void* stream) {
    VQA_REQUIRE(B > 0 && T > 0 && E > 0 && ldx >= E && (dirs == 1 || dirs == 2), "embed_fwd: bad dims");
    const Dropout d = make_dropout(seed, p);
    const unsigned grid = (unsigned)((int64_t)dirs * T * B);
    if (act_dtype == VQA_F32)
        VQA_CUDA(vqa_launch_pdl(embed_tanh_fwd_kernel<float>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, q, q_len, order, emb, (float*)xs, B, T, E, ldx, dirs, d));
    else if (act_dtype == VQA_BF16)
        VQA_CUDA(vqa_launch_pdl(embed_tanh_fwd_kernel<bf16>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, q, q_len, order, emb, (bf16*)xs, B, T, E, ldx, dirs, d));
    else VQA_REQUIRE(false, "embed_fwd: bad dtype");
    VQA_CHECK_LAUNCH("embed_tanh_fwd");
    return 0;
}

extern "C" int vqa_embed_tanh_fwd(const int64_t* q, const int64_t* q_len, const float* emb, void* xs, int act_dtype,
                                  int B, int T, int E, int ldx, int dirs, float p, uint64_t seed, void* stream) {
    return vqa_embed_tanh_fwd_ordered(q, q_len, nullptr, emb, xs, act_dtype, B, T, E, ldx, dirs, p, seed, stream);
}

extern "C" int vqa_embed_tanh_bwd_ordered(const int64_t* q, const int64_t* q_len, const int* order, const void* xs, const void* dxs,
                                          float* demb, int act_dtype, int B, int T, int E, int ldx, int dirs, float p, uint64_t seed,
                                          void* stream) {
    VQA_REQUIRE(B > 0 && T > 0 && E > 0 && ldx >= E && (dirs == 1 || dirs == 2), "embed_bwd: bad dims");
    const Dropout d = make_dropout(seed, p);
    const unsigned grid = (unsigned)((int64_t)dirs * T * B);
    if (act_dtype == VQA_F32)
        VQA_CUDA(vqa_launch_pdl(embed_tanh_bwd_kernel<float>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, q, q_len, order, (const float*)xs, (const float*)dxs, demb, B, T, E, ldx, dirs, d));
    else if (act_dtype == VQA_BF16)
        VQA_CUDA(vqa_launch_pdl(embed_tanh_bwd_kernel<bf16>, dim3(grid), dim3(128), 0, (cudaStream_t)stream, q, q_len, order, (const bf16*)xs, (const bf16*)dxs, demb, B, T, E, ldx, dirs, d));
    else VQA_REQUIRE(false, "embed_bwd: bad dtype");
    VQA_CHECK_LAUNCH("embed_tanh_bwd");
    return 0;
}

extern "C" int vqa_embed_tanh_bwd(const int64_t* q, const int64_t* q_len, const void* xs, const void* dxs, float* demb,
                                  int act_dtype, int B, int T, int E, int ldx, int dirs, float p, uint64_t seed,
                                  void* stream) {
    return vqa_embed_tanh_bwd_ordered(q, q_len, nullptr, xs, dxs, demb, act_dtype, B, T, E, ldx, dirs, p, seed, stream);
}

// ------------------------------------------------------------------------------------------
// Length order of a batch of questions (what pack_padded_sequence(enforce_sorted=False) computes on the host in the
// reference, models/model.py:160): order[j] = sample at position j when the batch is sorted by DESCENDING length, ties
// in sample order (stable, deterministic); len_sorted[j] = clamp(q_len[order[j]], 0, T).  With the batch in this order
// the rows that are still active at step s are a prefix, so the persistent LSTM kernels drop whole 128-row tiles from
// the late steps (mean question length is about half of T).  One CTA; rank by counting (B <= 8192).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) length_order_kernel(const int64_t* __restrict__ q_len, int* __restrict__ order,
                                                            int64_t* __restrict__ len_sorted, int B, int T_) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ int lens[];
    for (int b = threadIdx.x; b < B; b += blockDim.x) lens[b] = clamp_len(q_len[b], T_);
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int mine = lens[b];
        int rank = 0;
        for (int o = 0; o < B; ++o) {
            const int l = lens[o];                 // broadcast read
            rank += (l > mine) || (l == mine && o < b);
        }
        order[rank] = b;
        len_sorted[rank] = mine;
    }
}

extern "C" int vqa_length_order(const int64_t* q_len, int* order, int64_t* len_sorted, int B, int T, void* stream) {
    VQA_REQUIRE(q_len && order && len_sorted && B > 0 && T > 0, "length_order: bad arguments");
    VQA_REQUIRE(B <= 8192, "length_order: batch %d too large (<= 8192)", B);
    const int threads = B < 1024 ? ((B + 31) / 32) * 32 : 1024;
    VQA_CUDA(vqa_launch_pdl(length_order_kernel, dim3(1), dim3(threads), (size_t)B * sizeof(int), (cudaStream_t)stream, q_len, order, len_sorted, B, T));
    VQA_CHECK_LAUNCH("length_order");
    return 0;
}

// ------------------------------------------------------------------------------------------
// Reduction-block lists for the LSTM weight gradients (vqa_tc_gemm_kblocks).  The step-indexed buffers hold rows
// (s, r), s = 0..T-1, r = 0..B-1; the gate gradient dg[s][r] is exactly zero when s >= len[r] (question r has ended:
// pack_padded_sequence drops those positions, models/model.py:160-164), so a 64-row block all of whose rows have ended
// adds nothing to dW_ih = sum dg^T x or dW_hh = sum dg^T h_prev.  CTA c (c = 0, 1) lists, in ascending order, the blocks
// of steps c..T-1 (block index relative to step c) that still hold a live row.  Works for any row order; with the rows
// in descending length order (vqa_length_order) the live rows of a step are a prefix and about half of the blocks drop
// out at uniformly distributed lengths.
// ------------------------------------------------------------------------------------------
__global__ void lstm_active_kblocks_kernel(const int64_t* __restrict__ len, int* __restrict__ list0, int* __restrict__ list1,
                                           int B, int T_) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ int gmax[];                       // [B / 64] longest question of every 64-row group
    const int s_begin = blockIdx.x;
    int* list = s_begin == 0 ? list0 : list1;
    if (list == nullptr) return;
    const int G = B >> 6;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int g = warp; g < G; g += nwarps) {
        const int64_t a = len[g * 64 + lane], b = len[g * 64 + 32 + lane];
        int m = (int)(a > b ? a : b);
        m = m < 0 ? 0 : (m > T_ ? T_ : m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) gmax[g] = m;
    }
    __syncthreads();
    if (warp != 0) return;
    const int nblk = (T_ - s_begin) * G;
    int n = 0;
    for (int base = 0; base < nblk; base += 32) {        // ordered compaction, 32 blocks per ballot
        const int b = base + lane;
        const bool live = b < nblk && gmax[b % G] > s_begin + b / G;
        const unsigned int bal = __ballot_sync(0xffffffffu, live);
        if (live) list[1 + n + __popc(bal & ((1u << lane) - 1u))] = b;
        n += __popc(bal);
    }
    if (lane == 0) list[0] = n;
}

extern "C" int vqa_lstm_active_kblocks(const int64_t* len_rows, int32_t* list0, int32_t* list1, int B, int T, void* stream) {
    VQA_REQUIRE(len_rows && (list0 || list1) && B > 0 && T > 0, "lstm_active_kblocks: bad arguments");
    VQA_REQUIRE(B % 64 == 0 && B <= 8192, "lstm_active_kblocks: B = %d must be a multiple of 64 (<= 8192)", B);
    VQA_CUDA(vqa_launch_pdl(lstm_active_kblocks_kernel, dim3(2), dim3(256), (size_t)(B / 64) * sizeof(int), (cudaStream_t)stream,
                            len_rows, (int*)list0, (int*)list1, B, T));
    VQA_CHECK_LAUNCH("lstm_active_kblocks");
    return 0;
}

// ------------------------------------------------------------------------------------------
// LSTM backward, pointwise part of step s (BPTT through the cell of models/model.py:164)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void lstm_bwd_pointwise_kernel(const T* __restrict__ gates, const float* __restrict__ cs,
                                          float* __restrict__ dh, float* __restrict__ dc,
                                          const T* __restrict__ dc_init, T* __restrict__ dg,
                                          const int64_t* __restrict__ q_len, int s, int T_, int B, int H, int dirs) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over dirs*B*H
    if (i >= (int64_t)dirs * B * H) return;
    const int j = (int)(i % H);
    const int b = (int)((i / H) % B);
    const int dir = (int)(i / ((int64_t)B * H));
    const int64_t row = ((int64_t)dir * T_ + s) * B + b;
    T* o = dg + row * 4 * H;
    // gradient w.r.t. the final cell state enters at the first processed step (s == T-1)
    const float dc_in = dc_init ? to_f32(dc_init[(int64_t)b * dirs * H + (int64_t)dir * H + j]) : dc[i];
    if (s >= clamp_len(q_len[b], T_)) {
        o[j] = o[H + j] = o[2 * H + j] = o[3 * H + j] = from_f32<T>(0.f);
        dc[i] = dc_in;                              // frozen step: dc passes through unchanged
        return;
    }
    const T* g = gates + row * 4 * H;
    const float gi = to_f32(g[j]), gf = to_f32(g[H + j]), gg = to_f32(g[2 * H + j]), go = to_f32(g[3 * H + j]);
    const float c = cs[row * H + j];
    const float c_prev = s > 0 ? cs[(row - B) * H + j] : 0.f;
    const float tc = tanhf(c);
    const float dhv = dh[i];
    dh[i] = 0.f;                                    // read-and-clear: the next step's data gradient ACCUMULATES into dh (split-K)
    float dcv = dc_in + dhv * go * (1.f - tc * tc);
    const float d_o = dhv * tc;
    const float d_i = dcv * gg, d_g = dcv * gi, d_f = dcv * c_prev;
    dc[i] = dcv * gf;
    o[j] = from_f32<T>(d_i * gi * (1.f - gi));
    o[H + j] = from_f32<T>(d_f * gf * (1.f - gf));
    o[2 * H + j] = from_f32<T>(d_g * (1.f - gg * gg));
    o[3 * H + j] = from_f32<T>(d_o * go * (1.f - go));
}

// bf16 tensor-core arm: 8 hidden units per thread, 128-bit accesses (the scalar kernel above is latency-bound: 2-byte
// accesses to four gate planes)
__global__ void __launch_bounds__(128)
lstm_bwd_pointwise_vec8_kernel(const bf16* __restrict__ gates, const float* __restrict__ cs, float* __restrict__ dh,
                               float* __restrict__ dc, const bf16* __restrict__ dc_init, bf16* __restrict__ dg,
                               const int64_t* __restrict__ q_len, int s, int T_, int B, int H, int dirs) {
    pdl_trigger();
    pdl_wait();
    const int h8 = H >> 3;
    const int64_t i8 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over dirs*B*H/8
    if (i8 >= (int64_t)dirs * B * h8) return;
    lstm_bwd_pointwise_item8<false>(i8, gates, cs, dh, dc, dc_init, dg, q_len, s, T_, B, H, dirs);
}

extern "C" int vqa_lstm_step_bwd_pointwise(const void* gates, const float* cs, float* dh, float* dc,
                                           const void* dc_init, void* dg, const int64_t* q_len, int act_dtype, int s, int T, int B, int H, int dirs,
                                           void* stream) {
    VQA_REQUIRE(s >= 0 && s < T && B > 0 && H > 0 && (dirs == 1 || dirs == 2), "lstm bwd pointwise: bad dims");
    const int64_t n = (int64_t)dirs * B * H;
    const unsigned grid = (unsigned)ceil_div64(n, 256);
    if (act_dtype == VQA_F32)
        VQA_CUDA(vqa_launch_pdl(lstm_bwd_pointwise_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)gates, cs, dh, dc, (const float*)dc_init, (float*)dg, q_len, s, T, B, H, dirs));
    else if (act_dtype == VQA_BF16 && H % 8 == 0)
        VQA_CUDA(vqa_launch_pdl(lstm_bwd_pointwise_vec8_kernel, dim3((unsigned)ceil_div64(n / 8, 128)), dim3(128), 0, (cudaStream_t)stream,
                                (const bf16*)gates, cs, dh, dc, (const bf16*)dc_init, (bf16*)dg, q_len, s, T, B, H, dirs));
    else if (act_dtype == VQA_BF16)
        VQA_CUDA(vqa_launch_pdl(lstm_bwd_pointwise_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)gates, cs, dh, dc, (const bf16*)dc_init, (bf16*)dg, q_len, s, T, B, H, dirs));
    else VQA_REQUIRE(false, "lstm bwd pointwise: bad dtype");
    VQA_CHECK_LAUNCH("lstm_step_bwd_pointwise");
    return 0;
}

// ------------------------------------------------------------------------------------------
// dropout application / gradient merges / casts / column sums
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void dropout_apply_kernel(const T* __restrict__ in, int64_t ld_in, T* __restrict__ out, int64_t ld_out,
                                     int64_t rows, int cols, Dropout d, uint32_t site) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols; const int c = (int)(i - r * cols);
    out[r * ld_out + c] = from_f32<T>(to_f32(in[r * ld_in + c]) * dropout_mult(d, site, (uint64_t)i));
}

extern "C" int vqa_dropout_apply(const void* in, int64_t ld_in, void* out, int64_t ld_out, int dtype, int64_t rows,
                                 int cols, float p, uint64_t seed, uint32_t site, void* stream) {
    VQA_REQUIRE(rows >= 0 && cols > 0 && ld_in >= cols && ld_out >= cols, "dropout_apply: bad dims");
    if (rows == 0) return 0;
    const Dropout d = make_dropout(seed, p);
    const unsigned grid = (unsigned)ceil_div64(rows * cols, 256);
    if (dtype == VQA_F32)
        VQA_CUDA(vqa_launch_pdl(dropout_apply_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)in, ld_in, (float*)out, ld_out, rows, cols, d, site));
    else if (dtype == VQA_BF16)
        VQA_CUDA(vqa_launch_pdl(dropout_apply_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)in, ld_in, (bf16*)out, ld_out, rows, cols, d, site));
    else VQA_REQUIRE(false, "dropout_apply: bad dtype");
    VQA_CHECK_LAUNCH("dropout_apply");
    return 0;
}

// keep[i] = 1 when element i of dropout site `site` is kept.  vector8 = 0: the per-element scheme (dropout_mult: embedding,
// classifier, q' sites); 1: the 8-element vector scheme (dropout_mult8: image, v and attention x sites).
__global__ void dropout_mask_kernel(uint8_t* __restrict__ keep, int64_t n, Dropout d, uint32_t site, int vector8) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m;
    if (vector8) {
        float m8[8];
        dropout_mult8(make_dropout8(d, site), (uint32_t)(i >> 3), m8);
        m = m8[i & 7];
    } else {
        m = dropout_mult(d, site, (uint64_t)i);
    }
    keep[i] = m != 0.f;
}

extern "C" int vqa_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint32_t site, int vector8, void* stream) {
    VQA_REQUIRE(keep && n >= 0 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
    if (n == 0) return 0;
    const Dropout d = make_dropout(seed, p);
    VQA_CUDA(vqa_launch_pdl(dropout_mask_kernel, dim3((unsigned)ceil_div64(n, 256)), dim3(256), 0, (cudaStream_t)stream, keep, n, d, site, vector8));
    VQA_CHECK_LAUNCH("dropout_mask");
    return 0;
}

template <typename T>
__global__ void add_dropped_kernel(const T* __restrict__ a, int64_t lda, const T* __restrict__ b, int64_t ldb,
                                   T* __restrict__ dst, int64_t ldd, int64_t rows, int cols, Dropout d, uint32_t site) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols; const int c = (int)(i - r * cols);
    float v = to_f32(a[r * lda + c]);
    if (b) v += to_f32(b[r * ldb + c]) * dropout_mult(d, site, (uint64_t)i);
    dst[r * ldd + c] = from_f32<T>(v);
}

extern "C" int vqa_add_dropped(const void* a, int64_t lda, const void* b, int64_t ldb, void* dst, int64_t ldd, int dtype,
                               int64_t rows, int cols, float p, uint64_t seed, uint32_t site, void* stream) {
    VQA_REQUIRE(rows >= 0 && cols > 0 && a && dst, "add_dropped: bad arguments");
    if (rows == 0) return 0;
    const Dropout d = make_dropout(seed, p);
    const unsigned grid = (unsigned)ceil_div64(rows * cols, 256);
    if (dtype == VQA_F32)
        VQA_CUDA(vqa_launch_pdl(add_dropped_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)a, lda, (const float*)b, ldb, (float*)dst, ldd, rows, cols, d, site));
    else if (dtype == VQA_BF16)
        VQA_CUDA(vqa_launch_pdl(add_dropped_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)a, lda, (const bf16*)b, ldb, (bf16*)dst, ldd, rows, cols, d, site));
    else VQA_REQUIRE(false, "add_dropped: bad dtype");
    VQA_CHECK_LAUNCH("add_dropped");
    return 0;
}

template <typename T>
__global__ void relu_drop_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dz, int64_t n,
                                     float scale) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dz[i] = from_f32<T>(to_f32(y[i]) > 0.f ? to_f32(dy[i]) * scale : 0.f);
}

extern "C" int vqa_relu_drop_bwd(const void* dy, const void* y, void* dz, int dtype, int64_t n, float p, void* stream) {
    VQA_REQUIRE(n >= 0 && p >= 0.f && p < 1.f, "relu_drop_bwd: bad arguments");
    if (n == 0) return 0;
    const float scale = 1.f / (1.f - p);
    const unsigned grid = (unsigned)ceil_div64(n, 256);
    if (dtype == VQA_F32)
        VQA_CUDA(vqa_launch_pdl(relu_drop_bwd_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const float*)dy, (const float*)y, (float*)dz, n, scale));
    else if (dtype == VQA_BF16)
        VQA_CUDA(vqa_launch_pdl(relu_drop_bwd_kernel<bf16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, (const bf16*)dy, (const bf16*)y, (bf16*)dz, n, scale));
    else VQA_REQUIRE(false, "relu_drop_bwd: bad dtype");
    VQA_CHECK_LAUNCH("relu_drop_bwd");
    return 0;
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = from_f32<TD>(to_f32(s[i]));
}

extern "C" int vqa_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream) {
    VQA_REQUIRE(n >= 0, "cast: bad size");
    if (n == 0) return 0;
    const unsigned grid = (unsigned)ceil_div64(n, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == VQA_F32 && dst_dtype == VQA_BF16) VQA_CUDA(vqa_launch_pdl(cast_kernel<float, bf16>, dim3(grid), dim3(256), 0, st, (const float*)src, (bf16*)dst, n));
    else if (src_dtype == VQA_BF16 && dst_dtype == VQA_F32) VQA_CUDA(vqa_launch_pdl(cast_kernel<bf16, float>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (float*)dst, n));
    else if (src_dtype == VQA_F32 && dst_dtype == VQA_F32) VQA_CUDA(vqa_launch_pdl(cast_kernel<float, float>, dim3(grid), dim3(256), 0, st, (const float*)src, (float*)dst, n));
    else if (src_dtype == VQA_BF16 && dst_dtype == VQA_BF16) VQA_CUDA(vqa_launch_pdl(cast_kernel<bf16, bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)src, (bf16*)dst, n));
    else VQA_REQUIRE(false, "cast: bad dtypes");
    VQA_CHECK_LAUNCH("cast");
    return 0;
}

// 2-D cast with row pitches; destination columns cols..dst_cols-1 are zero filled (TMA-friendly padding)
template <typename TS, typename TD>
__global__ void cast2d_kernel(const TS* __restrict__ s, int64_t lds, TD* __restrict__ d, int64_t ldd, int64_t rows,
                              int cols, int dst_cols) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * dst_cols) return;
    const int64_t r = i / dst_cols; const int c = (int)(i - r * dst_cols);
    d[r * ldd + c] = from_f32<TD>(c < cols ? to_f32(s[r * lds + c]) : 0.f);
}

extern "C" int vqa_cast_2d(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd,
                           int64_t rows, int cols, int dst_cols, void* stream) {
    VQA_REQUIRE(rows >= 0 && cols > 0 && dst_cols >= cols && lds >= cols && ldd >= dst_cols, "cast_2d: bad dims");
    if (rows == 0) return 0;
    const unsigned grid = (unsigned)ceil_div64(rows * dst_cols, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == VQA_F32 && dst_dtype == VQA_BF16) VQA_CUDA(vqa_launch_pdl(cast2d_kernel<float, bf16>, dim3(grid), dim3(256), 0, st, (const float*)src, lds, (bf16*)dst, ldd, rows, cols, dst_cols));
    else if (src_dtype == VQA_BF16 && dst_dtype == VQA_F32) VQA_CUDA(vqa_launch_pdl(cast2d_kernel<bf16, float>, dim3(grid), dim3(256), 0, st, (const bf16*)src, lds, (float*)dst, ldd, rows, cols, dst_cols));
    else if (src_dtype == VQA_F32 && dst_dtype == VQA_F32) VQA_CUDA(vqa_launch_pdl(cast2d_kernel<float, float>, dim3(grid), dim3(256), 0, st, (const float*)src, lds, (float*)dst, ldd, rows, cols, dst_cols));
    else if (src_dtype == VQA_BF16 && dst_dtype == VQA_BF16) VQA_CUDA(vqa_launch_pdl(cast2d_kernel<bf16, bf16>, dim3(grid), dim3(256), 0, st, (const bf16*)src, lds, (bf16*)dst, ldd, rows, cols, dst_cols));
    else VQA_REQUIRE(false, "cast_2d: bad dtypes");
    VQA_CHECK_LAUNCH("cast_2d");
    return 0;
}

// column sums: block = 32 columns x 8 row lanes; rows split over gridDim.y, combined with atomics
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ in, int64_t ld, const uint8_t* __restrict__ mask,
                              float* __restrict__ out, int64_t rows, int cols, int64_t rows_per_block) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
    const int64_t r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    if (c < cols)
        for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
            const int64_t idx = r * ld + c;
            if (!mask || mask[idx] < 4) s += to_f32(in[idx]);
        }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

int vqa_colsum_impl(const void* in, int dtype, int64_t ld, const uint8_t* mask, float* out, int64_t rows, int cols,
                    cudaStream_t st) {
    VQA_REQUIRE(rows >= 0 && cols > 0 && ld >= cols, "colsum: bad dims");
    if (rows == 0) return 0;
    const int gx = ceil_div(cols, 32);
    int64_t gy = ceil_div64(148 * 8, gx);
    const int64_t max_gy = ceil_div64(rows, 64);
    if (gy > max_gy) gy = max_gy;
    if (gy < 1) gy = 1;
    const int64_t rpb = ceil_div64(rows, gy);
    dim3 grid(gx, (unsigned)ceil_div64(rows, rpb)), block(32, 8);
    if (dtype == VQA_F32) VQA_CUDA(vqa_launch_pdl(colsum_kernel<float>, dim3(grid), dim3(block), 0, st, (const float*)in, ld, mask, out, rows, cols, rpb));
    else if (dtype == VQA_BF16) VQA_CUDA(vqa_launch_pdl(colsum_kernel<bf16>, dim3(grid), dim3(block), 0, st, (const bf16*)in, ld, mask, out, rows, cols, rpb));
    else VQA_REQUIRE(false, "colsum: bad dtype");
    VQA_CHECK_LAUNCH("colsum");
    return 0;
}

extern "C" int vqa_colsum(const void* in, int dtype, int64_t ld, const uint8_t* mask, float* out, int64_t rows, int cols,
                          void* stream) {
    return vqa_colsum_impl(in, dtype, ld, mask, out, rows, cols, (cudaStream_t)stream);
}
