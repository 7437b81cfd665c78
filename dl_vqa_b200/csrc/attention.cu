// Fused attention: tile/broadcast + (+|*) + ReLU + dropout + 1x1 glimpse projection + spatial softmax +
// weighted pooling in ONE memory-bound kernel (models/model.py:187-195 and :208-221), and its backward.
//
// Layout: vp [B,P,A] and vn [B,P,C] are channel-contiguous (NHWC), so every global access is a 128-bit
// load/store per lane and a warp covers 512 contiguous bytes.  One CTA per sample; a warp owns a spatial
// position, a lane owns 8 consecutive channels of each 256-channel group.  The [B,A,P] "x" tensor, the
// tiled question and the [B,G,C,P] weighted tensor of the reference never exist.
//
// Algorithmic HBM bytes per sample (forward): (P*A + P*C + G*C) * sizeof(T) + A*4 + G*P*4 (prob, saved
// for backward) -- SURVEY.md section 8d config 3.
#include "common.cuh"

namespace {

constexpr int NW = 8;            // warps per CTA
constexpr int NTHREADS = NW * 32;

__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
    const uint4 u = __ldcs(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(p));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    __stcs(reinterpret_cast<uint4*>(p), u);
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ void load8f(const float* p, float (&v)[8]) {      // cached (re-used) fp32 data
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// dropout multipliers for 8 consecutive elements starting at (even) idx
__device__ __forceinline__ void drop8(const Dropout& d, uint32_t key, uint64_t idx, float (&m)[8]) {
    if (d.threshold == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = 1.f;
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) dropout_mult2(d, key, idx + 2 * i, m[2 * i], m[2 * i + 1]);
}

template <typename T, int G, int OP>
__global__ void __launch_bounds__(NTHREADS)
attention_fwd_kernel(const T* __restrict__ vp, const float* __restrict__ qp, const T* __restrict__ vn,
                     const float* __restrict__ wx, const float* __restrict__ bx, float* __restrict__ prob,
                     T* __restrict__ out, int64_t ldo, int P, int A, int C, Dropout drop) {
    extern __shared__ float sm[];
    float* logit = sm;                 // [G][P]
    float* red = sm + G * P;           // [NW][G*256]
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t key = dropout_key(drop, SITE_ATT_X);

    for (int i = tid; i < G * P; i += NTHREADS) logit[i] = bx[i / P];
    __syncthreads();

    // ---- phase 1: logits[g][s] = sum_a relu(vp[s,a] (+|*) qp[a]) * drop * wx[g][a]
    const int nchunk = A >> 3;
    const T* vpb = vp + (int64_t)b * P * A;
    for (int cg = 0; cg * 32 < nchunk; ++cg) {
        const int ch = cg * 32 + lane;
        const bool act = ch < nchunk;
        const int a0 = ch * 8;
        float qv[8], wv[G][8];
        if (act) {
            load8f(qp + (int64_t)b * A + a0, qv);
#pragma unroll
            for (int g = 0; g < G; ++g) load8f(wx + (int64_t)g * A + a0, wv[g]);
        }
        for (int s0 = warp; s0 < P; s0 += 2 * NW) {
            const int s1 = s0 + NW;
            float x0[8], x1[8], acc0[G], acc1[G];
#pragma unroll
            for (int g = 0; g < G; ++g) { acc0[g] = 0.f; acc1[g] = 0.f; }
            const bool has1 = s1 < P;
            if (act) {
                load8(vpb + (int64_t)s0 * A + a0, x0);
                if (has1) load8(vpb + (int64_t)s1 * A + a0, x1);
                float m[8];
                drop8(drop, key, ((uint64_t)b * P + s0) * A + a0, m);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float pre = OP == VQA_ATT_ADD ? x0[i] + qv[i] : x0[i] * qv[i];
                    const float r = fmaxf(pre, 0.f) * m[i];
#pragma unroll
                    for (int g = 0; g < G; ++g) acc0[g] = fmaf(r, wv[g][i], acc0[g]);
                }
                if (has1) {
                    drop8(drop, key, ((uint64_t)b * P + s1) * A + a0, m);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float pre = OP == VQA_ATT_ADD ? x1[i] + qv[i] : x1[i] * qv[i];
                        const float r = fmaxf(pre, 0.f) * m[i];
#pragma unroll
                        for (int g = 0; g < G; ++g) acc1[g] = fmaf(r, wv[g][i], acc1[g]);
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float t0 = warp_sum(acc0[g]);
                const float t1 = warp_sum(acc1[g]);
                if (lane == 0) {
                    logit[g * P + s0] += t0;
                    if (has1) logit[g * P + s1] += t1;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: spatial softmax per glimpse (warp-shuffle reductions)
    for (int g = warp; g < G; g += NW) {
        float mx = -INFINITY;
        for (int s = lane; s < P; s += 32) mx = fmaxf(mx, logit[g * P + s]);
        mx = warp_max(mx);
        float sum = 0.f;
        for (int s = lane; s < P; s += 32) { const float e = __expf(logit[g * P + s] - mx); logit[g * P + s] = e; sum += e; }
        sum = warp_sum(sum);
        const float inv = 1.f / sum;
        for (int s = lane; s < P; s += 32) {
            const float p = logit[g * P + s] * inv;
            logit[g * P + s] = p;
            if (prob) prob[((int64_t)b * G + g) * P + s] = p;
        }
    }
    __syncthreads();

    // ---- phase 3: out[g][c] = sum_s p[g][s] * vn[s][c]
    const int nchunkc = C >> 3;
    const T* vnb = vn + (int64_t)b * P * C;
    for (int cg = 0; cg * 32 < nchunkc; ++cg) {
        const int ch = cg * 32 + lane;
        const bool act = ch < nchunkc;
        float acc[G][8];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[g][i] = 0.f;
        if (act) {
            for (int s0 = warp; s0 < P; s0 += 2 * NW) {
                const int s1 = s0 + NW;
                float v0[8], v1[8];
                load8(vnb + (int64_t)s0 * C + ch * 8, v0);
                const bool has1 = s1 < P;
                if (has1) load8(vnb + (int64_t)s1 * C + ch * 8, v1);
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float p0 = logit[g * P + s0];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[g][i] = fmaf(p0, v0[i], acc[g][i]);
                    if (has1) {
                        const float p1 = logit[g * P + s1];
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[g][i] = fmaf(p1, v1[i], acc[g][i]);
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) red[(warp * G + g) * 256 + lane * 8 + i] = acc[g][i];
        __syncthreads();
        for (int t = tid; t < G * 256; t += NTHREADS) {
            const int g = t >> 8, c = cg * 256 + (t & 255);
            if (c < C) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += red[(w * G + g) * 256 + (t & 255)];
                out[(int64_t)b * ldo + (int64_t)g * C + c] = from_f32<T>(s);
            }
        }
        __syncthreads();
    }
}

template <typename T, int G, int OP>
__global__ void __launch_bounds__(NTHREADS)
attention_bwd_kernel(const T* __restrict__ dout, int64_t ldd, const T* __restrict__ vp, const float* __restrict__ qp,
                     const T* __restrict__ vn, const float* __restrict__ wx, const float* __restrict__ prob,
                     T* __restrict__ dvp, T* __restrict__ dvn, float* __restrict__ dqp, float* __restrict__ dwx_part,
                     float* __restrict__ dbx_part, int P, int A, int C, Dropout drop) {
    extern __shared__ float sm[];
    float* pr = sm;                       // [G][P] softmax
    float* dl = pr + G * P;               // [G][P] dp, then dlogit
    float* dsm = dl + G * P;              // [G][C] upstream gradient
    float* red = dsm + G * C;             // [NW][(1+G)*256]
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t key = dropout_key(drop, SITE_ATT_X);

    for (int i = tid; i < G * P; i += NTHREADS) { pr[i] = prob[(int64_t)b * G * P + i]; dl[i] = 0.f; }
    for (int i = tid; i < G * C; i += NTHREADS) dsm[i] = to_f32(dout[(int64_t)b * ldd + i]);
    __syncthreads();

    // ---- phase A: dp[g][s] = <dout[g], vn[s]>,  dvn[s] = sum_g p[g][s] dout[g]
    const int nchunkc = C >> 3;
    const T* vnb = vn + (int64_t)b * P * C;
    T* dvnb = dvn + (int64_t)b * P * C;
    for (int cg = 0; cg * 32 < nchunkc; ++cg) {
        const int ch = cg * 32 + lane;
        const bool act = ch < nchunkc;
        float dv[G][8];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) dv[g][i] = act ? dsm[g * C + ch * 8 + i] : 0.f;
        for (int s = warp; s < P; s += NW) {
            float part[G];
#pragma unroll
            for (int g = 0; g < G; ++g) part[g] = 0.f;
            if (act) {
                float v[8], o[8];
                load8(vnb + (int64_t)s * C + ch * 8, v);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
                for (int g = 0; g < G; ++g) {
                    const float pg = pr[g * P + s];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { part[g] = fmaf(dv[g][i], v[i], part[g]); o[i] = fmaf(pg, dv[g][i], o[i]); }
                }
                store8(dvnb + (int64_t)s * C + ch * 8, o);
            }
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float t = warp_sum(part[g]);
                if (lane == 0) dl[g * P + s] += t;
            }
        }
    }
    __syncthreads();

    // ---- phase B: softmax backward, dlogit = p * (dp - <p, dp>)
    for (int g = warp; g < G; g += NW) {
        float dot = 0.f;
        for (int s = lane; s < P; s += 32) dot += pr[g * P + s] * dl[g * P + s];
        dot = warp_sum(dot);
        float sb = 0.f;
        for (int s = lane; s < P; s += 32) {
            const float d = pr[g * P + s] * (dl[g * P + s] - dot);
            dl[g * P + s] = d;
            sb += d;
        }
        sb = warp_sum(sb);
        if (lane == 0) dbx_part[(int64_t)b * G + g] = sb;
    }
    __syncthreads();

    // ---- phase C: through x_conv, dropout, ReLU and the (+|*) fusion
    const int nchunk = A >> 3;
    const T* vpb = vp + (int64_t)b * P * A;
    T* dvpb = dvp + (int64_t)b * P * A;
    for (int cg = 0; cg * 32 < nchunk; ++cg) {
        const int ch = cg * 32 + lane;
        const bool act = ch < nchunk;
        const int a0 = ch * 8;
        float qv[8], wv[G][8], dq[8], dw[G][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { dq[i] = 0.f; qv[i] = 0.f; }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) { dw[g][i] = 0.f; wv[g][i] = 0.f; }
        if (act) {
            load8f(qp + (int64_t)b * A + a0, qv);
#pragma unroll
            for (int g = 0; g < G; ++g) load8f(wx + (int64_t)g * A + a0, wv[g]);
            for (int s = warp; s < P; s += NW) {
                float x[8], m[8], o[8], dls[G];
                load8(vpb + (int64_t)s * A + a0, x);
                drop8(drop, key, ((uint64_t)b * P + s) * A + a0, m);
#pragma unroll
                for (int g = 0; g < G; ++g) dls[g] = dl[g * P + s];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float pre = OP == VQA_ATT_ADD ? x[i] + qv[i] : x[i] * qv[i];
                    const bool alive = pre > 0.f;
                    const float xr = alive ? pre * m[i] : 0.f;
                    float dxt = 0.f;
#pragma unroll
                    for (int g = 0; g < G; ++g) { dxt = fmaf(dls[g], wv[g][i], dxt); dw[g][i] = fmaf(dls[g], xr, dw[g][i]); }
                    const float dpre = alive ? dxt * m[i] : 0.f;
                    if (OP == VQA_ATT_ADD) { o[i] = dpre; dq[i] += dpre; }
                    else { o[i] = dpre * qv[i]; dq[i] = fmaf(dpre, x[i], dq[i]); }
                }
                store8(dvpb + (int64_t)s * A + a0, o);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) red[(warp * (1 + G)) * 256 + lane * 8 + i] = dq[i];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) red[(warp * (1 + G) + 1 + g) * 256 + lane * 8 + i] = dw[g][i];
        __syncthreads();
        for (int t = tid; t < (1 + G) * 256; t += NTHREADS) {
            const int k = t >> 8, a = cg * 256 + (t & 255);
            if (a < A) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += red[(w * (1 + G) + k) * 256 + (t & 255)];
                if (k == 0) dqp[(int64_t)b * A + a] = s;
                else dwx_part[((int64_t)b * G + (k - 1)) * A + a] = s;
            }
        }
        __syncthreads();
    }
}

template <typename T, int G, int OP>
int launch_fwd(const void* vp, const float* qp, const void* vn, const float* wx, const float* bx, float* prob,
               void* out, int64_t ldo, int B, int P, int A, int C, Dropout d, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)G * P + (size_t)NW * G * 256);
    auto kern = attention_fwd_kernel<T, G, OP>;
    if (smem > 48 * 1024) VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, NTHREADS, smem, st>>>((const T*)vp, qp, (const T*)vn, wx, bx, prob, (T*)out, ldo, P, A, C, d);
    VQA_CHECK_LAUNCH("attention_fwd");
    return 0;
}

template <typename T, int G, int OP>
int launch_bwd(const void* dout, int64_t ldd, const void* vp, const float* qp, const void* vn, const float* wx,
               const float* prob, void* dvp, void* dvn, float* dqp, float* dwx_part, float* dbx_part,
               int B, int P, int A, int C, Dropout d, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)2 * G * P + (size_t)G * C + (size_t)NW * (1 + G) * 256);
    auto kern = attention_bwd_kernel<T, G, OP>;
    if (smem > 48 * 1024) VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, NTHREADS, smem, st>>>((const T*)dout, ldd, (const T*)vp, qp, (const T*)vn, wx, prob, (T*)dvp, (T*)dvn,
                                    dqp, dwx_part, dbx_part, P, A, C, d);
    VQA_CHECK_LAUNCH("attention_bwd");
    return 0;
}

}  // namespace

#define ATT_DISPATCH(FN, ...)                                                                        \
    do {                                                                                             \
        const int key__ = (act_dtype == VQA_BF16 ? 100 : 0) + G * 10 + op;                           \
        switch (key__) {                                                                             \
            case 10:  return FN<float, 1, VQA_ATT_ADD>(__VA_ARGS__);                                 \
            case 11:  return FN<float, 1, VQA_ATT_MUL>(__VA_ARGS__);                                 \
            case 20:  return FN<float, 2, VQA_ATT_ADD>(__VA_ARGS__);                                 \
            case 21:  return FN<float, 2, VQA_ATT_MUL>(__VA_ARGS__);                                 \
            case 30:  return FN<float, 3, VQA_ATT_ADD>(__VA_ARGS__);                                 \
            case 31:  return FN<float, 3, VQA_ATT_MUL>(__VA_ARGS__);                                 \
            case 40:  return FN<float, 4, VQA_ATT_ADD>(__VA_ARGS__);                                 \
            case 41:  return FN<float, 4, VQA_ATT_MUL>(__VA_ARGS__);                                 \
            case 110: return FN<bf16, 1, VQA_ATT_ADD>(__VA_ARGS__);                                  \
            case 111: return FN<bf16, 1, VQA_ATT_MUL>(__VA_ARGS__);                                  \
            case 120: return FN<bf16, 2, VQA_ATT_ADD>(__VA_ARGS__);                                  \
            case 121: return FN<bf16, 2, VQA_ATT_MUL>(__VA_ARGS__);                                  \
            case 130: return FN<bf16, 3, VQA_ATT_ADD>(__VA_ARGS__);                                  \
            case 131: return FN<bf16, 3, VQA_ATT_MUL>(__VA_ARGS__);                                  \
            case 140: return FN<bf16, 4, VQA_ATT_ADD>(__VA_ARGS__);                                  \
            case 141: return FN<bf16, 4, VQA_ATT_MUL>(__VA_ARGS__);                                  \
        }                                                                                            \
    } while (0)

static int att_check(int act_dtype, int op, int B, int P, int A, int C, int G, int64_t ld) {
    VQA_REQUIRE(act_dtype == VQA_F32 || act_dtype == VQA_BF16, "attention: bad dtype %d", act_dtype);
    VQA_REQUIRE(op == VQA_ATT_ADD || op == VQA_ATT_MUL, "attention: do_option code %d not supported", op);
    VQA_REQUIRE(B > 0 && P > 0 && A > 0 && C > 0, "attention: bad dims");
    VQA_REQUIRE(G >= 1 && G <= 4, "attention: glimpses=%d not supported (1..4)", G);
    VQA_REQUIRE(A % 8 == 0 && C % 8 == 0, "attention: hidden_dim (%d) and image features (%d) must be multiples of 8", A, C);
    VQA_REQUIRE(ld >= (int64_t)G * C, "attention: row pitch %lld < G*C", (long long)ld);
    VQA_REQUIRE((size_t)(2 * G * P + G * C + NW * (1 + G) * 256) * 4 <= 200 * 1024, "attention: spatial grid too large for shared memory");
    return 0;
}

extern "C" int vqa_attention_fwd(const void* vp, const float* qp, const void* vn, const float* wx, const float* bx,
                                 float* prob, void* out, int64_t ldo, int act_dtype, int op,
                                 int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream) {
    if (int e = att_check(act_dtype, op, B, P, A, C, G, ldo)) return e;
    const Dropout d = make_dropout(seed, p_drop);
    cudaStream_t st = (cudaStream_t)stream;
    ATT_DISPATCH(launch_fwd, vp, qp, vn, wx, bx, prob, out, ldo, B, P, A, C, d, st);
    VQA_REQUIRE(false, "attention_fwd: no kernel for this configuration");
    return 0;
}

extern "C" int vqa_attention_bwd(const void* dout, int64_t ldd, const void* vp, const float* qp, const void* vn,
                                 const float* wx, const float* prob, void* dvp, void* dvn, float* dqp,
                                 float* dwx_part, float* dbx_part, int act_dtype, int op,
                                 int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream) {
    if (int e = att_check(act_dtype, op, B, P, A, C, G, ldd)) return e;
    const Dropout d = make_dropout(seed, p_drop);
    cudaStream_t st = (cudaStream_t)stream;
    ATT_DISPATCH(launch_bwd, dout, ldd, vp, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, A, C, d, st);
    VQA_REQUIRE(false, "attention_bwd: no kernel for this configuration");
    return 0;
}
