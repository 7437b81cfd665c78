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
#include "tc_common.cuh"

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

// dropout multipliers for the 8 consecutive elements starting at idx (a multiple of 8)
__device__ __forceinline__ void drop8(const Dropout8& d, uint64_t idx, float (&m)[8]) { dropout_mult8(d, (uint32_t)(idx >> 3), m); }

template <typename T, int G, int OP>
__global__ void __launch_bounds__(NTHREADS)
attention_fwd_kernel(const T* __restrict__ vp, const float* __restrict__ qp, const T* __restrict__ vn,
                     const float* __restrict__ wx, const float* __restrict__ bx, float* __restrict__ prob,
                     T* __restrict__ out, int64_t ldo, int P, int A, int C, Dropout drop) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    float* logit = sm;                 // [G][P]
    float* red = sm + G * P;           // [NW][G*256]
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Dropout8 d8 = make_dropout8(drop, SITE_ATT_X);

    for (int i = tid; i < G * P; i += NTHREADS) logit[i] = bx[i / P];
    __syncthreads();

    // ---- phase 1: logits[g][s] = sum_a relu(vp[s,a] (+|*) qp[a]) * drop * wx[g][a]
    const int nchunk = A >> 3;
    const T* vpb = vp + (int64_t)b * P * A;
    for (int cg = 0; cg * 32 < nchunk; ++cg) {
        const int ch = cg * 32 + lane;
        const bool act = ch < nchunk;
        const int a0 = ch * 8;
        // CAT ('|', models/model.py:192-193): x = relu(cat[v', q']) has 2A channels, x_conv's weight is [G][2A]; the q' half
        // of x is the same for every position except for its dropout mask (element index (b, s, A + a) of the 2A-wide row).
        constexpr bool CAT = OP == VQA_ATT_CAT;
        const int W = CAT ? 2 * A : A;                 // row length of wx and of the dropout index space
        float qv[8], wv[G][8], wq[CAT ? G : 1][8];
        if (act) {
            load8f(qp + (int64_t)b * A + a0, qv);
#pragma unroll
            for (int g = 0; g < G; ++g) load8f(wx + (int64_t)g * W + a0, wv[g]);
            if (CAT) {
#pragma unroll
                for (int g = 0; g < G; ++g) load8f(wx + (int64_t)g * W + A + a0, wq[CAT ? g : 0]);
#pragma unroll
                for (int i = 0; i < 8; ++i) qv[i] = fmaxf(qv[i], 0.f);
            }
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
                drop8(d8, ((uint64_t)b * P + s0) * W + a0, m);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float pre = CAT ? x0[i] : (OP == VQA_ATT_ADD ? x0[i] + qv[i] : x0[i] * qv[i]);
                    const float r = fmaxf(pre, 0.f) * m[i];
#pragma unroll
                    for (int g = 0; g < G; ++g) acc0[g] = fmaf(r, wv[g][i], acc0[g]);
                }
                if (CAT) {
                    drop8(d8, ((uint64_t)b * P + s0) * W + A + a0, m);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int g = 0; g < G; ++g) acc0[g] = fmaf(qv[i] * m[i], wq[CAT ? g : 0][i], acc0[g]);
                }
                if (has1) {
                    drop8(d8, ((uint64_t)b * P + s1) * W + a0, m);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float pre = CAT ? x1[i] : (OP == VQA_ATT_ADD ? x1[i] + qv[i] : x1[i] * qv[i]);
                        const float r = fmaxf(pre, 0.f) * m[i];
#pragma unroll
                        for (int g = 0; g < G; ++g) acc1[g] = fmaf(r, wv[g][i], acc1[g]);
                    }
                    if (CAT) {
                        drop8(d8, ((uint64_t)b * P + s1) * W + A + a0, m);
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int g = 0; g < G; ++g) acc1[g] = fmaf(qv[i] * m[i], wq[CAT ? g : 0][i], acc1[g]);
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
    pdl_trigger();
    pdl_wait();
    extern __shared__ float sm[];
    float* pr = sm;                       // [G][P] softmax
    float* dl = pr + G * P;               // [G][P] dp, then dlogit
    float* dsm = dl + G * P;              // [G][C] upstream gradient
    float* red = dsm + G * C;             // [NW][(1+G)*256]   (CAT: [NW][(1+2G)*256])
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const Dropout8 d8 = make_dropout8(drop, SITE_ATT_X);
    constexpr bool CAT = OP == VQA_ATT_CAT;
    constexpr int NR = CAT ? 1 + 2 * G : 1 + G;       // reduced rows per chunk: dq, dwx (v' half) and for CAT dwx (q' half)
    const int W = CAT ? 2 * A : A;

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
        float qv[8], wv[G][8], dq[8], dw[G][8], wq[CAT ? G : 1][8], dwq[CAT ? G : 1][8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { dq[i] = 0.f; qv[i] = 0.f; }
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) { dw[g][i] = 0.f; wv[g][i] = 0.f; }
#pragma unroll
        for (int g = 0; g < (CAT ? G : 1); ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) { dwq[g][i] = 0.f; wq[g][i] = 0.f; }
        if (act) {
            load8f(qp + (int64_t)b * A + a0, qv);
#pragma unroll
            for (int g = 0; g < G; ++g) load8f(wx + (int64_t)g * W + a0, wv[g]);
            if (CAT) {
#pragma unroll
                for (int g = 0; g < G; ++g) load8f(wx + (int64_t)g * W + A + a0, wq[CAT ? g : 0]);
            }
            for (int s = warp; s < P; s += NW) {
                float x[8], m[8], o[8], dls[G];
                load8(vpb + (int64_t)s * A + a0, x);
                drop8(d8, ((uint64_t)b * P + s) * W + a0, m);
#pragma unroll
                for (int g = 0; g < G; ++g) dls[g] = dl[g * P + s];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float pre = CAT ? x[i] : (OP == VQA_ATT_ADD ? x[i] + qv[i] : x[i] * qv[i]);
                    const bool alive = pre > 0.f;
                    const float xr = alive ? pre * m[i] : 0.f;
                    float dxt = 0.f;
#pragma unroll
                    for (int g = 0; g < G; ++g) { dxt = fmaf(dls[g], wv[g][i], dxt); dw[g][i] = fmaf(dls[g], xr, dw[g][i]); }
                    const float dpre = alive ? dxt * m[i] : 0.f;
                    if (CAT || OP == VQA_ATT_ADD) { o[i] = dpre; if (!CAT) dq[i] += dpre; }
                    else { o[i] = dpre * qv[i]; dq[i] = fmaf(dpre, x[i], dq[i]); }
                }
                if (CAT) {                           // the q' half of x: relu(q') * dropout(b, s, A + a)
                    drop8(d8, ((uint64_t)b * P + s) * W + A + a0, m);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const bool alive = qv[i] > 0.f;
                        const float xr = alive ? qv[i] * m[i] : 0.f;
                        float dxt = 0.f;
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            dxt = fmaf(dls[g], wq[CAT ? g : 0][i], dxt);
                            dwq[CAT ? g : 0][i] = fmaf(dls[g], xr, dwq[CAT ? g : 0][i]);
                        }
                        if (alive) dq[i] = fmaf(dxt, m[i], dq[i]);
                    }
                }
                store8(dvpb + (int64_t)s * A + a0, o);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) red[(warp * NR) * 256 + lane * 8 + i] = dq[i];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) red[(warp * NR + 1 + g) * 256 + lane * 8 + i] = dw[g][i];
        if (CAT) {
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int i = 0; i < 8; ++i) red[(warp * NR + 1 + G + g) * 256 + lane * 8 + i] = dwq[CAT ? g : 0][i];
        }
        __syncthreads();
        for (int t = tid; t < NR * 256; t += NTHREADS) {
            const int k = t >> 8, a = cg * 256 + (t & 255);
            if (a < A) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) s += red[(w * NR + k) * 256 + (t & 255)];
                if (k == 0) dqp[(int64_t)b * A + a] = s;
                else if (k <= G) dwx_part[((int64_t)b * G + (k - 1)) * W + a] = s;
                else dwx_part[((int64_t)b * G + (k - 1 - G)) * W + A + a] = s;
            }
        }
        __syncthreads();
    }
}

// ================================================================================================================
// Streaming forward kernel for the tensor-core arm (bf16 activations, A = 1024, C = 256: the config.yaml shape).
//
// Persistent grid (one CTA per SM), samples round-robin.  A producer warp streams the sample's v' [P][A] and then
// its v [P][C] rows through a 16 x 8 KB shared-memory ring with cp.async.bulk + mbarriers and never stops at
// sample or phase boundaries, so HBM stays busy while the consumers reduce / normalise.  16 consumer warps work as
// 8 pairs; a pair owns every 8th 8 KB chunk (two private ring slots):
//   phase 1 (chunk = 4 positions x 1024 channels): each warp of the pair takes half a row (lane = 16 channels,
//           two conflict-free LDS.128 per position), computes relu(v' (+|*) q') & dropout-mask in packed bf16x2,
//           accumulates the G glimpse dot products in fp32 and folds its 4 x G per-lane partial sums with a
//           transposing butterfly (9 shuffles instead of 40).
//   softmax over the P positions per glimpse (warp shuffles), probabilities kept in shared memory and saved.
//   phase 3 (chunk = 16 positions x 256 channels): lane = 8 channels, fp32 accumulators, one cross-warp reduction.
// Algorithmic HBM bytes per sample as for the generic kernel; nothing is read twice.
//
// Sample order: both streaming kernels walk the batch from the LAST sample to the first.  The forward follows the
// v_conv GEMM, whose persistent tiles wrote v' in ascending row order: the tail of v' (about a third of its 354 MB at
// B = 256) is still dirty in the 126 MB L2 when this kernel starts, and reading it first turns those HBM reads into L2
// hits instead of forcing their write-back while the head is streamed in.  The backward writes dv' last-sample-first for
// the same reason: the v_conv data-gradient GEMM that follows reads dv' in ascending order.
// ================================================================================================================
namespace stream {

constexpr int A_ = 1024, C_ = 256;
constexpr int NCW = 16, NPAIR = 8;
constexpr int NTHR = (NCW + 1) * 32;
constexpr int CHUNK = 8192, NST = 16;
static_assert(NST % NPAIR == 0, "every ring slot must belong to exactly one consumer pair (see wait_chunk)");
constexpr int POS1 = CHUNK / (A_ * 2);           // 4 positions of v' per chunk
constexpr int POS3 = CHUNK / (C_ * 2);           // 16 positions of v per chunk
__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
// Slot = chunk % NST and pair = chunk % NPAIR with NST a multiple of NPAIR: a slot is only ever used by ONE pair, so
// when a pair asks for chunk cc it has already consumed the slot's previous chunk and a plain parity wait is exact.
// (With slots shared between pairs a fast pair could ask before the previous use of the slot has landed, and the
// 1-bit phase parity would alias with the phase before that.)
__device__ __forceinline__ void wait_chunk(uint64_t* full, uint32_t cc) { tc::mbar_wait(&full[cc % NST], (cc / NST) & 1); }
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCW * 32) : "memory"); }
__device__ __forceinline__ uint32_t hadd2_bf16(uint32_t a, uint32_t b) { uint32_t r; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hmul2_bf16(uint32_t a, uint32_t b) { uint32_t r; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t hrelu2_bf16(uint32_t a) { uint32_t r; asm("max.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0u)); return r; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t x) { return __uint_as_float(x << 16); }
__device__ __forceinline__ float bf_hi(uint32_t x) { return __uint_as_float(x & 0xffff0000u); }

// ---- v' (and the packed copy of q') as bf16x2 (H = false) or f16x2 (H = true).  The '+' fusion rounds v' + q' to the
// 16-bit format; q' is an order of magnitude larger than the spatial variation of v' that the x_conv weight gradient
// sees (sum_s dlogit[s] = 0 removes the constant part), so the 8-bit mantissa of bf16 puts 1-2 % of max-norm error on
// that one gradient.  With v' written as fp16 by the v_conv GEMM (11-bit mantissa; |v'| <= ||W_row||, nowhere near the
// fp16 range, and the GEMM saturates instead of producing infinities) the same instructions are 8 x more precise, and
// the widening to fp32 moves from the integer pipe (shift / and) to the half-precision pipe.
template <bool H> __device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    uint32_t r;
    if (H) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool H> __device__ __forceinline__ float2 widen_h2(uint32_t x) {
    if (H) {
        float2 f;
        asm("{\n.reg .f16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}" : "=f"(f.x), "=f"(f.y) : "r"(x));
        return f;
    }
    return make_float2(bf_lo(x), bf_hi(x));
}
// relu(a + b) / relu(a * b) in ONE half-precision-pipe instruction (fma.rn.relu): the same single rounding as add.rn /
// mul.rn followed by max(., 0), without the min/max instruction on the integer pipe
template <bool H> __device__ __forceinline__ uint32_t add_relu_h2(uint32_t a, uint32_t b) {
    uint32_t r;
    if (H) asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x3C003C00u), "r"(b));
    else asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x3F803F80u), "r"(b));
    return r;
}
template <bool H> __device__ __forceinline__ uint32_t mul_relu_h2(uint32_t a, uint32_t b) {
    uint32_t r;
    if (H) asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(0x80008000u));
    else asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(0x80008000u));
    return r;
}
template <bool H> __device__ __forceinline__ uint32_t ne_mask_h2(uint32_t a) {      // 0xFFFF in each half that is non-zero
    uint32_t r;
    if (H) asm("set.ne.u32.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0u));
    else asm("set.ne.u32.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(0u));
    return r;
}

// fold n = 8*GP per-lane values over the 32 lanes; afterwards v[0] of lane l holds the total of value index
// l >> (5 - log2 n) (the other lanes of that group hold the same total)
template <int N>
__device__ __forceinline__ void transpose_reduce(float (&v)[N], int lane) {
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = up ? v[n / 2 + i] : v[i];
            const float send = up ? v[i] : v[n / 2 + i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
#pragma unroll
    for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}

template <int G, int OP, bool TRAIN, bool VPH>
__global__ void __launch_bounds__(NTHR, 1)
attention_fwd_stream_kernel(const bf16* __restrict__ vp, const float* __restrict__ qp, const bf16* __restrict__ vn,
                            const float* __restrict__ wx, const float* __restrict__ bx, float* __restrict__ prob,
                            bf16* __restrict__ out, int64_t ldo, int B, int P, Dropout drop) {
    pdl_trigger();
    pdl_wait();
    constexpr int GP = G <= 2 ? G : 4;                 // glimpses padded to a power of two for the butterfly
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (tc::smem_u32(smem_raw) & 127u)) & 127u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* ring = smem;                                                    // NST x CHUNK
    float* red = reinterpret_cast<float*>(ring + NST * CHUNK);              // [NCW][G][C_]
    float* lpart = red + NCW * G * C_;                                       // [2 halves][G][P]
    float* pr = lpart + 2 * G * P;                                           // [G][P] softmax
    uint64_t* full = reinterpret_cast<uint64_t*>(pr + G * P + ((G * P) & 1));
    uint64_t* empty = full + NST;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n1 = (P + POS1 - 1) / POS1, n3 = (P + POS3 - 1) / POS3;

    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 2); }
        tc::fence_barrier_init();
    }
    __syncthreads();

    if (warp == NCW) {
        // ===== producer: one thread streams every chunk of every sample of this CTA, in consumption order =====
        if (lane == 0) {
            uint32_t c = 0;
            for (int bi = blockIdx.x; bi < B; bi += gridDim.x) {
        const int b = B - 1 - bi;                              // last sample first: see "sample order" at the top of the streaming kernels
                const uint8_t* src1 = reinterpret_cast<const uint8_t*>(vp + (int64_t)b * P * A_);
                const uint8_t* src3 = reinterpret_cast<const uint8_t*>(vn + (int64_t)b * P * C_);
                const int64_t bytes1 = (int64_t)P * A_ * 2, bytes3 = (int64_t)P * C_ * 2;
                for (int i = 0; i < n1 + n3; ++i, ++c) {
                    const int slot = c % NST;
                    tc::mbar_wait(&empty[slot], ((c / NST) & 1) ^ 1);
                    const bool ph1 = i < n1;
                    const int64_t off = (int64_t)(ph1 ? i : i - n1) * CHUNK;
                    const int64_t left = (ph1 ? bytes1 : bytes3) - off;
                    const uint32_t bytes = (uint32_t)(left < CHUNK ? left : CHUNK);
                    tc::mbar_expect_tx(&full[slot], bytes);
                    bulk_g2s(ring + slot * CHUNK, (ph1 ? src1 : src3) + off, bytes, &full[slot]);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const int pair = warp >> 1, half = warp & 1;
    const Dropout8 d8 = make_dropout8(drop, SITE_ATT_X);
    const float wscale = TRAIN ? d8.scale : 1.f;
    // lane's channels in phase 1: half*512 + (j*32 + lane)*8 + [0,8), j = 0,1
    // as (even, odd) channel pairs: one packed FFMA2 (fma.rn.f32x2) per bf16x2 word and glimpse instead of two FFMA
    float2 wv[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 w2 = *reinterpret_cast<const float2*>(wx + g * A_ + half * 512 + (j * 32 + lane) * 8 + 2 * i);
                wv[g][j * 4 + i] = make_float2(w2.x * wscale, w2.y * wscale);
            }

    uint32_t c = 0;                                   // global chunk counter (same sequence as the producer)
    for (int bi = blockIdx.x; bi < B; bi += gridDim.x) {
        const int b = B - 1 - bi;                              // last sample first: see "sample order" at the top of the streaming kernels
        uint32_t q2[8];                               // q' of this sample, rounded to the format of v' (packed pairs)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = *reinterpret_cast<const float2*>(qp + (int64_t)b * A_ + half * 512 + (j * 32 + lane) * 8 + 2 * i);
                q2[j * 4 + i] = pack_h2<VPH>(f.x, f.y);
            }
        // ---- phase 1: per-position glimpse logits
        for (int i = (int)((pair + NPAIR - (c % NPAIR)) % NPAIR); i < n1; i += NPAIR) {
            const uint32_t cc = c + i;
            const int slot = cc % NST;
            wait_chunk(full, cc);
            const uint8_t* base = ring + slot * CHUNK + half * 1024 + lane * 16;
            const int pos0 = i * POS1;
            float2 acc2[POS1 * GP];                 // .x: even channels, .y: odd channels; folded before the butterfly
#pragma unroll
            for (int k = 0; k < POS1 * GP; ++k) acc2[k] = make_float2(0.f, 0.f);
#pragma unroll
            for (int ps = 0; ps < POS1; ++ps) {
                if (pos0 + ps < P) {                 // warp-uniform
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint4 u = *reinterpret_cast<const uint4*>(base + ps * (A_ * 2) + j * 512);
                        uint32_t x[4] = {u.x, u.y, u.z, u.w};
                        uint32_t t[4];
                        if (TRAIN) dropout_flags8(d8, (uint32_t)(((int64_t)b * P + pos0 + ps) * (A_ / 8) + half * 64 + j * 32 + lane), t);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            uint32_t r = OP == VQA_ATT_ADD ? add_relu_h2<VPH>(x[e], q2[j * 4 + e]) : mul_relu_h2<VPH>(x[e], q2[j * 4 + e]);
                            if (TRAIN) r &= dropout_mask_bf16x2(t[e]);
                            const float2 lh = widen_h2<VPH>(r);
#pragma unroll
                            for (int g = 0; g < G; ++g) acc2[g * POS1 + ps] = __ffma2_rn(lh, wv[g][j * 4 + e], acc2[g * POS1 + ps]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&empty[slot]);          // this warp is done with the chunk
            float acc[POS1 * GP];
#pragma unroll
            for (int k = 0; k < POS1 * GP; ++k) acc[k] = acc2[k].x + acc2[k].y;
            transpose_reduce<POS1 * GP>(acc, lane);
            // value index k = g*POS1 + ps sits in lanes with (lane >> SH) == k
            constexpr int SH = 5 - ilog2(POS1 * GP);
            const int k = lane >> SH, g = k / POS1, ps = k % POS1;
            if ((lane & ((1 << SH) - 1)) == 0 && g < G && pos0 + ps < P) lpart[(half * G + g) * P + pos0 + ps] = acc[0];
        }
        c += n1;
        consumer_sync();

        // ---- phase 2: spatial softmax per glimpse (warp-shuffle reductions)
        if (warp < G) {
            const int g = warp;
            const float bias = bx[g];
            float mx = -INFINITY;
            for (int s = lane; s < P; s += 32) {
                const float l = lpart[g * P + s] + lpart[(G + g) * P + s] + bias;
                pr[g * P + s] = l;
                mx = fmaxf(mx, l);
            }
            mx = warp_max(mx);
            float sum = 0.f;
            for (int s = lane; s < P; s += 32) { const float e = __expf(pr[g * P + s] - mx); pr[g * P + s] = e; sum += e; }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
            for (int s = lane; s < P; s += 32) {
                const float pv = pr[g * P + s] * inv;
                pr[g * P + s] = pv;
                if (prob) prob[((int64_t)b * G + g) * P + s] = pv;
            }
        }
        consumer_sync();

        // ---- phase 3: out[g][c] = sum_s p[g][s] * vn[s][c]; lane = 8 channels, the pair's warps split the rows
        float2 o[G][4];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[g][e] = make_float2(0.f, 0.f);
        for (int i = (int)((pair + NPAIR - (c % NPAIR)) % NPAIR); i < n3; i += NPAIR) {
            const uint32_t cc = c + i;
            const int slot = cc % NST;
            wait_chunk(full, cc);
            const uint8_t* base = ring + slot * CHUNK + lane * 16;
            const int pos0 = i * POS3 + half * (POS3 / 2);
#pragma unroll 4
            for (int ps = 0; ps < POS3 / 2; ++ps) {
                const int s = pos0 + ps;
                if (s < P) {
                    const uint4 u = *reinterpret_cast<const uint4*>(base + (half * (POS3 / 2) + ps) * (C_ * 2));
                    const uint32_t x[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float pv = pr[g * P + s];
                        const float2 pv2 = make_float2(pv, pv);
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[g][e] = __ffma2_rn(make_float2(bf_lo(x[e]), bf_hi(x[e])), pv2, o[g][e]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&empty[slot]);
        }
        c += n3;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            *reinterpret_cast<float4*>(red + (warp * G + g) * C_ + lane * 8) = make_float4(o[g][0].x, o[g][0].y, o[g][1].x, o[g][1].y);
            *reinterpret_cast<float4*>(red + (warp * G + g) * C_ + lane * 8 + 4) = make_float4(o[g][2].x, o[g][2].y, o[g][3].x, o[g][3].y);
        }
        consumer_sync();
        for (int t = tid; t < G * C_; t += NCW * 32) {
            float sacc = 0.f;
#pragma unroll
            for (int w = 0; w < NCW; ++w) sacc += red[w * G * C_ + t];
            out[(int64_t)b * ldo + t] = __float2bfloat16_rn(sacc);
        }
        // `red`, `lpart` and `pr` are next written after the barriers of the next sample's phases 1 / 2 / 3
    }
}

// shared memory of the streaming forward kernel; it needs [3][G][P] floats beside the ring
static size_t fwd_stream_smem(int G, int P) {
    return (size_t)NST * CHUNK + sizeof(float) * ((size_t)NCW * G * C_ + 3 * (size_t)G * P + 2) + 2 * NST * 8 + 256;
}

template <int G, int OP, bool TRAIN, bool VPH>
int launch_fwd_stream_t(const void* vp, const float* qp, const void* vn, const float* wx, const float* bx, float* prob,
                        void* out, int64_t ldo, int B, int P, Dropout d, cudaStream_t st) {
    const size_t smem = fwd_stream_smem(G, P);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = B < sms ? B : sms;
    auto kern = attention_fwd_stream_kernel<G, OP, TRAIN, VPH>;
    VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VQA_CUDA(vqa_launch_pdl(kern, dim3(grid), dim3(NTHR), smem, st, (const bf16*)vp, qp, (const bf16*)vn, wx, bx, prob, (bf16*)out, ldo, B, P, d));
    VQA_CHECK_LAUNCH("attention_fwd_stream");
    return 0;
}
// vph: v' is fp16 (written so by the v_conv GEMM), else bf16
template <int G, int OP>
int launch_fwd_stream(const void* vp, bool vph, const float* qp, const void* vn, const float* wx, const float* bx, float* prob,
                      void* out, int64_t ldo, int B, int P, Dropout d, cudaStream_t st) {
    const bool train = d.threshold != 0;
    if (vph) return train ? launch_fwd_stream_t<G, OP, true, true>(vp, qp, vn, wx, bx, prob, out, ldo, B, P, d, st)
                          : launch_fwd_stream_t<G, OP, false, true>(vp, qp, vn, wx, bx, prob, out, ldo, B, P, d, st);
    return train ? launch_fwd_stream_t<G, OP, true, false>(vp, qp, vn, wx, bx, prob, out, ldo, B, P, d, st)
                 : launch_fwd_stream_t<G, OP, false, false>(vp, qp, vn, wx, bx, prob, out, ldo, B, P, d, st);
}

// ----------------------------------------------------------------------------------------------------------------
// Streaming backward kernel (same arm / shape as the forward).  Every consumer warp takes part in every 32 KB chunk,
// in order (plain parity waits):
//   phase A (chunk = 64 rows of v): warp w owns rows w, w+16, ...; lane = 8 channels.  dp[g][s] = <dout[g], v[s]>
//           (transposing butterfly over 4 rows x G) and dv[s] = sum_g p[g][s] dout[g] is written straight to HBM.
//   phase B: softmax backward, dlogit = p (dp - <p, dp>), and the x_conv bias gradient.
//   phase C (chunk = 16 rows of v'): warp (cg, pi) owns channels [256 cg, +256) of rows pi, pi+4, ...; lane = 8
//           channels, so the dq' / dW_x partial sums are 8 + 8 G registers and only 4 copies meet in shared memory.
//           The fusion, ReLU gate and dropout mask are recomputed exactly as in the forward (packed bf16x2).
// ----------------------------------------------------------------------------------------------------------------
constexpr int BCHUNK = 32768, BNST = 5;
constexpr int BPOSA = BCHUNK / (C_ * 2);          // 64 rows of v per chunk
constexpr int BPOSC = BCHUNK / (A_ * 2);          // 16 rows of v' per chunk

template <int G, int OP, bool TRAIN, bool VPH>
__global__ void __launch_bounds__(NTHR, 1)
attention_bwd_stream_kernel(const bf16* __restrict__ dout, int64_t ldd, const bf16* __restrict__ vp, const float* __restrict__ qp,
                            const bf16* __restrict__ vn, const float* __restrict__ wx, const float* __restrict__ prob,
                            bf16* __restrict__ dvp, bf16* __restrict__ dvn, float* __restrict__ dqp, float* __restrict__ dwx_part,
                            float* __restrict__ dbx_part, int B, int P, Dropout drop) {
    pdl_trigger();
    pdl_wait();
    constexpr int GP = G <= 2 ? G : 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (tc::smem_u32(smem_raw) & 127u)) & 127u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* ring = smem;                                                    // BNST x BCHUNK
    float* red = reinterpret_cast<float*>(ring + BNST * BCHUNK);            // [4 pi][1 + G][A_]
    float* pr = red + 4 * (1 + G) * A_;                                      // [P][G] softmax (interleaved)
    float* dl = pr + G * P;                                                  // [P][G] dp, then dlogit
    float* dsm = dl + G * P;                                                 // [G][C_] upstream gradient
    uint64_t* full = reinterpret_cast<uint64_t*>(dsm + G * C_);
    uint64_t* empty = full + BNST;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nA = (P + BPOSA - 1) / BPOSA, nC = (P + BPOSC - 1) / BPOSC;

    if (tid == 0) {
        for (int i = 0; i < BNST; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], NCW); }
        tc::fence_barrier_init();
    }
    __syncthreads();

    if (warp == NCW) {
        if (lane == 0) {
            uint32_t c = 0;
            for (int bi = blockIdx.x; bi < B; bi += gridDim.x) {
        const int b = B - 1 - bi;                              // last sample first: see "sample order" at the top of the streaming kernels
                const uint8_t* srcA = reinterpret_cast<const uint8_t*>(vn + (int64_t)b * P * C_);
                const uint8_t* srcC = reinterpret_cast<const uint8_t*>(vp + (int64_t)b * P * A_);
                const int64_t bytesA = (int64_t)P * C_ * 2, bytesC = (int64_t)P * A_ * 2;
                for (int i = 0; i < nA + nC; ++i, ++c) {
                    const int slot = c % BNST;
                    tc::mbar_wait(&empty[slot], ((c / BNST) & 1) ^ 1);
                    const bool phA = i < nA;
                    const int64_t off = (int64_t)(phA ? i : i - nA) * BCHUNK;
                    const int64_t left = (phA ? bytesA : bytesC) - off;
                    const uint32_t bytes = (uint32_t)(left < BCHUNK ? left : BCHUNK);
                    tc::mbar_expect_tx(&full[slot], bytes);
                    bulk_g2s(ring + slot * BCHUNK, (phA ? srcA : srcC) + off, bytes, &full[slot]);
                }
            }
        }
        return;
    }

    const Dropout8 d8 = make_dropout8(drop, SITE_ATT_X);
    const float scale = TRAIN ? d8.scale : 1.f;
    const int cg = warp & 3, pi = warp >> 2;
    const int a0 = cg * 256 + lane * 8;               // lane's 8 channels of v' in phase C
    float wv[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int i = 0; i < 8; ++i) wv[g][i] = wx[g * A_ + a0 + i] * scale;

    uint32_t c = 0;
    for (int bi = blockIdx.x; bi < B; bi += gridDim.x) {
        const int b = B - 1 - bi;                              // last sample first: see "sample order" at the top of the streaming kernels
        // per-sample inputs: probabilities, upstream gradient
        for (int i = tid; i < G * P; i += NCW * 32) { const int g = i / P, sidx = i - g * P; pr[sidx * G + g] = prob[(int64_t)b * G * P + i]; }
        for (int i = tid; i < G * C_; i += NCW * 32) dsm[i] = __bfloat162float(dout[(int64_t)b * ldd + i]);
        consumer_sync();

        // ---- phase A
        {
            float dv[G][8];
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int i = 0; i < 8; ++i) dv[g][i] = dsm[g * C_ + lane * 8 + i];
            bf16* dvnb = dvn + (int64_t)b * P * C_;
            for (int i = 0; i < nA; ++i, ++c) {
                const int slot = c % BNST;
                tc::mbar_wait(&full[slot], (c / BNST) & 1);
                const uint8_t* base = ring + slot * BCHUNK + lane * 16;
                float part[4 * GP];
#pragma unroll
                for (int k = 0; k < 4 * GP; ++k) part[k] = 0.f;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int r = warp + NCW * t, sidx = i * BPOSA + r;
                    if (sidx < P) {
                        const uint4 u = *reinterpret_cast<const uint4*>(base + r * (C_ * 2));
                        const uint32_t x[4] = {u.x, u.y, u.z, u.w};
                        float o[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
                        for (int g = 0; g < G; ++g) {
                            const float pg = pr[sidx * G + g];
                            float a = 0.f;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                a = fmaf(dv[g][2 * e], bf_lo(x[e]), a);
                                a = fmaf(dv[g][2 * e + 1], bf_hi(x[e]), a);
                                o[2 * e] = fmaf(pg, dv[g][2 * e], o[2 * e]);
                                o[2 * e + 1] = fmaf(pg, dv[g][2 * e + 1], o[2 * e + 1]);
                            }
                            part[g * 4 + t] = a;
                        }
                        uint4 w4;
                        w4.x = pack_bf16x2(o[0], o[1]); w4.y = pack_bf16x2(o[2], o[3]);
                        w4.z = pack_bf16x2(o[4], o[5]); w4.w = pack_bf16x2(o[6], o[7]);
                        __stcs(reinterpret_cast<uint4*>(dvnb + (int64_t)sidx * C_ + lane * 8), w4);
                    }
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&empty[slot]);
                transpose_reduce<4 * GP>(part, lane);
                constexpr int SH = 5 - ilog2(4 * GP);                    // 4*GP values -> index = lane >> SH
                const int k = lane >> SH, g = k >> 2, t = k & 3;
                const int sidx = i * BPOSA + warp + NCW * t;
                if ((lane & ((1 << SH) - 1)) == 0 && g < G && sidx < P) dl[sidx * G + g] = part[0];
            }
        }
        consumer_sync();

        // ---- phase B: softmax backward per glimpse
        if (warp < G) {
            const int g = warp;
            float dot = 0.f;
            for (int sidx = lane; sidx < P; sidx += 32) dot += pr[sidx * G + g] * dl[sidx * G + g];
            dot = warp_sum(dot);
            float sb = 0.f;
            for (int sidx = lane; sidx < P; sidx += 32) {
                const float d = pr[sidx * G + g] * (dl[sidx * G + g] - dot);
                dl[sidx * G + g] = d;
                sb += d;
            }
            sb = warp_sum(sb);
            if (lane == 0) dbx_part[(int64_t)b * G + g] = sb;
        }
        consumer_sync();

        // ---- phase C
        {
            uint32_t q2[4];
            float qf[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = *reinterpret_cast<const float2*>(qp + (int64_t)b * A_ + a0 + 2 * e);
                q2[e] = pack_h2<VPH>(f.x, f.y);
                const float2 qw = widen_h2<VPH>(q2[e]);
                qf[2 * e] = qw.x; qf[2 * e + 1] = qw.y;
            }
            float dq[8], dw[G][8];
#pragma unroll
            for (int e = 0; e < 8; ++e) dq[e] = 0.f;
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int e = 0; e < 8; ++e) dw[g][e] = 0.f;
            bf16* dvpb = dvp + (int64_t)b * P * A_;
            for (int i = 0; i < nC; ++i, ++c) {
                const int slot = c % BNST;
                tc::mbar_wait(&full[slot], (c / BNST) & 1);
                const uint8_t* base = ring + slot * BCHUNK + cg * 512 + lane * 16;
#pragma unroll
                for (int t = 0; t < BPOSC / 4; ++t) {
                    const int r = pi + 4 * t, sidx = i * BPOSC + r;
                    if (sidx < P) {
                        const uint4 u = *reinterpret_cast<const uint4*>(base + r * (A_ * 2));
                        const uint32_t x[4] = {u.x, u.y, u.z, u.w};
                        uint32_t fl[4];
                        if (TRAIN) dropout_flags8(d8, (uint32_t)(((int64_t)b * P + sidx) * (A_ / 8) + cg * 32 + lane), fl);
                        float dls[G];
#pragma unroll
                        for (int g = 0; g < G; ++g) dls[g] = dl[sidx * G + g];
                        uint4 o4;
                        uint32_t* o = reinterpret_cast<uint32_t*>(&o4);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            uint32_t xr = OP == VQA_ATT_ADD ? add_relu_h2<VPH>(x[e], q2[e]) : mul_relu_h2<VPH>(x[e], q2[e]);
                            if (TRAIN) xr &= dropout_mask_bf16x2(fl[e]);          // relu(pre) where kept (unscaled), else 0
                            const float2 xw = widen_h2<VPH>(xr);
                            const float xl = xw.x, xh = xw.y;
                            float dl_lo = 0.f, dl_hi = 0.f;
#pragma unroll
                            for (int g = 0; g < G; ++g) {
                                dl_lo = fmaf(dls[g], wv[g][2 * e], dl_lo);
                                dl_hi = fmaf(dls[g], wv[g][2 * e + 1], dl_hi);
                                dw[g][2 * e] = fmaf(dls[g], xl, dw[g][2 * e]);
                                dw[g][2 * e + 1] = fmaf(dls[g], xh, dw[g][2 * e + 1]);
                            }
                            // gradient w.r.t. the pre-activation, gated by (alive and kept) <=> xr != 0
                            const uint32_t dpre = pack_bf16x2(dl_lo, dl_hi) & ne_mask_h2<VPH>(xr);
                            const float pl = bf_lo(dpre), phh = bf_hi(dpre);
                            if (OP == VQA_ATT_ADD) {
                                o[e] = dpre;
                                dq[2 * e] += pl; dq[2 * e + 1] += phh;
                            } else {
                                o[e] = pack_bf16x2(pl * qf[2 * e], phh * qf[2 * e + 1]);
                                const float2 vw = widen_h2<VPH>(x[e]);
                                dq[2 * e] = fmaf(pl, vw.x, dq[2 * e]);
                                dq[2 * e + 1] = fmaf(phh, vw.y, dq[2 * e + 1]);
                            }
                        }
                        __stcs(reinterpret_cast<uint4*>(dvpb + (int64_t)sidx * A_ + a0), o4);
                    }
                }
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&empty[slot]);
            }
            // four row-interleaved copies of every channel's partial sums meet in shared memory
            float* rp = red + pi * (1 + G) * A_ + a0;
            *reinterpret_cast<float4*>(rp) = make_float4(dq[0], dq[1], dq[2], dq[3]);
            *reinterpret_cast<float4*>(rp + 4) = make_float4(dq[4], dq[5], dq[6], dq[7]);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                *reinterpret_cast<float4*>(rp + (1 + g) * A_) = make_float4(dw[g][0], dw[g][1], dw[g][2], dw[g][3]);
                *reinterpret_cast<float4*>(rp + (1 + g) * A_ + 4) = make_float4(dw[g][4], dw[g][5], dw[g][6], dw[g][7]);
            }
        }
        consumer_sync();
        for (int t = tid; t < (1 + G) * A_; t += NCW * 32) {
            const float sacc = red[t] + red[(1 + G) * A_ + t] + red[2 * (1 + G) * A_ + t] + red[3 * (1 + G) * A_ + t];
            const int k = t / A_, a = t - k * A_;
            if (k == 0) dqp[(int64_t)b * A_ + a] = sacc;
            else dwx_part[((int64_t)b * G + (k - 1)) * A_ + a] = sacc * scale;       // dW_x sees the scaled, dropped x
        }
        // the next sample's first consumer_sync (after staging pr / dsm) separates these reads from the next writes of `red`
    }
}

static size_t bwd_stream_smem(int G, int P) {
    return (size_t)BNST * BCHUNK + sizeof(float) * ((size_t)4 * (1 + G) * A_ + 2 * (size_t)G * P + (size_t)G * C_) + 2 * BNST * 8 + 256;
}

template <int G, int OP, bool TRAIN, bool VPH>
int launch_bwd_stream_t(const void* dout, int64_t ldd, const void* vp, const float* qp, const void* vn, const float* wx,
                        const float* prob, void* dvp, void* dvn, float* dqp, float* dwx_part, float* dbx_part,
                        int B, int P, Dropout d, cudaStream_t st) {
    const size_t smem = bwd_stream_smem(G, P);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = B < sms ? B : sms;
    auto kern = attention_bwd_stream_kernel<G, OP, TRAIN, VPH>;
    VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VQA_CUDA(vqa_launch_pdl(kern, dim3(grid), dim3(NTHR), smem, st, (const bf16*)dout, ldd, (const bf16*)vp, qp, (const bf16*)vn, wx, prob, (bf16*)dvp,
                            (bf16*)dvn, dqp, dwx_part, dbx_part, B, P, d));
    VQA_CHECK_LAUNCH("attention_bwd_stream");
    return 0;
}
template <int G, int OP>
int launch_bwd_stream(const void* dout, int64_t ldd, const void* vp, bool vph, const float* qp, const void* vn, const float* wx,
                      const float* prob, void* dvp, void* dvn, float* dqp, float* dwx_part, float* dbx_part,
                      int B, int P, Dropout d, cudaStream_t st) {
    const bool train = d.threshold != 0;
    if (vph) return train ? launch_bwd_stream_t<G, OP, true, true>(dout, ldd, vp, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, d, st)
                          : launch_bwd_stream_t<G, OP, false, true>(dout, ldd, vp, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, d, st);
    return train ? launch_bwd_stream_t<G, OP, true, false>(dout, ldd, vp, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, d, st)
                 : launch_bwd_stream_t<G, OP, false, false>(dout, ldd, vp, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, d, st);
}

}  // namespace stream

template <typename T, int G, int OP>
int launch_fwd(const void* vp, const float* qp, const void* vn, const float* wx, const float* bx, float* prob,
               void* out, int64_t ldo, int B, int P, int A, int C, Dropout d, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)G * P + (size_t)NW * G * 256);
    auto kern = attention_fwd_kernel<T, G, OP>;
    if (smem > 48 * 1024) VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VQA_CUDA(vqa_launch_pdl(kern, dim3(B), dim3(NTHREADS), smem, st, (const T*)vp, qp, (const T*)vn, wx, bx, prob, (T*)out, ldo, P, A, C, d));
    VQA_CHECK_LAUNCH("attention_fwd");
    return 0;
}

template <typename T, int G, int OP>
int launch_bwd(const void* dout, int64_t ldd, const void* vp, const float* qp, const void* vn, const float* wx,
               const float* prob, void* dvp, void* dvn, float* dqp, float* dwx_part, float* dbx_part,
               int B, int P, int A, int C, Dropout d, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)2 * G * P + (size_t)G * C + (size_t)NW * (OP == VQA_ATT_CAT ? 1 + 2 * G : 1 + G) * 256);
    auto kern = attention_bwd_kernel<T, G, OP>;
    if (smem > 48 * 1024) VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VQA_CUDA(vqa_launch_pdl(kern, dim3(B), dim3(NTHREADS), smem, st, (const T*)dout, ldd, (const T*)vp, qp, (const T*)vn, wx, prob, (T*)dvp, (T*)dvn,
                                    dqp, dwx_part, dbx_part, P, A, C, d));
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
            case 12:  return FN<float, 1, VQA_ATT_CAT>(__VA_ARGS__);                                 \
            case 20:  return FN<float, 2, VQA_ATT_ADD>(__VA_ARGS__);                                 \
            case 21:  return FN<float, 2, VQA_ATT_MUL>(__VA_ARGS__);                                 \
            case 22:  return FN<float, 2, VQA_ATT_CAT>(__VA_ARGS__);                                 \
            case 30:  return FN<float, 3, VQA_ATT_ADD>(__VA_ARGS__);                                 \
            case 31:  return FN<float, 3, VQA_ATT_MUL>(__VA_ARGS__);                                 \
            case 32:  return FN<float, 3, VQA_ATT_CAT>(__VA_ARGS__);                                 \
            case 40:  return FN<float, 4, VQA_ATT_ADD>(__VA_ARGS__);                                 \
            case 41:  return FN<float, 4, VQA_ATT_MUL>(__VA_ARGS__);                                 \
            case 42:  return FN<float, 4, VQA_ATT_CAT>(__VA_ARGS__);                                 \
            case 110: return FN<bf16, 1, VQA_ATT_ADD>(__VA_ARGS__);                                  \
            case 111: return FN<bf16, 1, VQA_ATT_MUL>(__VA_ARGS__);                                  \
            case 112: return FN<bf16, 1, VQA_ATT_CAT>(__VA_ARGS__);                                  \
            case 120: return FN<bf16, 2, VQA_ATT_ADD>(__VA_ARGS__);                                  \
            case 121: return FN<bf16, 2, VQA_ATT_MUL>(__VA_ARGS__);                                  \
            case 122: return FN<bf16, 2, VQA_ATT_CAT>(__VA_ARGS__);                                  \
            case 130: return FN<bf16, 3, VQA_ATT_ADD>(__VA_ARGS__);                                  \
            case 131: return FN<bf16, 3, VQA_ATT_MUL>(__VA_ARGS__);                                  \
            case 132: return FN<bf16, 3, VQA_ATT_CAT>(__VA_ARGS__);                                  \
            case 140: return FN<bf16, 4, VQA_ATT_ADD>(__VA_ARGS__);                                  \
            case 141: return FN<bf16, 4, VQA_ATT_MUL>(__VA_ARGS__);                                  \
            case 142: return FN<bf16, 4, VQA_ATT_CAT>(__VA_ARGS__);                                  \
        }                                                                                            \
    } while (0)

static int att_check(int act_dtype, int op, int B, int P, int A, int C, int G, int64_t ld) {
    VQA_REQUIRE(act_dtype == VQA_F32 || act_dtype == VQA_BF16, "attention: bad dtype %d", act_dtype);
    VQA_REQUIRE(op == VQA_ATT_ADD || op == VQA_ATT_MUL || op == VQA_ATT_CAT, "attention: do_option code %d not supported", op);
    VQA_REQUIRE(B > 0 && P > 0 && A > 0 && C > 0, "attention: bad dims");
    VQA_REQUIRE(G >= 1 && G <= 4, "attention: glimpses=%d not supported (1..4)", G);
    VQA_REQUIRE(A % 8 == 0 && C % 8 == 0, "attention: hidden_dim (%d) and image features (%d) must be multiples of 8", A, C);
    VQA_REQUIRE(ld >= (int64_t)G * C, "attention: row pitch %lld < G*C", (long long)ld);
    VQA_REQUIRE((size_t)(2 * G * P + G * C + NW * (op == VQA_ATT_CAT ? 1 + 2 * G : 1 + G) * 256) * 4 <= 200 * 1024, "attention: spatial grid too large for shared memory");
    return 0;
}

// 1 when the streaming kernels take this configuration (the shapes v' may be handed over as fp16 for), else 0
extern "C" int vqa_attention_streaming_ok(int act_dtype, int op, int P, int A, int C, int G) {
    return act_dtype == VQA_BF16 && (op == VQA_ATT_ADD || op == VQA_ATT_MUL) && A == stream::A_ && C == stream::C_ && G >= 1 && G <= 2 &&
           P > 0 && stream::fwd_stream_smem(G, P) <= 232448 && stream::bwd_stream_smem(G, P) <= 232448;
}

extern "C" int vqa_attention_fwd_x(const void* vp, int vp_dtype, const float* qp, const void* vn, const float* wx, const float* bx,
                                   float* prob, void* out, int64_t ldo, int act_dtype, int op,
                                   int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream) {
    if (int e = att_check(act_dtype, op, B, P, A, C, G, ldo)) return e;
    const Dropout d = make_dropout(seed, p_drop);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vph = vp_dtype == VQA_F16;
    VQA_REQUIRE(vph || vp_dtype == act_dtype, "attention: v' dtype %d (the activation dtype, or fp16 on the streaming kernels)", vp_dtype);
    if (vqa_attention_streaming_ok(act_dtype, op, P, A, C, G)) {     // tensor-core arm at the config.yaml shape
        VQA_REQUIRE((int64_t)B * P * (A / 8) < (1ll << 32), "attention: batch too large for the 32-bit dropout counter");
        if (G == 1) return op == VQA_ATT_ADD ? stream::launch_fwd_stream<1, VQA_ATT_ADD>(vp, vph, qp, vn, wx, bx, prob, out, ldo, B, P, d, st)
                                             : stream::launch_fwd_stream<1, VQA_ATT_MUL>(vp, vph, qp, vn, wx, bx, prob, out, ldo, B, P, d, st);
        return op == VQA_ATT_ADD ? stream::launch_fwd_stream<2, VQA_ATT_ADD>(vp, vph, qp, vn, wx, bx, prob, out, ldo, B, P, d, st)
                                 : stream::launch_fwd_stream<2, VQA_ATT_MUL>(vp, vph, qp, vn, wx, bx, prob, out, ldo, B, P, d, st);
    }
    VQA_REQUIRE(!vph, "attention: fp16 v' is only taken by the streaming kernels (see vqa_attention_streaming_ok)");
    ATT_DISPATCH(launch_fwd, vp, qp, vn, wx, bx, prob, out, ldo, B, P, A, C, d, st);
    VQA_REQUIRE(false, "attention_fwd: no kernel for this configuration");
    return 0;
}

extern "C" int vqa_attention_fwd(const void* vp, const float* qp, const void* vn, const float* wx, const float* bx,
                                 float* prob, void* out, int64_t ldo, int act_dtype, int op,
                                 int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream) {
    return vqa_attention_fwd_x(vp, act_dtype, qp, vn, wx, bx, prob, out, ldo, act_dtype, op, B, P, A, C, G, p_drop, seed, stream);
}

extern "C" int vqa_attention_bwd_x(const void* dout, int64_t ldd, const void* vp, int vp_dtype, const float* qp, const void* vn,
                                   const float* wx, const float* prob, void* dvp, void* dvn, float* dqp,
                                   float* dwx_part, float* dbx_part, int act_dtype, int op,
                                   int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream) {
    if (int e = att_check(act_dtype, op, B, P, A, C, G, ldd)) return e;
    const Dropout d = make_dropout(seed, p_drop);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vph = vp_dtype == VQA_F16;
    VQA_REQUIRE(vph || vp_dtype == act_dtype, "attention: v' dtype %d (the activation dtype, or fp16 on the streaming kernels)", vp_dtype);
    if (vqa_attention_streaming_ok(act_dtype, op, P, A, C, G)) {     // same condition as the forward: both recompute the fusion in 16 bits
        VQA_REQUIRE((int64_t)B * P * (A / 8) < (1ll << 32), "attention: batch too large for the 32-bit dropout counter");
        if (G == 1) return op == VQA_ATT_ADD ? stream::launch_bwd_stream<1, VQA_ATT_ADD>(dout, ldd, vp, vph, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, d, st)
                                             : stream::launch_bwd_stream<1, VQA_ATT_MUL>(dout, ldd, vp, vph, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, d, st);
        return op == VQA_ATT_ADD ? stream::launch_bwd_stream<2, VQA_ATT_ADD>(dout, ldd, vp, vph, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, d, st)
                                 : stream::launch_bwd_stream<2, VQA_ATT_MUL>(dout, ldd, vp, vph, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, d, st);
    }
    VQA_REQUIRE(!vph, "attention: fp16 v' is only taken by the streaming kernels (see vqa_attention_streaming_ok)");
    ATT_DISPATCH(launch_bwd, dout, ldd, vp, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, B, P, A, C, d, st);
    VQA_REQUIRE(false, "attention_bwd: no kernel for this configuration");
    return 0;
}

extern "C" int vqa_attention_bwd(const void* dout, int64_t ldd, const void* vp, const float* qp, const void* vn,
                                 const float* wx, const float* prob, void* dvp, void* dvn, float* dqp,
                                 float* dwx_part, float* dbx_part, int act_dtype, int op,
                                 int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream) {
    return vqa_attention_bwd_x(dout, ldd, vp, act_dtype, qp, vn, wx, prob, dvp, dvn, dqp, dwx_part, dbx_part, act_dtype, op,
                               B, P, A, C, G, p_drop, seed, stream);
}
