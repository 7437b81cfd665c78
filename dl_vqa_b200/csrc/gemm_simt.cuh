// fp32 SIMT implicit-GEMM core: C[m,n] = sum_k A(m,k) * B(n,k), with functor loaders and epilogues.
//
// This is the exact-fp32 arm of the hot path (the 1e-4 parity bar of BASELINE.json north_star) and
// the generic-shape arm (any channel list / stride / kernel size of config.yaml).  The bf16 arm for
// the config.yaml default shapes is the tcgen05 kernel family in gemm_tc.cu / conv_tc.cu.
//
// Tile 128x128x16, 256 threads, 8x8 outputs per thread as four 4x4 blocks
// (rows ty*4+{0..3} and 64+ty*4+{0..3}; cols tx*4+{0..3} and 64+tx*4+{0..3}) so that
//   * a 2x2 max-pool window (4 consecutive implicit-GEMM rows) lives in one thread, and
//   * the 4 LSTM gates of one hidden unit (4 consecutive columns) live in one thread.
#pragma once
#include "common.cuh"

namespace simt {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256, PAD = 4;

// ------------------------------------------------------------------------------------------
// loaders.  Protocol:  Row ctx;  init_row(ctx, r) ;  at(ctx, k) -> float (0 outside the matrix)
// ------------------------------------------------------------------------------------------
template <typename T>
struct DenseLoader {
    const T* p;
    int rows, K;
    int64_t sr, sk, sbatch;   // element strides
    int ilv_h;                // >0: logical row r maps to source row (r%4)*ilv_h + r/4 (LSTM gate interleave)
    struct Row { const T* p; bool ok; };
    __device__ void set_batch(int z) { p += (int64_t)z * sbatch; }
    __device__ void init_row(Row& c, int r) const {
        c.ok = r < rows;
        int src = r;
        if (ilv_h > 0) src = (r & 3) * ilv_h + (r >> 2);
        c.p = p + (int64_t)src * sr;
    }
    __device__ float at(const Row& c, int k) const {
        return (c.ok && k < K) ? to_f32(c.p[(int64_t)k * sk]) : 0.f;
    }
};

struct ConvGeom {
    int B, IH, IW, Cin, Cout, KS, stride, PH, PW;   // pooled output PH x PW; conv output 2PH x 2PW used
};

// conv forward A operand: row m = (b, ph, pw, e) with e = dy*2+dx the pool-window element,
// k = (kh, kw, ci).  TIn/NCHW describe the input tensor.
template <typename TIn, bool NCHW>
struct ConvFwdALoader {
    const TIn* x;
    ConvGeom g;
    int M, K;
    struct Row { int64_t base; bool ok; };
    __device__ void set_batch(int) {}
    __device__ void init_row(Row& c, int m) const {
        c.ok = m < M;
        if (!c.ok) { c.base = 0; return; }
        const int e = m & 3, p = m >> 2;
        const int pw = p % g.PW, t = p / g.PW, ph = t % g.PH, b = t / g.PH;
        const int ih = (2 * ph + (e >> 1)) * g.stride, iw = (2 * pw + (e & 1)) * g.stride;
        c.base = NCHW ? (((int64_t)b * g.Cin) * g.IH + ih) * g.IW + iw
                      : (((int64_t)b * g.IH + ih) * g.IW + iw) * g.Cin;
    }
    __device__ float at(const Row& c, int k) const {
        if (!c.ok || k >= K) return 0.f;
        const int tap = k / g.Cin, ci = k - tap * g.Cin;
        const int kh = tap / g.KS, kw = tap - kh * g.KS;
        const int64_t off = NCHW ? ((int64_t)ci * g.IH + kh) * g.IW + kw
                                 : ((int64_t)kh * g.IW + kw) * g.Cin + ci;
        return to_f32(x[c.base + off]);
    }
};

// conv weights read straight from the PyTorch OIHW tensor: B(n = co, k = (tap, ci))
struct ConvWLoader {
    const float* w;
    int Cout, Cin, KK, K;
    struct Row { const float* p; bool ok; };
    __device__ void set_batch(int) {}
    __device__ void init_row(Row& c, int n) const { c.ok = n < Cout; c.p = w + (int64_t)n * Cin * KK; }
    __device__ float at(const Row& c, int k) const {
        if (!c.ok || k >= K) return 0.f;
        const int tap = k / Cin, ci = k - tap * Cin;
        return c.p[ci * KK + tap];
    }
};

// gradient w.r.t. the conv output, un-pooled on the fly from (dpool, mask):
// value at conv-output position (b, oh, ow, co)
template <typename T>
struct UnpoolView {
    const T* dpool;          // [B, PH, PW, Cout]
    const uint8_t* mask;     // same shape; 0..3 = window element holding the max, 4 = ReLU-dead
    ConvGeom g;
    __device__ float at(int b, int oh, int ow, int co) const {
        if (oh < 0 || ow < 0 || oh >= 2 * g.PH || ow >= 2 * g.PW) return 0.f;
        const int64_t idx = (((int64_t)b * g.PH + (oh >> 1)) * g.PW + (ow >> 1)) * g.Cout + co;
        const int e = ((oh & 1) << 1) | (ow & 1);
        return mask[idx] == e ? to_f32(dpool[idx]) : 0.f;
    }
};

// conv dgrad A operand: row m = (b, ih, iw), k = (tap, co)
template <typename T>
struct ConvDgradALoader {
    UnpoolView<T> u;
    int M, K;
    struct Row { int b, ih, iw; bool ok; };
    __device__ void set_batch(int) {}
    __device__ void init_row(Row& c, int m) const {
        c.ok = m < M;
        const int iw = m % u.g.IW, t = m / u.g.IW;
        c.iw = iw; c.ih = t % u.g.IH; c.b = t / u.g.IH;
    }
    __device__ float at(const Row& c, int k) const {
        if (!c.ok || k >= K) return 0.f;
        const int tap = k / u.g.Cout, co = k - tap * u.g.Cout;
        const int kh = tap / u.g.KS, kw = tap - kh * u.g.KS;
        int oh = c.ih - kh, ow = c.iw - kw;
        if (oh < 0 || ow < 0) return 0.f;
        if (u.g.stride > 1) {
            if (oh % u.g.stride || ow % u.g.stride) return 0.f;
            oh /= u.g.stride; ow /= u.g.stride;
        }
        return u.at(c.b, oh, ow, co);
    }
};
// dgrad B operand from OIHW weights: B(n = ci, k = (tap, co))
struct ConvDgradWLoader {
    const float* w;
    int Cout, Cin, KK, K;
    struct Row { int ci; bool ok; };
    __device__ void set_batch(int) {}
    __device__ void init_row(Row& c, int n) const { c.ok = n < Cin; c.ci = n; }
    __device__ float at(const Row& c, int k) const {
        if (!c.ok || k >= K) return 0.f;
        const int tap = k / Cout, co = k - tap * Cout;
        return w[((int64_t)co * Cin + c.ci) * KK + tap];
    }
};

// conv wgrad: A(m = co, k = (p, e)),  B(n = (tap, ci), k = (p, e))
template <typename T>
struct ConvWgradALoader {
    const T* dpool;
    const uint8_t* mask;
    int Cout, K;
    struct Row { int co; bool ok; };
    __device__ void set_batch(int) {}
    __device__ void init_row(Row& c, int m) const { c.ok = m < Cout; c.co = m; }
    __device__ float at(const Row& c, int k) const {
        if (!c.ok || k >= K) return 0.f;
        const int64_t idx = (int64_t)(k >> 2) * Cout + c.co;
        return mask[idx] == (k & 3) ? to_f32(dpool[idx]) : 0.f;
    }
};
template <typename TIn, bool NCHW>
struct ConvWgradBLoader {
    const TIn* x;
    ConvGeom g;
    int N, K;
    struct Row { int kh, kw, ci; bool ok; };
    __device__ void set_batch(int) {}
    __device__ void init_row(Row& c, int n) const {
        c.ok = n < N;
        const int tap = n / g.Cin;
        c.ci = n - tap * g.Cin; c.kh = tap / g.KS; c.kw = tap - c.kh * g.KS;
    }
    __device__ float at(const Row& c, int k) const {
        if (!c.ok || k >= K) return 0.f;
        const int e = k & 3, p = k >> 2;
        const int pw = p % g.PW, t = p / g.PW, ph = t % g.PH, b = t / g.PH;
        const int ih = (2 * ph + (e >> 1)) * g.stride + c.kh, iw = (2 * pw + (e & 1)) * g.stride + c.kw;
        const int64_t off = NCHW ? (((int64_t)b * g.Cin + c.ci) * g.IH + ih) * g.IW + iw
                                 : (((int64_t)b * g.IH + ih) * g.IW + iw) * g.Cin + c.ci;
        return to_f32(x[off]);
    }
};

// ------------------------------------------------------------------------------------------
// epilogues.  Protocol: set_batch(z); apply(m0, n0, acc[4][4], M, N) for rows m0..m0+3, cols n0..n0+3
// ------------------------------------------------------------------------------------------
template <typename TOut>
struct EpStore {
    TOut* out;
    int64_t ldc, sbatch;
    const float* bias;      // [N] or null
    const float* bias2;     // [N] or null (second bias added, e.g. b_ih + b_hh)
    int64_t bias_sbatch;
    int relu;
    int accumulate;         // out += value
    int use_dropout;
    uint32_t site;
    Dropout drop;
    int ilv_h;              // >0: logical column n (gate-interleaved) is stored at column (n%4)*ilv_h + n/4
    __device__ void set_batch(int z) {
        out += (int64_t)z * sbatch;
        if (bias) bias += (int64_t)z * bias_sbatch;
        if (bias2) bias2 += (int64_t)z * bias_sbatch;
    }
    __device__ void apply(int m0, int n0, const float (&acc)[4][4], int M, int N) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + i;
            if (m >= M) break;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + j;
                if (n >= N) break;
                const int nc = ilv_h > 0 ? (n & 3) * ilv_h + (n >> 2) : n;
                float v = acc[i][j];
                if (bias) v += bias[nc];
                if (bias2) v += bias2[nc];
                if (relu) v = fmaxf(v, 0.f);
                if (use_dropout) v *= dropout_mult(drop, site, (uint64_t)m * N + nc);
                const int64_t o = (int64_t)m * ldc + nc;
                if (accumulate) v += to_f32(out[o]);
                out[o] = from_f32<TOut>(v);
            }
        }
    }
};

// conv forward: bias + ReLU + 2x2 max-pool over the 4 rows of the block (one pool window)
template <typename TOut>
struct EpPool {
    TOut* out;          // [M/4, N]
    uint8_t* mask;      // [M/4, N]
    const float* bias;
    __device__ void set_batch(int) {}
    __device__ void apply(int m0, int n0, const float (&acc)[4][4], int M, int N) const {
        if (m0 >= M) return;
        const int64_t p = m0 >> 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + j;
            if (n >= N) break;
            float best = acc[0][j];
            int arg = 0;
#pragma unroll
            for (int i = 1; i < 4; ++i)
                if (acc[i][j] > best) { best = acc[i][j]; arg = i; }
            best += bias[n];
            if (!(best > 0.f)) { best = 0.f; arg = 4; }
            out[p * N + n] = from_f32<TOut>(best);
            mask[p * N + n] = (uint8_t)arg;
        }
    }
};

// split-K accumulation into an fp32 matrix with arbitrary (row, col) -> offset mapping
struct EpAtomic {
    float* out;
    int64_t ldc, sbatch;
    int conv_w;     // 1: out is OIHW conv weight grad, m = co, n = tap*Cin + ci
    int Cin, KK;
    __device__ void set_batch(int z) { out += (int64_t)z * sbatch; }
    __device__ void apply(int m0, int n0, const float (&acc)[4][4], int M, int N) const {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + i;
            if (m >= M) break;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + j;
                if (n >= N) break;
                int64_t o;
                if (conv_w) { const int tap = n / Cin, ci = n - tap * Cin; o = ((int64_t)m * Cin + ci) * KK + tap; }
                else o = (int64_t)m * ldc + n;
                atomicAdd(out + o, acc[i][j]);
            }
        }
    }
};

// LSTM cell, forward step s.  Columns are gate-interleaved: n = 4*j + gate (i,f,g,o).
// All sequence buffers are step-indexed: [dir][s][b][...]  (see lstm.cu).
template <typename T>
struct EpLstmCell {
    T* gx;              // [dirs][T][B][4H]  in: x-projection + biases; out: activated gates (saved)
    float* cs;          // [dirs][T][B][H]   cell state after step s
    T* hs;              // [dirs][T][B][H]   hidden state after step s
    T* qf;              // [B][dirs*H]       final cell state (written at s == T-1)
    const int64_t* len; // [B]
    int s, T_, B, H, dirs;
    int dir;
    __device__ void set_batch(int z) { dir = z; }
    __device__ void apply(int m0, int n0, const float (&acc)[4][4], int M, int N) const {
        const int j = n0 >> 2;
        if (j >= H) return;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int b = m0 + i;
            if (b >= B) break;
            const int64_t row = ((int64_t)dir * T_ + s) * B + b;
            const int64_t prow = row - B;
            const float c_prev = s > 0 ? cs[prow * H + j] : 0.f;
            float c_new, h_new;
            if (s < (int)len[b]) {
                T* g = gx + row * 4 * H;
                const float gi = sigmoidf_(acc[i][0] + to_f32(g[j]));
                const float gf = sigmoidf_(acc[i][1] + to_f32(g[H + j]));
                const float gg = tanhf(acc[i][2] + to_f32(g[2 * H + j]));
                const float go = sigmoidf_(acc[i][3] + to_f32(g[3 * H + j]));
                c_new = gf * c_prev + gi * gg;
                h_new = go * tanhf(c_new);
                g[j] = from_f32<T>(gi); g[H + j] = from_f32<T>(gf);
                g[2 * H + j] = from_f32<T>(gg); g[3 * H + j] = from_f32<T>(go);
            } else {
                c_new = c_prev;
                h_new = s > 0 ? to_f32(hs[prow * H + j]) : 0.f;
            }
            cs[row * H + j] = c_new;
            hs[row * H + j] = from_f32<T>(h_new);
            if (s == T_ - 1) qf[(int64_t)b * dirs * H + dir * H + j] = from_f32<T>(c_new);
        }
    }
};

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
template <class AL, class BL, class EP>
__global__ void __launch_bounds__(NT, 2)
gemm_kernel(AL al, BL bl, EP ep, int M, int N, int K, int nsplit, int k_per_split) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int batch = blockIdx.z / nsplit, split = blockIdx.z - batch * nsplit;
    al.set_batch(batch); bl.set_batch(batch); ep.set_batch(batch);
    const int m_base = blockIdx.x * BM, n_base = blockIdx.y * BN;
    const int k_begin = split * k_per_split;
    const int k_end = min(K, k_begin + k_per_split);

    // loader mapping: thread -> one row, 8 consecutive k
    const int lrow = tid >> 1, lk = (tid & 1) * 8;
    typename AL::Row arow; typename BL::Row brow;
    al.init_row(arow, m_base + lrow);
    bl.init_row(brow, n_base + lrow);

    const int tx = tid & 15, ty = tid >> 4;
    float acc[2][2][4][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[a][b][i][j] = 0.f;

    float ra[8], rb[8];
    const int ntiles = (k_end > k_begin) ? (k_end - k_begin + BK - 1) / BK : 0;

    auto fetch = [&](int t) {
        const int k0 = k_begin + t * BK + lk;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + i;
            ra[i] = k < k_end ? al.at(arow, k) : 0.f;
            rb[i] = k < k_end ? bl.at(brow, k) : 0.f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            As[buf][lk + i][lrow] = ra[i];
            Bs[buf][lk + i][lrow] = rb[i];
        }
    };

    if (ntiles > 0) { fetch(0); stash(0); }
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) fetch(t + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float av[2][4] = {{a0.x, a0.y, a0.z, a0.w}, {a1.x, a1.y, a1.z, a1.w}};
            const float bv[2][4] = {{b0.x, b0.y, b0.z, b0.w}, {b1.x, b1.y, b1.z, b1.w}};
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[a][b][i][j] = fmaf(av[a][i], bv[b][j], acc[a][b][i][j]);
        }
        if (t + 1 < ntiles) stash(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
            ep.apply(m_base + a * 64 + ty * 4, n_base + b * 64 + tx * 4, acc[a][b], M, N);
}

template <class AL, class BL, class EP>
inline int launch(const AL& al, const BL& bl, const EP& ep, int M, int N, int K, int nbatch, int nsplit,
                  cudaStream_t st, const char* name) {
    if (M <= 0 || N <= 0) return 0;
    if (nsplit < 1) nsplit = 1;
    int kps = ((K + nsplit - 1) / nsplit + BK - 1) / BK * BK;
    if (kps < BK) kps = BK;
    nsplit = K > 0 ? (K + kps - 1) / kps : 1;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, nbatch * nsplit);
    gemm_kernel<AL, BL, EP><<<grid, NT, 0, st>>>(al, bl, ep, M, N, K, nsplit, kps);
    VQA_CHECK_LAUNCH(name);
    return 0;
}

// choose a split-K factor so that the grid has about 4 CTAs per SM
inline int pick_split(int M, int N, int K, int nbatch) {
    const int64_t tiles = (int64_t)((M + BM - 1) / BM) * ((N + BN - 1) / BN) * nbatch;
    int64_t want = (148 * 4 + tiles - 1) / tiles;
    const int64_t maxs = (K + 4 * BK - 1) / (4 * BK);
    if (want > maxs) want = maxs;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace simt
