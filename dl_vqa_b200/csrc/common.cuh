// Shared device/host helpers for libvqa_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vqa_b200.h"

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns 0 or an error code; message kept per thread
// ------------------------------------------------------------------------------------------
void vqa_set_error(const char* fmt, ...);
void vqa_count_launch();

#define VQA_REQUIRE(cond, ...)                         \
    do {                                               \
        if (!(cond)) {                                 \
            vqa_set_error(__VA_ARGS__);                \
            return VQA_ERR_INVALID_ARGUMENT;           \
        }                                              \
    } while (0)

#define VQA_CHECK_LAUNCH(name)                                                        \
    do {                                                                              \
        vqa_count_launch();                                                           \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            vqa_set_error("%s: %s", name, cudaGetErrorString(e__));                   \
            return (int)e__;                                                          \
        }                                                                             \
    } while (0)

// ---- programmatic dependent launch (PDL).  A kernel launched through vqa_launch_pdl may become resident while its
// predecessor in the stream is still running: it runs its prologue (barrier init, TMEM allocation, descriptor
// prefetch), then pdl_wait() blocks until the predecessor has completed and its writes are visible.  Rules for a
// kernel launched this way: pdl_trigger() first (lets ITS successor do the same), and NO global-memory access of any
// kind before pdl_wait().  Hides the ~2.5 us dependent-launch gap between the ~110 kernels of a step.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int vqa_pdl_enabled();          // env VQA_PDL=0 turns the attribute off (kernels then serialise as usual)

template <typename... KArgs, typename... Args>
static inline cudaError_t vqa_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = vqa_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

#define VQA_CUDA(call)                                                                \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) {                                                     \
            vqa_set_error("%s: %s", #call, cudaGetErrorString(e__));                  \
            return (int)e__;                                                          \
        }                                                                             \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------
// type helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float x) { return x; }
__device__ __forceinline__ float to_f32(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sign bits of the four bytes of z replicated over two 16-bit lanes each: LO -> bytes 0,1; HI -> bytes 2,3
// (prmt's sign-replicate selector mode, which __byte_perm does not expose).
__device__ __forceinline__ uint32_t sign_mask16_lo(uint32_t z) { uint32_t m; asm("prmt.b32 %0, %1, %1, 0x9988;" : "=r"(m) : "r"(z)); return m; }
__device__ __forceinline__ uint32_t sign_mask16_hi(uint32_t z) { uint32_t m; asm("prmt.b32 %0, %1, %1, 0xBBAA;" : "=r"(m) : "r"(z)); return m; }

// Column sums of a 32 x 32 tile held one ROW per lane (v[j] = element j of this lane's row): afterwards v[0] of lane l is
// the sum over all 32 lanes of element l.  Transposing butterfly: 31 shuffles instead of 32 x 5.
__device__ __forceinline__ void warp_colsum32(float (&v)[32], int lane) {
    int off = 16;
#pragma unroll
    for (int n = 32; n > 1; n >>= 1, off >>= 1) {
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float keep = up ? v[n / 2 + i] : v[i];
            const float send = up ? v[i] : v[n / 2 + i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// block-wide sum; `red` is >= 32 floats of shared memory; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// fire-and-forget fp32 atomic add of four consecutive elements (REDG.E.ADD.F32x4, sm_90+); p must be 16-byte aligned
__device__ __forceinline__ void red_add_f32x4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// hardware tanh (bf16 tensor-core arm only; the fp32 arm keeps the precise forms)
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }

// ---- 8-wide bf16 / fp32 register <-> memory helpers (128-bit accesses)
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]);
// question length as the kernels use it: the reference's pack_padded_sequence raises for lengths outside 1..T; a device
// kernel cannot, so the value is clamped to [0, T] (0 = no active step) and can never index outside the padded question
__device__ __forceinline__ int clamp_len(int64_t len, int T_) { return len < 0 ? 0 : (len > T_ ? T_ : (int)len); }

// 8 values from `a` (bf16) when it is non-null, else from `b` (fp32)
__device__ __forceinline__ void ld8(const bf16* a, const float* b, float (&v)[8]) {
    if (a) ld8(a, v); else ld8(b, v);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {       // 8 packed bf16 -> fp32
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void st8(bf16* p, const float (&v)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ void st8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *(reinterpret_cast<float4*>(p) + 1) = make_float4(v[4], v[5], v[6], v[7]);
}

// One item of the LSTM backward pointwise step (bf16 arm): 8 hidden units of sample b, direction dir, step s.
// Gate activations `gates` [dirs][T][B][4H] (i,f,g,o as stored by the forward), cell states `cs` [dirs][T][B][H];
// reads and CLEARS dh (the data-gradient GEMM of the next step accumulates into it), updates dc in place, writes the
// gate gradients dg [dirs][T][B][4H].  Shared by lstm_bwd_pointwise_vec8_kernel and the persistent backward kernel.
// DH_L2: read dh with ld.global.cg (L2 only) -- needed when OTHER SMs accumulate into dh between two calls inside one kernel.
template <bool DH_L2>
__device__ __forceinline__ void lstm_bwd_pointwise_item8(int64_t i8, const bf16* __restrict__ gates, const float* __restrict__ cs,
                                                         float* dh, float* __restrict__ dc,
                                                         const bf16* __restrict__ dc_init, bf16* __restrict__ dg,
                                                         const int64_t* __restrict__ q_len, int s, int T_, int B, int H, int dirs,
                                                         const int* __restrict__ order = nullptr) {
    // `order` (length-sorted batches, vqa_length_order): row b of every LSTM tensor is sample order[b]; only dc_init
    // [B][dirs*H] is in the caller's sample order
    const int h8 = H >> 3;
    const int j = (int)(i8 % h8) * 8;
    const int b = (int)((i8 / h8) % B);
    const int dir = (int)(i8 / ((int64_t)B * h8));
    const int64_t i = ((int64_t)dir * B + b) * H + j;
    const int64_t row = ((int64_t)dir * T_ + s) * B + b;
    bf16* o = dg + row * 4 * H + j;
    // every load is issued before the first dependent branch (this kernel is one L2 round trip long: a load that waits
    // for q_len, or sits in an if/else diamond, doubles it)
    const int len = clamp_len(q_len[b], T_);
    const bf16* g = gates + row * 4 * H + j;
    float dc_in[8], gi[8], gf[8], gg[8], go[8], c[8], cp[8], dhv[8];
    ld8(dc_init ? dc_init + (int64_t)(order ? order[b] : b) * dirs * H + (int64_t)dir * H + j : nullptr, dc + i, dc_in);
    ld8(g, gi); ld8(g + H, gf); ld8(g + 2 * H, gg); ld8(g + 3 * H, go);
    ld8(cs + row * H + j, c);
    ld8(cs + (s > 0 ? row - B : row) * H + j, cp);
    if (DH_L2) {
        const float4 a = __ldcg(reinterpret_cast<const float4*>(dh + i)), b4 = __ldcg(reinterpret_cast<const float4*>(dh + i) + 1);
        dhv[0] = a.x; dhv[1] = a.y; dhv[2] = a.z; dhv[3] = a.w; dhv[4] = b4.x; dhv[5] = b4.y; dhv[6] = b4.z; dhv[7] = b4.w;
    } else {
        ld8(dh + i, dhv);
    }
    if (s >= len) {                               // frozen step: zero gate gradients, dc passes through unchanged
        const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) *reinterpret_cast<uint4*>(o + gq * H) = z;
        st8(dc + i, dc_in);
        return;
    }
    if (s == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cp[k] = 0.f;
    }
    {   // read-and-clear: the next step's data gradient ACCUMULATES into dh (split-K with vector reductions)
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        st8(dh + i, z);
    }
    float di[8], df[8], dgg[8], dox[8], dcn[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float tc = tanh_approx(c[k]);            // the same hardware tanh as the persistent forward kernel
        const float dcv = dc_in[k] + dhv[k] * go[k] * (1.f - tc * tc);
        dcn[k] = dcv * gf[k];
        di[k] = dcv * gg[k] * gi[k] * (1.f - gi[k]);
        df[k] = dcv * cp[k] * gf[k] * (1.f - gf[k]);
        dgg[k] = dcv * gi[k] * (1.f - gg[k] * gg[k]);
        dox[k] = dhv[k] * tc * go[k] * (1.f - go[k]);
    }
    st8(dc + i, dcn);
    st8(o, di); st8(o + H, df); st8(o + 2 * H, dgg); st8(o + 3 * H, dox);
}

// ------------------------------------------------------------------------------------------
// counter-based dropout.  Element i of dropout site `site` is kept iff a 16-bit word derived from
// hash32(key(seed, site) ^ (i >> 1)) is >= threshold (p quantised to 1/65536).  Stateless, so the
// backward pass regenerates the same mask and no mask tensor ever touches HBM.  One hash (2 IMUL +
// 3 shift-xor) serves two elements, which keeps the fused attention kernel HBM-bound in train mode
// (a Philox4x32-10 stream would make it ALU-bound).  The mask cannot match torch's RNG stream bit for
// bit, so parity runs use eval() / dropout 0 (SURVEY.md section 7 "Dropout parity").
// ------------------------------------------------------------------------------------------
struct Dropout {
    uint64_t seed;
    const uint64_t* seed_ptr;   // non-null: the seed is read from device memory when the kernel RUNS (VQA_SEED_ON_DEVICE)
    uint32_t threshold;   // 16-bit threshold: keep iff word >= threshold; 0 disables dropout
    float scale;          // 1/(1-p)
    uint32_t thr2;        // vector scheme (Dropout8): the threshold as a packed bf16x2 bit pattern; 0 disables dropout
};

// Vector scheme: a 16-bit random field h is read as a bf16 BIT PATTERN and the element is dropped iff h < T (ordered
// compare), so one `set.geu.u32.bf16x2` turns a 32-bit random word into the keep masks (0xFFFF / 0) of two elements.
// Number of the 65536 patterns below a threshold T (inf = 0x7F80; NaN patterns compare unordered = kept):
//   T = -t (t >= 1): 0x7F80 - t      T = -0: 0x7F80      T = +t: 0x7F81 + t
// so every drop count up to 0xFF01 (p <= 0.996) is reachable except 0x7F81.  Thresholds are kept NORMAL numbers (or
// zero): whether the comparison flushes subnormal inputs then does not matter; the cost is |p' - p| <= 2^-9 inside
// 0.496 < p < 0.5002 and <= 2^-16 elsewhere (p = 0.5 itself lands on 32769 / 65536).
static inline uint32_t dropout_bf16_threshold(uint32_t count) {
    if (count == 0) return 0;
    uint32_t pat;
    if (count <= 0x7F80u) {
        const uint32_t t = 0x7F80u - count;                       // T = -t
        pat = t == 0 ? 0x8000u : (t < 0x80u ? (t < 0x40u ? 0x8000u : 0x8080u) : (0x8000u | t));
    } else {
        uint32_t t = count - 0x7F81u;                             // T = +t
        if (t > 0x7F80u) t = 0x7F80u;
        pat = t < 0x80u ? (t < 0x40u ? 0x8000u : 0x0080u) : t;
    }
    return pat | (pat << 16);
}

static inline Dropout make_dropout(uint64_t seed, float p) {
    Dropout d;
    d.seed = seed;
    d.seed_ptr = (seed & VQA_SEED_ON_DEVICE) ? reinterpret_cast<const uint64_t*>(seed & ~VQA_SEED_ON_DEVICE) : nullptr;
    if (p <= 0.f) { d.threshold = 0; d.scale = 1.f; d.thr2 = 0; }
    else {
        double t = (double)p * 65536.0 + 0.5;
        d.threshold = (t >= 65535.0) ? 65535u : (uint32_t)t;
        d.scale = 1.f / (1.f - p);
        d.thr2 = dropout_bf16_threshold(d.threshold);
    }
    return d;
}

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
    return x;
}
__device__ __forceinline__ uint32_t dropout_key(const Dropout& d, uint32_t site) {
    const uint64_t seed = d.seed_ptr ? __ldg(d.seed_ptr) : d.seed;      // after pdl_wait(): every caller is past it
    return hash32((uint32_t)seed ^ hash32((uint32_t)(seed >> 32) + site * 0x9E3779B9u + 0x85ebca6bu));
}
// 32 random bits covering elements 2*pair and 2*pair+1
__device__ __forceinline__ uint32_t dropout_word(uint32_t key, uint64_t pair) {
    return hash32(((uint32_t)pair ^ key) + (uint32_t)(pair >> 32) * 0xc2b2ae35u);
}
// multiplier (0 or scale) for one element
__device__ __forceinline__ float dropout_mult(const Dropout& d, uint32_t site, uint64_t idx) {
    if (d.threshold == 0) return 1.f;
    const uint32_t w = dropout_word(dropout_key(d, site), idx >> 1);
    const uint32_t h = (idx & 1) ? (w >> 16) : (w & 0xffffu);
    return h >= d.threshold ? d.scale : 0.f;
}
// multipliers for 2 consecutive elements starting at even idx, with a precomputed key
__device__ __forceinline__ void dropout_mult2(const Dropout& d, uint32_t key, uint64_t idx, float& m0, float& m1) {
    const uint32_t w = dropout_word(key, idx >> 1);
    m0 = (w & 0xffffu) >= d.threshold ? d.scale : 0.f;
    m1 = (w >> 16) >= d.threshold ? d.scale : 0.f;
}

// ---- vector form used by the fused attention / dropout+L2-norm kernels (0.7 G elements per step at site ATT_X) ------
// One hash32 per group of 8 consecutive elements, three LCG steps for the other three words; each 32-bit word serves
// two elements: its 16-bit halves are compared as bf16 bit patterns against the threshold (see dropout_bf16_threshold),
// one HSET2 per word on the half-precision pipe instead of AND + ADD + PRMT on the integer pipe, which is the pipe the
// streaming attention kernels saturate in train mode.  Forward and backward regenerate the same masks.
struct Dropout8 {
    uint32_t key;       // dropout_key(seed, site)
    uint32_t thr2;      // packed bf16x2 threshold pattern; 0 disables dropout
    float scale;        // 1/(1-p)
};
__device__ __forceinline__ Dropout8 make_dropout8(const Dropout& d, uint32_t site) {
    Dropout8 r;
    r.key = dropout_key(d, site);
    r.thr2 = d.threshold == 0 ? 0u : d.thr2;
    r.scale = d.scale;
    return r;
}
// keep masks for elements 8*group .. 8*group+7: element 2j <-> low half of t[j], element 2j+1 <-> high half (0xFFFF = kept)
__device__ __forceinline__ void dropout_flags8(const Dropout8& d, uint32_t group, uint32_t (&t)[4]) {
    uint32_t w = hash32(group ^ d.key);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        asm("set.geu.u32.bf16x2 %0, %1, %2;" : "=r"(t[j]) : "r"(w), "r"(d.thr2));
        w = w * 0x9E3779B1u + 0x7F4A7C15u;
    }
}
// the flags ARE the bf16x2 bit masks
__device__ __forceinline__ uint32_t dropout_mask_bf16x2(uint32_t t) { return t; }
__device__ __forceinline__ void dropout_mult8(const Dropout8& d, uint32_t group, float (&m)[8]) {
    if (d.thr2 == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = 1.f;
        return;
    }
    uint32_t t[4];
    dropout_flags8(d, group, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        m[2 * j] = (t[j] & 0x8000u) ? d.scale : 0.f;
        m[2 * j + 1] = (t[j] & 0x80000000u) ? d.scale : 0.f;
    }
}

// dropout sites (one independent mask stream per nn.Dropout call site of models/model.py)
enum : uint32_t {
    SITE_IMAGE = 0,      // models/model.py:84   image.drop
    SITE_ATT_V = 1,      // models/model.py:185  attention.drop(v)
    SITE_EMBED = 2,      // models/model.py:156  text.drop
    SITE_ATT_Q = 3,      // models/model.py:186  attention.drop(q)
    SITE_ATT_X = 4,      // models/model.py:194  attention.drop(x)
    SITE_CLS_IN = 5,     // models/model.py:201  classifier.drop1
    SITE_CLS_HID = 6,    // models/model.py:204  classifier.drop2
};
