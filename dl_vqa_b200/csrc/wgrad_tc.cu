// Weight gradient of the 3x3 convolutions on tcgen05 (autograd of models/model.py:80):
//   dW[co][ci][kh][kw] = sum_{b,oh,ow} dY[b,oh,ow,co] * X[b,oh+kh,ow+kw,ci]
// The reduction index is the spatial position.  With NHWC activations a TMA box [64 ch, w, 4 h] lands in
// shared memory as position-rows x 128 bytes (128B swizzle) = the canonical MN-MAJOR UMMA operand
// (K = positions run across rows, 64 channels contiguous inside a row), so both operands are consumed
// straight from the NHWC tensors -- no im2col, no transposed copies:
//   A = dY[b, h0:h0+4, w0:w0+16, co0:co0+128]            two 64-channel blocks (LBO = 8 KB apart)
//   B = X [b, h0+kh:+4, w0:w0+18, 64-channel block]      ONE box with a 2-column halo serves the three kw taps:
//       tap kw is the same tile one position-row (128 B) further, so the three taps are three "N blocks" of one
//       MMA with LBO = 128 B (the swizzle is a function of the absolute address, tools/umma_probe.cu).
// One CTA owns (128 output channels, one filter row kh) and a slice of the position tiles (split-K); per 16
// positions it issues one M=128 x N=192 MMA per 64-channel block of X (A is read once for three taps).  The
// accumulators (192 TMEM columns per block) are flushed once with fp32 atomics.
#include "tc_common.cuh"

namespace tc {

constexpr int WG_THREADS = 192;
constexpr int WG_TH = 4, WG_TW = 16;               // 64 positions per k-chunk
constexpr int BLK_BYTES = 64 * 128;                // one 64-position x 64-channel block of dY
constexpr int XBLK_BYTES = WG_TH * (WG_TW + 2) * 128;   // 72 position-rows (with halo) x 64 channels of X: 9216 B

template <int CIN>
struct WgSmem {
    static constexpr int NB = CIN / 64;            // channel blocks of the B operand
    static constexpr int A_BYTES = 2 * BLK_BYTES, B_BYTES = NB * XBLK_BYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (232448 - 2048) / STAGE_BYTES > 8 ? 8 : (232448 - 2048) / STAGE_BYTES;
    static constexpr int BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

struct WgParams {
    int B, tiles_h, tiles_w;  // position tiles per image over the un-pooled gradient region
    int chunks_per_split;
    int Cin, Cout;
    float* dw;                // fp32 OIHW, pre-zeroed
};

template <int CIN>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tma_dy, const __grid_constant__ CUtensorMap tma_x, WgParams p) {
    pdl_trigger();
    using S = WgSmem<CIN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::STAGES * S::STAGE_BYTES);
    uint64_t* empty = full + S::STAGES;
    uint64_t* tmem_full = empty + S::STAGES;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);
    constexpr uint32_t TMEM_COLS = S::NB * 192 <= 256 ? 256 : 512;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kh = blockIdx.x % 3, co0 = (blockIdx.x / 3) * 128;
    const int tiles_per_img = p.tiles_h * p.tiles_w;
    const int total = p.B * tiles_per_img;
    const int c_begin = blockIdx.y * p.chunks_per_split;
    const int c_end = min(total, c_begin + p.chunks_per_split);
    const int nch = max(0, c_end - c_begin);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_dy); tma_prefetch_desc(&tma_x);
        for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_smem, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nch; ++i) {
                const int c = c_begin + i;
                const int b = c / tiles_per_img, r = c - b * tiles_per_img;
                const int h0 = (r / p.tiles_w) * WG_TH, w0 = (r % p.tiles_w) * WG_TW;
                const int s = i % S::STAGES;
                const uint32_t ph = (i / S::STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], S::STAGE_BYTES);
                uint8_t* st = smem + s * S::STAGE_BYTES;
                tma_load_4d(st, &tma_dy, &full[s], co0, w0, h0, b);
                tma_load_4d(st + BLK_BYTES, &tma_dy, &full[s], co0 + 64, w0, h0, b);
#pragma unroll
                for (int nb = 0; nb < S::NB; ++nb)
                    tma_load_4d(st + S::A_BYTES + nb * XBLK_BYTES, &tma_x, &full[s], nb * 64, w0, h0 + kh, b);
            }
        }
    } else if (warp == 1) {
        // all lanes converged, one elected lane issues (see tc_common.cuh)
        constexpr uint32_t idesc = idesc_bf16(128, 192, /*a_mn_major=*/1, /*b_mn_major=*/1);
        const uint32_t elected = elect_one();
        const uint64_t a_desc0 = smem_desc_mn_sw128(smem_u32(smem), BLK_BYTES);
        const uint64_t b_desc0 = smem_desc_mn_sw128(smem_u32(smem + S::A_BYTES), 128);     // LBO = one position = one kw tap
        for (int i = 0; i < nch; ++i) {
            const int s = i % S::STAGES;
            mbar_wait(&full[s], (i / S::STAGES) & 1);
            tcgen05_fence_after();
            const uint64_t so = (uint64_t)(s * (S::STAGE_BYTES >> 4));
#pragma unroll
            for (int h = 0; h < WG_TH; ++h) {            // 16 positions (one tile row) per MMA
#pragma unroll
                for (int nb = 0; nb < S::NB; ++nb)
                    umma_issue<1>(tmem_base + nb * 192, a_desc0 + so + (uint64_t)(h * (2048 >> 4)),
                                  b_desc0 + so + (uint64_t)((nb * XBLK_BYTES + h * (WG_TW + 2) * 128) >> 4), idesc,
                                  (i > 0 || h > 0) ? 1u : 0u, elected);
            }
            umma_commit_issue<1>(&empty[s], elected);
        }
        umma_commit_issue<1>(tmem_full, elected);
    } else if (nch > 0) {
        const int quarter = warp & 3;
        const int co = co0 + quarter * 32 + lane;
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
#pragma unroll 1
        for (int nb = 0; nb < S::NB; ++nb) {
#pragma unroll 1
            for (int kw = 0; kw < 3; ++kw) {
                const int tap = kh * 3 + kw;
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + nb * 192 + kw * 64 + c0, v);
                    if (co < p.Cout) {
                        float* o = p.dw + ((int64_t)co * p.Cin + nb * 64 + c0) * 9 + tap;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(o + j * 9, v[j]);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// NHWC tensor map with the position box [64 ch, 16 (+halo) w, 4 h, 1]
static int pos_tmap(CUtensorMap* m, const void* base, int B, int H, int W, int C, int halo) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, (uint32_t)(WG_TW + halo), WG_TH, 1};
    return make_tmap_bf16(m, base, 4, dims, str, box);
}

template <int CIN>
static int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, WgParams p, cudaStream_t st) {
    auto kern = wgrad_tc_kernel<CIN>;
    static bool attr_set = false;
    if (!attr_set) {
        VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<CIN>::BYTES));
        attr_set = true;
    }
    const int units = (p.Cout / 128) * 3;
    const int total = p.B * p.tiles_h * p.tiles_w;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int nsplit = sms / units;                      // one CTA per SM, a single wave
    if (nsplit < 1) nsplit = 1;
    if (nsplit > total) nsplit = total;
    p.chunks_per_split = (total + nsplit - 1) / nsplit;
    nsplit = (total + p.chunks_per_split - 1) / p.chunks_per_split;
    dim3 grid(units, nsplit);
    VQA_CUDA(vqa_launch_pdl(kern, grid, dim3(WG_THREADS), WgSmem<CIN>::BYTES, st, tdy, tx, p));
    VQA_CHECK_LAUNCH("wgrad_tc");
    return 0;
}

}  // namespace tc

using namespace tc;

// x [B,IH,IW,Cin] bf16 NHWC (the layer input); dy [B,2PH,2PW,Cout] bf16 NHWC (un-pooled gradient from
// vqa_unpool_bf16); dw fp32 OIHW (overwritten).
extern "C" int vqa_tc_conv3x3_bwd_weight(const void* x, const void* dy, float* dw,
                                         int B, int IH, int IW, int Cin, int Cout, void* stream) {
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4, "tc conv wgrad: bad dims");
    VQA_REQUIRE(Cin == 64 || Cin == 128, "tc conv wgrad: Cin=%d must be 64 or 128", Cin);
    VQA_REQUIRE(Cout % 128 == 0, "tc conv wgrad: Cout=%d must be a multiple of 128", Cout);
    const int OHp = ((IH - 2) / 2) * 2, OWp = ((IW - 2) / 2) * 2;
    cudaStream_t st = (cudaStream_t)stream;
    VQA_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * 9, st));
    CUtensorMap tdy, tx;
    if (int e = pos_tmap(&tdy, dy, B, OHp, OWp, Cout, 0)) return e;
    if (int e = pos_tmap(&tx, x, B, IH, IW, Cin, 2)) return e;
    WgParams p{};
    p.B = B; p.tiles_h = (OHp + WG_TH - 1) / WG_TH; p.tiles_w = (OWp + WG_TW - 1) / WG_TW;
    p.Cin = Cin; p.Cout = Cout; p.dw = dw;
    if (Cin == 64) return launch_wgrad<64>(tdy, tx, p, st);
    return launch_wgrad<128>(tdy, tx, p, st);
}

// ------------------------------------------------------------------------------------------
// layout kernels feeding the weight gradient
// ------------------------------------------------------------------------------------------
// x [B,H,W,C] bf16 NHWC -> xT [B,C,H,Wp] (zero padded columns W..Wp-1)
__global__ void nhwc_to_nchw_pad_kernel(const bf16* __restrict__ x, bf16* __restrict__ xT, int H, int W, int C, int Wp) {
    pdl_trigger();
    pdl_wait();
    __shared__ bf16 tile[32][34];
    const int bh = blockIdx.z;                       // b*H + h
    const int b = bh / H, h = bh - b * H;
    const int w0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {      // i = w offset, threadIdx.x = c offset
        const int w = w0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (w < W && c < C) ? x[((int64_t)bh * W + w) * C + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {      // i = c offset, threadIdx.x = w offset
        const int c = c0 + i, w = w0 + threadIdx.x;
        if (c < C && w < Wp) xT[(((int64_t)b * C + c) * H + h) * Wp + w] = tile[threadIdx.x][i];
    }
}

extern "C" int vqa_nhwc_to_nchw_pad_bf16(const void* x, void* xT, int B, int H, int W, int C, int Wp, void* stream) {
    VQA_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Wp >= W, "nhwc_to_nchw: bad dims");
    VQA_REQUIRE((int64_t)B * H <= 65535 * 32ll, "nhwc_to_nchw: too many rows");
    dim3 grid((Wp + 31) / 32, (C + 31) / 32, B * H), block(32, 8);
    VQA_REQUIRE(grid.z <= 65535u * 1024u, "nhwc_to_nchw: grid too large");
    VQA_CUDA(vqa_launch_pdl(nhwc_to_nchw_pad_kernel, grid, block, 0, (cudaStream_t)stream, (const bf16*)x, (bf16*)xT, H, W, C, Wp));
    VQA_CHECK_LAUNCH("nhwc_to_nchw_pad");
    return 0;
}

// (dpool, mask) [B,PH,PW,C] -> dyT [B,C,2PH,OWpp]: un-pooled gradient, channel-major, zero padded
__global__ void unpool_nchw_kernel(const bf16* __restrict__ dpool, const uint8_t* __restrict__ mask, bf16* __restrict__ dyT,
                                   int PH, int PW, int C, int OWpp) {
    pdl_trigger();
    pdl_wait();
    __shared__ bf16 g[32][34];
    __shared__ uint8_t m[32][36];
    const int bp = blockIdx.z;                       // b*PH + ph
    const int b = bp / PH, ph = bp - b * PH;
    const int pw0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {      // i = pw offset, x = c offset
        const int pw = pw0 + i, c = c0 + threadIdx.x;
        const bool ok = pw < PW && c < C;
        const int64_t idx = ((int64_t)bp * PW + pw) * C + c;
        g[i][threadIdx.x] = ok ? dpool[idx] : __float2bfloat16_rn(0.f);
        m[i][threadIdx.x] = ok ? mask[idx] : (uint8_t)4;
    }
    __syncthreads();
    // each thread writes 2 consecutive output columns (one pool window wide) for both rows
    for (int i = threadIdx.y; i < 32; i += 8) {      // i = c offset, x = pw offset
        const int c = c0 + i, pw = pw0 + threadIdx.x;
        if (c >= C) continue;
        const bf16 gv = g[threadIdx.x][i];
        const int mv = m[threadIdx.x][i];
        const bf16 z = __float2bfloat16_rn(0.f);
        const int ow = 2 * pw;
        if (ow + 1 < OWpp || ow < OWpp) {
            bf16* r0 = dyT + (((int64_t)b * C + c) * (2 * PH) + 2 * ph) * OWpp;
            bf16* r1 = r0 + OWpp;
            if (ow < OWpp) { r0[ow] = mv == 0 ? gv : z; r1[ow] = mv == 2 ? gv : z; }
            if (ow + 1 < OWpp) { r0[ow + 1] = mv == 1 ? gv : z; r1[ow + 1] = mv == 3 ? gv : z; }
        }
    }
}

extern "C" int vqa_unpool_nchw_bf16(const void* dpool, const uint8_t* mask, void* dyT, int B, int PH, int PW, int C,
                                    int OWpp, void* stream) {
    VQA_REQUIRE(B > 0 && PH > 0 && PW > 0 && C > 0 && OWpp >= 2 * PW, "unpool_nchw: bad dims");
    // columns [2PW, OWpp) are covered because the pw range is rounded up to OWpp/2
    dim3 grid(((OWpp + 1) / 2 + 31) / 32, (C + 31) / 32, B * PH), block(32, 8);
    VQA_CUDA(vqa_launch_pdl(unpool_nchw_kernel, grid, block, 0, (cudaStream_t)stream, (const bf16*)dpool, mask, (bf16*)dyT, PH, PW, C, OWpp));
    VQA_CHECK_LAUNCH("unpool_nchw");
    return 0;
}
