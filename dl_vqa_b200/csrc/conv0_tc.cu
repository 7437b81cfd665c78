// First image-encoder layer (3 input channels, models/model.py:80 with num_channels[0] = 3) on tcgen05.
//
// K = 3*3*3 = 27 is far too small for a TMA-fed pipeline to pay off and the NCHW fp32 network input is not a
// UMMA operand, so builder warps write the im2col tile straight into 128-byte-swizzled shared memory:
// one thread per output position gathers its 27 inputs (L1-cached, each input is reused 9x), converts to
// bf16 and stores 4 x 16 B at the swizzled chunk positions -> a [128 positions][k] tile with 128-byte rows.
//   forward : that tile is the K-major A operand (K = 32, two MMAs), weights [64 co][k] are built once per
//             CTA; epilogue = bias + ReLU + 2x2 max-pool + mask exactly as conv_tc.cu (the 64x222x222 un-pooled
//             activation, 1.6 GB per 256-sample batch, never exists).
//   wgrad   : the SAME bytes read as an MN-major B operand (rows = positions = reduction index), the gradient
//             tile dY[128 positions][64 co] arrives by TMA as the MN-major A operand:
//             dW[co][k] += sum_pos dY[pos][co] * patch[pos][k]  -- accumulated in TMEM over the CTA's slice of
//             positions, flushed once with atomics (64 x 27 outputs).
// Both are bandwidth/epilogue-bound kernels; the tensor core only removes the 27x64 FMAs per position that a
// SIMT formulation spends.
#include "tc_common.cuh"

namespace tc {

constexpr int C0_THREADS = 288;        // warps 0-3 epilogue, 4-7 im2col builders, 8 MMA issuer
constexpr int C0_TILE_BYTES = 128 * 128;
constexpr int C0_STAGES = 3;
constexpr int C0_K = 27;

struct Conv0Params {
    const float* x;                    // [B,3,IH,IW] NCHW fp32
    int B, IH, IW, tiles_h, tiles_w;
    // forward
    const float* w; const float* bias; bf16* pooled; uint8_t* mask; int PH, PW;
    // wgrad
    float* dw; int chunks_per_cta;
};

// gather the 27 inputs of output position (h,w) of image b (clamped so that out-of-range tile positions read
// valid memory; their results are discarded) and write one swizzled 64-byte row segment of the tile
__device__ __forceinline__ void build_im2col_row(const Conv0Params& p, uint8_t* tile, int row, int b, int h, int w) {
    const int hc = min(h, p.IH - 3), wc = min(w, p.IW - 3);
    const float* src = p.x + ((int64_t)b * 3 * p.IH + hc) * p.IW + wc;
    float v[32];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
                v[ci * 9 + kh * 3 + kw] = __ldg(src + ((int64_t)ci * p.IH + kh) * p.IW + kw);
#pragma unroll
    for (int k = C0_K; k < 32; ++k) v[k] = 0.f;
    uint8_t* rowp = tile + row * 128;
    const int sw = row & 7;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint4 u;
        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int t = 0; t < 4; ++t) hh[t] = __floats2bfloat162_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
        *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = u;
    }
}

// ------------------------------------------------------------------------------------------ forward
__global__ void __launch_bounds__(C0_THREADS, 2) conv0_fwd_tc_kernel(Conv0Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sw_tile = smem;                                   // [64 co][128 B]
    uint8_t* sa = smem + 64 * 128;                             // C0_STAGES x [128 pos][128 B]
    uint64_t* a_full = reinterpret_cast<uint64_t*>(sa + C0_STAGES * C0_TILE_BYTES);
    uint64_t* a_empty = a_full + C0_STAGES;
    uint64_t* tmem_full = a_empty + C0_STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    constexpr int BN = 64;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_h * p.tiles_w;
    const int ntiles = p.B * tiles_per_img;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C0_STAGES; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        fence_barrier_init();
    }
    if (threadIdx.x < 64) {                                    // weight tile: row = co, 32 k-values (27 valid)
        const int co = threadIdx.x;
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = k < C0_K ? p.w[co * C0_K + k] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (j < 4) {
                __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int t = 0; t < 4; ++t) hh[t] = __floats2bfloat162_rn(v[8 * j + 2 * t], v[8 * j + 2 * t + 1]);
            }
            *reinterpret_cast<uint4*>(sw_tile + co * 128 + ((j ^ (co & 7)) << 4)) = u;
        }
        fence_proxy_async();
    }
    if (warp == 8) tmem_alloc(tmem_base_smem, 2 * BN);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp >= 4 && warp < 8) {
        // ---- im2col builders: thread -> tile position
        const int t = threadIdx.x - 128;
        const int rr = t >> 4, cc = t & 15;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
            const int h0 = (r / p.tiles_w) * 8, w0 = (r % p.tiles_w) * 16;
            const int s = it % C0_STAGES;
            const uint32_t ph = (it / C0_STAGES) & 1;
            mbar_wait(&a_empty[s], ph ^ 1);
            build_im2col_row(p, sa + s * C0_TILE_BYTES, t, b, h0 + rr, w0 + cc);
            fence_proxy_async();
            mbar_arrive(&a_full[s]);
        }
    } else if (warp == 8) {
        // all lanes converged, one elected lane issues (see tc_common.cuh)
        constexpr uint32_t idesc = idesc_bf16(128, BN);
        const uint32_t elected = elect_one();
        const uint64_t a_desc0 = smem_desc_k_sw128(smem_u32(sa)), w_desc = smem_desc_k_sw128(smem_u32(sw_tile));
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, use = it >> 1;
            const int s = it % C0_STAGES;
            mbar_wait(&tmem_empty[acc], (use & 1) ^ 1);
            mbar_wait(&a_full[s], (it / C0_STAGES) & 1);
            tcgen05_fence_after();
            const uint64_t ad = a_desc0 + (uint64_t)(s * (C0_TILE_BYTES >> 4));
#pragma unroll
            for (uint32_t k = 0; k < 2; ++k)
                umma_issue<1>(tmem_base + acc * BN, ad + 2 * k, w_desc + 2 * k, idesc, k, elected);
            umma_commit_issue<1>(&a_empty[s], elected);
            umma_commit_issue<1>(&tmem_full[acc], elected);
        }
    } else if (warp < 4) {
        // ---- epilogue: bias + ReLU + 2x2 max-pool + mask (same butterfly as conv_tc.cu)
        const int quarter = warp;
        const int cc = lane & 15;
        const int bit0 = lane & 1, bit4 = (lane >> 4) & 1;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, use = it >> 1;
            const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
            const int h0 = (r / p.tiles_w) * 8, w0 = (r % p.tiles_w) * 16;
            mbar_wait(&tmem_full[acc], use & 1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            const int ph_ = (h0 >> 1) + quarter, pw_ = (w0 >> 1) + (cc >> 1);
            const bool ok = ph_ < p.PH && pw_ < p.PW;
            const int64_t obase = (((int64_t)b * p.PH + ph_) * p.PW + pw_) * BN;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                float v[32];
                tmem_ld_32x32(taddr + c0, v);
                float k1[16]; int i1[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float mine = bit0 ? v[16 + j] : v[j];
                    const float send = bit0 ? v[j] : v[16 + j];
                    const float other = __shfl_xor_sync(0xffffffffu, send, 1);
                    const bool take_other = bit0 ? (other >= mine) : (other > mine);
                    k1[j] = take_other ? other : mine;
                    i1[j] = take_other ? (bit0 ^ 1) : bit0;
                }
                float k2[8]; int i2[8];
                uint32_t pack_send = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) pack_send |= (uint32_t)(bit4 ? i1[j] : i1[8 + j]) << (2 * j);
                const uint32_t pack_other = __shfl_xor_sync(0xffffffffu, pack_send, 16);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float mine = bit4 ? k1[8 + j] : k1[j];
                    const int mine_i = (bit4 ? i1[8 + j] : i1[j]) + 2 * bit4;
                    const float send = bit4 ? k1[j] : k1[8 + j];
                    const float other = __shfl_xor_sync(0xffffffffu, send, 16);
                    const int other_i = (int)((pack_other >> (2 * j)) & 3u) + 2 * (bit4 ^ 1);
                    const bool take_other = bit4 ? (other >= mine) : (other > mine);
                    k2[j] = take_other ? other : mine;
                    i2[j] = take_other ? other_i : mine_i;
                }
                if (ok) {
                    const int nb = c0 + 16 * bit0 + 8 * bit4;
                    const float4 b0 = *reinterpret_cast<const float4*>(p.bias + nb);
                    const float4 b1 = *reinterpret_cast<const float4*>(p.bias + nb + 4);
                    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                    uint4 u; uint2 mk;
                    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
                    uint8_t* mb = reinterpret_cast<uint8_t*>(&mk);
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float x = k2[j] + bb[j];
                        int id = i2[j];
                        if (!(x > 0.f)) { x = 0.f; id = 4; }
                        o[j] = x; mb[j] = (uint8_t)id;
                    }
#pragma unroll
                    for (int t = 0; t < 4; ++t) hh[t] = __floats2bfloat162_rn(o[2 * t], o[2 * t + 1]);
                    *reinterpret_cast<uint4*>(p.pooled + obase + nb) = u;
                    *reinterpret_cast<uint2*>(p.mask + obase + nb) = mk;
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 2 * BN); }
}

// ------------------------------------------------------------------------------------------ weight gradient
// thread roles: warps 0-3 final epilogue, warps 4-7 im2col builders (B operand), warp 8 MMA, warp 9 TMA (A operand)
constexpr int C0W_THREADS = 320;

__global__ void __launch_bounds__(C0W_THREADS, 1)
conv0_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tma_dy, Conv0Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // stage = [A: dY tile 128 pos x 64 co][B: im2col tile 128 pos x 64 k]
    constexpr int STAGE = 2 * C0_TILE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + C0_STAGES * STAGE);
    uint64_t* empty = full + C0_STAGES;
    uint64_t* tmem_full = empty + C0_STAGES;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);
    constexpr uint32_t TMEM_COLS = 32;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_h * p.tiles_w;
    const int total = p.B * tiles_per_img;
    const int c_begin = blockIdx.x * p.chunks_per_cta;
    const int c_end = min(total, c_begin + p.chunks_per_cta);
    const int nch = max(0, c_end - c_begin);

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_dy);
        for (int i = 0; i < C0_STAGES; ++i) { mbar_init(&full[i], 129); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 8) tmem_alloc(tmem_base_smem, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 9) {
        if (lane == 0) {
            for (int i = 0; i < nch; ++i) {
                const int c = c_begin + i;
                const int b = c / tiles_per_img, r = c - b * tiles_per_img;
                const int h0 = (r / p.tiles_w) * 8, w0 = (r % p.tiles_w) * 16;
                const int s = i % C0_STAGES;
                const uint32_t ph = (i / C0_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], C0_TILE_BYTES);
                tma_load_4d(smem + s * STAGE, &tma_dy, &full[s], 0, w0, h0, b);
            }
        }
    } else if (warp >= 4 && warp < 8) {
        const int t = threadIdx.x - 128;
        const int rr = t >> 4, cc = t & 15;
        for (int i = 0; i < nch; ++i) {
            const int c = c_begin + i;
            const int b = c / tiles_per_img, r = c - b * tiles_per_img;
            const int h0 = (r / p.tiles_w) * 8, w0 = (r % p.tiles_w) * 16;
            const int s = i % C0_STAGES;
            const uint32_t ph = (i / C0_STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            build_im2col_row(p, smem + s * STAGE + C0_TILE_BYTES, t, b, h0 + rr, w0 + cc);
            fence_proxy_async();
            mbar_arrive(&full[s]);
        }
    } else if (warp == 8) {
        // D[128 (co; rows 64..127 unused)][32 k] ; A = dY tile (MN-major, 64 channels = one block; the second
        // block address is arbitrary valid shared memory: its rows only feed the unused accumulator rows)
        constexpr uint32_t idesc = idesc_bf16(128, 32, 1, 1);
        const uint32_t elected = elect_one();
        const uint64_t a_desc0 = smem_desc_mn_sw128(smem_u32(smem), 1024);
        for (int i = 0; i < nch; ++i) {
            const int s = i % C0_STAGES;
            mbar_wait(&full[s], (i / C0_STAGES) & 1);
            tcgen05_fence_after();
            const uint64_t ad = a_desc0 + (uint64_t)(s * (STAGE >> 4)), bd = ad + (uint64_t)(C0_TILE_BYTES >> 4);
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k)
                umma_issue<1>(tmem_base, ad + k * (2048 >> 4), bd + k * (2048 >> 4), idesc, (i > 0 || k > 0) ? 1u : 0u, elected);
            umma_commit_issue<1>(&empty[s], elected);
        }
        umma_commit_issue<1>(tmem_full, elected);
    } else if (warp < 4 && nch > 0) {
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
        const int co = warp * 32 + lane;
        if (co < 64) {
#pragma unroll
            for (int k = 0; k < C0_K; ++k) atomicAdd(p.dw + co * C0_K + k, v[k]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) { tcgen05_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

}  // namespace tc

using namespace tc;

static int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

// x [B,3,IH,IW] fp32 NCHW; w [64,3,3,3] fp32; out/mask [B,PH,PW,64] (bf16 / uint8)
extern "C" int vqa_tc_conv0_relu_pool_fwd(const float* x, const float* w, const float* bias, void* out, uint8_t* mask,
                                          int B, int IH, int IW, int Cin, int Cout, void* stream) {
    VQA_REQUIRE(Cin == 3 && Cout == 64, "tc conv0 fwd: only Cin=3, Cout=64 (got %d, %d); use vqa_conv_relu_pool_fwd", Cin, Cout);
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4, "tc conv0 fwd: bad dims");
    Conv0Params p{};
    p.x = x; p.B = B; p.IH = IH; p.IW = IW;
    p.PH = (IH - 2) / 2; p.PW = (IW - 2) / 2;
    p.tiles_h = (2 * p.PH + 7) / 8; p.tiles_w = (2 * p.PW + 15) / 16;
    p.w = w; p.bias = bias; p.pooled = (bf16*)out; p.mask = mask;
    const int smem = 64 * 128 + C0_STAGES * C0_TILE_BYTES + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) { VQA_CUDA(cudaFuncSetAttribute(conv0_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
    const int ntiles = B * p.tiles_h * p.tiles_w;
    const int sms = sm_count();
    conv0_fwd_tc_kernel<<<ntiles < 2 * sms ? ntiles : 2 * sms, C0_THREADS, smem, (cudaStream_t)stream>>>(p);
    VQA_CHECK_LAUNCH("conv0_fwd_tc");
    return 0;
}

// x [B,3,IH,IW] fp32 NCHW; dy [B,2PH,2PW,64] bf16 (vqa_unpool_bf16); dw [64,3,3,3] fp32 (overwritten)
extern "C" int vqa_tc_conv0_bwd_weight(const float* x, const void* dy, float* dw, int B, int IH, int IW, int Cin, int Cout,
                                       void* stream) {
    VQA_REQUIRE(Cin == 3 && Cout == 64, "tc conv0 wgrad: only Cin=3, Cout=64 (got %d, %d)", Cin, Cout);
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4, "tc conv0 wgrad: bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    VQA_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 64 * C0_K, st));
    Conv0Params p{};
    p.x = x; p.B = B; p.IH = IH; p.IW = IW;
    p.PH = (IH - 2) / 2; p.PW = (IW - 2) / 2;
    p.tiles_h = (2 * p.PH + 7) / 8; p.tiles_w = (2 * p.PW + 15) / 16;
    p.dw = dw;
    CUtensorMap tdy;
    {
        const uint64_t dims[4] = {64, (uint64_t)(2 * p.PW), (uint64_t)(2 * p.PH), (uint64_t)B};
        const uint64_t str[3] = {64 * 2, (uint64_t)(2 * p.PW) * 64 * 2, (uint64_t)(2 * p.PH) * (2 * p.PW) * 64 * 2};
        const uint32_t box[4] = {64, 16, 8, 1};
        if (int e = make_tmap_bf16(&tdy, dy, 4, dims, str, box)) return e;
    }
    const int total = B * p.tiles_h * p.tiles_w;
    const int sms = sm_count();
    int ctas = total < sms ? total : sms;
    p.chunks_per_cta = (total + ctas - 1) / ctas;
    ctas = (total + p.chunks_per_cta - 1) / p.chunks_per_cta;
    const int smem = C0_STAGES * 2 * C0_TILE_BYTES + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) { VQA_CUDA(cudaFuncSetAttribute(conv0_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
    conv0_wgrad_tc_kernel<<<ctas, C0W_THREADS, smem, st>>>(tdy, p);
    VQA_CHECK_LAUNCH("conv0_wgrad_tc");
    return 0;
}
