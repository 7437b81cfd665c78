// 3x3 / stride-1 convolutions of the image encoder on tcgen05 tensor cores (models/model.py:72-84).
//
// Implicit GEMM without im2col: one CTA tile = 8 x 16 spatial positions (M = 128) x all output channels
// (N = BN <= 256); for filter tap (kh,kw) and a 64-channel slice the A operand is ONE 4-D TMA box
// [64 ch, 16 w, 8 h, 1 image] of the NHWC activation shifted by (kh,kw) -- it lands in shared memory as 128
// rows x 128 bytes with the 128-byte swizzle, exactly the K-major UMMA operand.  Out-of-bounds rows/cols are
// zero-filled by TMA (no padding copies).  The B operand is the packed weight [N][tap][C] (K-major).
//
// Persistent kernel: grid = #SMs, tiles round-robin; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 =
// epilogue.  Two TMEM accumulators (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Epilogues:
//   POOL  (forward)  bias + ReLU + 2x2 max-pool + arg-max mask: window partners are lanes l, l^1, l^16, l^17 of
//                    one warp (tile rows are 16 wide); a 2-step butterfly leaves each lane 8 of every 32 channels.
//   STORE (dgrad)    plain bf16 store of the 128 x BN tile (gradient w.r.t. the layer input).
#include "tc_common.cuh"

namespace tc {

constexpr int CONV_THREADS = 192;
constexpr int TILE_H = 8, TILE_W = 16;
enum { EPI_POOL = 0, EPI_STORE = 1 };

struct ConvParams {
    int B, tiles_h, tiles_w;       // tile grid per image
    int chunks;                    // 64-channel slices of the reduction channel dimension
    int sign;                      // +1: forward taps (h+kh, w+kw);  -1: data-gradient taps (h-kh, w-kw)
    int valid_h, valid_w;          // extent of valid output positions (conv-output for POOL, input for STORE)
    int N;                         // output channels (== BN)
    // POOL
    const float* bias; bf16* pooled; uint8_t* mask; int PH, PW;
    // STORE
    bf16* dx;
};

template <int BN>
struct ConvSmem {
    static constexpr int STAGES = BN >= 256 ? 4 : 6;
    static constexpr int A_BYTES = 128 * 64 * 2, B_BYTES = BN * 64 * 2;
    static constexpr int BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 + 256;
};

template <int BN, int EPI>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, ConvParams p) {
    using S = ConvSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;
    uint8_t* sb = smem + S::STAGES * S::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::STAGES * (S::A_BYTES + S::B_BYTES));
    uint64_t* empty = full + S::STAGES;
    uint64_t* tmem_full = empty + S::STAGES;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;        // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_img = p.tiles_h * p.tiles_w;
    const int ntiles = p.B * tiles_per_img;
    const int nkb = 9 * p.chunks;
    constexpr uint32_t TMEM_COLS = 2 * BN;       // 128, 256 or 512: all powers of two

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b);
        for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_smem, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;                     // running k-block counter across tiles (smem ring position)
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
                const int h0 = (r / p.tiles_w) * TILE_H, w0 = (r % p.tiles_w) * TILE_W;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % S::STAGES;
                    const uint32_t ph = (it / S::STAGES) & 1;
                    const int tap = kb / p.chunks, cc = kb - tap * p.chunks;
                    const int kh = tap / 3, kw = tap - kh * 3;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], S::A_BYTES + S::B_BYTES);
                    tma_load_4d(sa + s * S::A_BYTES, &tma_a, &full[s], cc * 64, w0 + p.sign * kw, h0 + p.sign * kh, b);
                    tma_load_2d(sb + s * S::B_BYTES, &tma_b, &full[s], kb * 64, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_bf16(128, BN);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
                const uint32_t acc = tcount & 1, use = tcount >> 1;
                mbar_wait(&tmem_empty[acc], (use & 1) ^ 1);      // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % S::STAGES;
                    const uint32_t ph = (it / S::STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tcgen05_fence_after();
                    const uint32_t a_addr = smem_u32(sa + s * S::A_BYTES), b_addr = smem_u32(sb + s * S::B_BYTES);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16(d_tmem, smem_desc_k_sw128(a_addr + k * 32), smem_desc_k_sw128(b_addr + k * 32), idesc,
                                 (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[acc]);
            }
        }
    } else {
        const int quarter = warp & 3;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const uint32_t acc = tcount & 1, use = tcount >> 1;
            const int b = tile / tiles_per_img, r = tile - b * tiles_per_img;
            const int h0 = (r / p.tiles_w) * TILE_H, w0 = (r % p.tiles_w) * TILE_W;
            mbar_wait(&tmem_full[acc], use & 1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            const int rr = 2 * quarter + (lane >> 4), cc = lane & 15;       // position inside the 8x16 tile
            if (EPI == EPI_STORE) {
                const int h = h0 + rr, w = w0 + cc;
                const bool ok = h < p.valid_h && w < p.valid_w;
                bf16* o = p.dx + (((int64_t)b * p.valid_h + h) * p.valid_w + w) * p.N;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(taddr + c0, v);
                    if (ok) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 u;
                            __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                            for (int t = 0; t < 4; ++t) hh[t] = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
                            *reinterpret_cast<uint4*>(o + c0 + j) = u;
                        }
                    }
                }
            } else {
                // window id: pooled row = h0/2 + quarter, pooled col = w0/2 + cc/2; element e = dy*2+dx
                const int ph_ = (h0 >> 1) + quarter, pw_ = (w0 >> 1) + (cc >> 1);
                const bool ok = ph_ < p.PH && pw_ < p.PW;
                const int bit0 = lane & 1, bit4 = (lane >> 4) & 1;
                const int64_t obase = (((int64_t)b * p.PH + ph_) * p.PW + pw_) * p.N;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(taddr + c0, v);
                    // step 1 (partner lane^1, dx): keep 16 channels: [0,16) if bit0==0 else [16,32)
                    float k1[16]; int i1[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float mine = bit0 ? v[16 + j] : v[j];
                        const float send = bit0 ? v[j] : v[16 + j];
                        const float other = __shfl_xor_sync(0xffffffffu, send, 1);
                        // element ids: mine = bit0, other = bit0^1 (same dy); ties go to the lower id
                        const bool take_other = bit0 ? (other >= mine) : (other > mine);
                        k1[j] = take_other ? other : mine;
                        i1[j] = take_other ? (bit0 ^ 1) : bit0;
                    }
                    // step 2 (partner lane^16, dy): keep 8 channels: first 8 if bit4==0 else last 8
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
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <int BN, int EPI>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb, const ConvParams& p, cudaStream_t st) {
    auto kern = conv_tc_kernel<BN, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ConvSmem<BN>::BYTES));
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = p.B * p.tiles_h * p.tiles_w;
    const int grid = ntiles < sms ? ntiles : sms;
    kern<<<grid, CONV_THREADS, ConvSmem<BN>::BYTES, st>>>(ta, tb, p);
    VQA_CHECK_LAUNCH("conv_tc");
    return 0;
}

// activation tensor map: NHWC bf16 [B, H, W, C] with box [64, 16, 8, 1]
static int act_tmap(CUtensorMap* m, const void* base, int B, int H, int W, int C) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, TILE_W, TILE_H, 1};
    return make_tmap_bf16(m, base, 4, dims, str, box);
}
static int weight_tmap(CUtensorMap* m, const void* base, int N, int K) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, (uint32_t)N};
    return make_tmap_bf16(m, base, 2, dims, str, box);
}

}  // namespace tc

using namespace tc;

// x [B,IH,IW,Cin] bf16 NHWC; wp [Cout][9*Cin] bf16 (tap-major, channel-minor); out/mask [B,PH,PW,Cout]
extern "C" int vqa_tc_conv3x3_relu_pool_fwd(const void* x, const void* wp, const float* bias, void* out, uint8_t* mask,
                                            int B, int IH, int IW, int Cin, int Cout, void* stream) {
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4, "tc conv fwd: bad dims");
    VQA_REQUIRE(Cin % 64 == 0, "tc conv fwd: Cin=%d must be a multiple of 64 (use vqa_conv_relu_pool_fwd)", Cin);
    VQA_REQUIRE(Cout == 64 || Cout == 128 || Cout == 256, "tc conv fwd: Cout=%d must be 64, 128 or 256", Cout);
    const int OH = IH - 2, OW = IW - 2, PH = OH / 2, PW = OW / 2;
    CUtensorMap ta, tb;
    if (int e = act_tmap(&ta, x, B, IH, IW, Cin)) return e;
    if (int e = weight_tmap(&tb, wp, Cout, 9 * Cin)) return e;
    ConvParams p{};
    p.B = B; p.tiles_h = (2 * PH + TILE_H - 1) / TILE_H; p.tiles_w = (2 * PW + TILE_W - 1) / TILE_W;
    p.chunks = Cin / 64; p.sign = 1; p.valid_h = 2 * PH; p.valid_w = 2 * PW; p.N = Cout;
    p.bias = bias; p.pooled = (bf16*)out; p.mask = mask; p.PH = PH; p.PW = PW;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout == 64) return launch_conv<64, EPI_POOL>(ta, tb, p, st);
    if (Cout == 128) return launch_conv<128, EPI_POOL>(ta, tb, p, st);
    return launch_conv<256, EPI_POOL>(ta, tb, p, st);
}

// dy [B,OHp,OWp,Cout] bf16: gradient w.r.t. the conv output restricted to the pooled region (OHp = 2PH,
// OWp = 2PW), zero where a position was not its window's maximum (vqa_unpool_bf16).
// wd [Cin][9*Cout] bf16 with wd[ci][tap][co] = w[co][ci][kh][kw].  dx [B,IH,IW,Cin] bf16.
extern "C" int vqa_tc_conv3x3_bwd_data(const void* dy, const void* wd, void* dx,
                                       int B, int IH, int IW, int Cin, int Cout, void* stream) {
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4, "tc conv dgrad: bad dims");
    VQA_REQUIRE(Cout % 64 == 0, "tc conv dgrad: Cout=%d must be a multiple of 64", Cout);
    VQA_REQUIRE(Cin == 64 || Cin == 128 || Cin == 256, "tc conv dgrad: Cin=%d must be 64, 128 or 256", Cin);
    const int OHp = ((IH - 2) / 2) * 2, OWp = ((IW - 2) / 2) * 2;
    CUtensorMap ta, tb;
    if (int e = act_tmap(&ta, dy, B, OHp, OWp, Cout)) return e;
    if (int e = weight_tmap(&tb, wd, Cin, 9 * Cout)) return e;
    ConvParams p{};
    p.B = B; p.tiles_h = (IH + TILE_H - 1) / TILE_H; p.tiles_w = (IW + TILE_W - 1) / TILE_W;
    p.chunks = Cout / 64; p.sign = -1; p.valid_h = IH; p.valid_w = IW; p.N = Cin;
    p.dx = (bf16*)dx;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 64) return launch_conv<64, EPI_STORE>(ta, tb, p, st);
    if (Cin == 128) return launch_conv<128, EPI_STORE>(ta, tb, p, st);
    return launch_conv<256, EPI_STORE>(ta, tb, p, st);
}

// ------------------------------------------------------------------------------------------
// weight packing (fp32 OIHW -> the two bf16 K-major layouts) and un-pooling of the gradient
// ------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, bf16* __restrict__ wp, bf16* __restrict__ wd,
                                        int Cout, int Cin) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // over Cout*Cin*9 in OIHW order
    if (i >= (int64_t)Cout * Cin * 9) return;
    const int tap = (int)(i % 9);
    const int ci = (int)((i / 9) % Cin);
    const int co = (int)(i / (9 * (int64_t)Cin));
    const bf16 v = __float2bfloat16_rn(w[i]);
    if (wp) wp[((int64_t)co * 9 + tap) * Cin + ci] = v;
    if (wd) wd[((int64_t)ci * 9 + tap) * Cout + co] = v;
}

extern "C" int vqa_pack_conv3x3_weight(const float* w, void* wp, void* wd, int Cout, int Cin, void* stream) {
    VQA_REQUIRE(w && (wp || wd) && Cout > 0 && Cin > 0, "pack_conv_weight: bad arguments");
    const int64_t n = (int64_t)Cout * Cin * 9;
    pack_conv_weight_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(w, (bf16*)wp, (bf16*)wd, Cout, Cin);
    VQA_CHECK_LAUNCH("pack_conv_weight");
    return 0;
}

// dy[b, 2ph+dy, 2pw+dx, c] = (mask[b,ph,pw,c] == dy*2+dx) ? dpool[b,ph,pw,c] : 0      (8 channels per thread)
__global__ void unpool_bf16_kernel(const bf16* __restrict__ dpool, const uint8_t* __restrict__ mask, bf16* __restrict__ dy,
                                   int64_t npos, int PH, int PW, int C) {
    const int c8 = C >> 3;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npos * c8) return;
    const int64_t pos = i / c8;
    const int c0 = (int)(i - pos * c8) * 8;
    const int pw = (int)(pos % PW);
    const int64_t t = pos / PW;
    const int ph = (int)(t % PH);
    const int64_t b = t / PH;
    const uint4 g = *reinterpret_cast<const uint4*>(dpool + pos * C + c0);
    const uint2 m = *reinterpret_cast<const uint2*>(mask + pos * C + c0);
    const uint16_t* gs = reinterpret_cast<const uint16_t*>(&g);
    const uint8_t* ms = reinterpret_cast<const uint8_t*>(&m);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        uint4 o;
        uint16_t* os = reinterpret_cast<uint16_t*>(&o);
#pragma unroll
        for (int j = 0; j < 8; ++j) os[j] = ms[j] == e ? gs[j] : (uint16_t)0;
        const int64_t opos = ((b * 2 * PH + 2 * ph + (e >> 1)) * (2 * PW) + 2 * pw + (e & 1));
        *reinterpret_cast<uint4*>(dy + opos * C + c0) = o;
    }
}

extern "C" int vqa_unpool_bf16(const void* dpool, const uint8_t* mask, void* dy, int B, int PH, int PW, int C, void* stream) {
    VQA_REQUIRE(B > 0 && PH > 0 && PW > 0 && C % 8 == 0, "unpool: bad dims (C must be a multiple of 8)");
    const int64_t npos = (int64_t)B * PH * PW;
    unpool_bf16_kernel<<<(unsigned)ceil_div64(npos * (C / 8), 256), 256, 0, (cudaStream_t)stream>>>(
        (const bf16*)dpool, mask, (bf16*)dy, npos, PH, PW, C);
    VQA_CHECK_LAUNCH("unpool_bf16");
    return 0;
}
