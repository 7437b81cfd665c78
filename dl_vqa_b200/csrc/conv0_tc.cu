// First image-encoder layer (3 input channels, models/model.py:80 with num_channels[0] = 3) on tcgen05.
//
// K = 3*3*3 = 27 is far too small for a TMA-fed pipeline and the NCHW fp32 network input is not a UMMA operand, so
// builder warps write the im2col rows straight into 128-byte-swizzled shared memory.  Both kernels are WINDOW-major:
// the M / reduction index of a tile is a 2x2 pooling window (8 x 16 windows = 128 per tile) and the four window
// elements e = (dy,dx) are four separate operand tiles built from the window's 4x4x3 input patch (48 loads per
// window instead of 4 x 27):
//   patch_e[window][k] = x[ci][2ph+dy+kh][2pw+dx+kw],  k = ci*9 + kh*3 + kw  (k = 27 holds 1.0, k = 28..31 zero)
// Two elements share one 128-byte row (k-range 0..31 of e at byte 64*(e&1)), so a stage is two 16 KB tiles.
//
//   forward : D[window][e*64 + co] = patch_e W^T -- four N = 64 accumulators side by side in TMEM (256 columns,
//             double buffered).  The 2x2 max-pool is then a per-thread max over four registers per channel: no
//             shuffles, no cross-lane traffic; epilogue = max / arg-max + bias + ReLU + mask, written as 128 B of
//             pooled bf16 and 64 B of mask per window.  The 64x222x222 un-pooled activation never exists.
//   backward: dW[co][k] = sum_{window,e} [mask[window][co] == e] dpool[window][co] * patch_e[window][k].  The builders
//             expand (dpool, mask) into the four masked MN-major A tiles in shared memory, so the 1.6 GB un-pooled
//             gradient never exists either (it used to be written by vqa_unpool_bf16 and re-read twice); the patch
//             tiles are the MN-major B operand (the same bytes the forward uses K-major).  Column k = 27 of the
//             patch is 1.0, hence D[co][27] is the bias gradient.  One [128 x 32] fp32 accumulator per CTA, flushed
//             once with atomics.
// Both kernels are HBM / issue-bound; the tensor core only removes the 27 x 64 FMAs per position.
#include "tc_common.cuh"

namespace tc {

constexpr int C0_K = 27;
constexpr int C0_TILE_BYTES = 128 * 128;           // 128 windows x 128 B
constexpr int C0_WH = 8, C0_WW = 16;               // windows per tile

struct Conv0Params {
    const void* x;                     // [B,3,IH,IW] NCHW, fp32 or fp16 (template parameter XHALF)
    int B, IH, IW, PH, PW, tiles_h, tiles_w;
    // forward
    uint32_t magic_img, magic_w;       // ceil(2^32 / tiles per image), ceil(2^32 / tiles_w); 0 = divisor 1
    const float* w; const float* bias; bf16* pooled; uint8_t* mask;
    // backward
    const bf16* dpool; const uint8_t* bmask; float* dw; float* db; int tiles_per_cta;
};

// Tile index -> (image, first pooled row, first pooled column) without integer division: multiply-high by the host's
// reciprocals, exact while tile * divisor < 2^32 (checked at launch).
template <int WH = C0_WH>
__device__ __forceinline__ void conv0_tile_pos(const Conv0Params& p, int tile, int& b, int& ph0, int& pw0) {
    b = p.magic_img ? (int)__umulhi((uint32_t)tile, p.magic_img) : tile;
    const int r = tile - b * (p.tiles_h * p.tiles_w);
    const int th = p.magic_w ? (int)__umulhi((uint32_t)r, p.magic_w) : r;
    ph0 = th * WH;
    pw0 = (r - th * p.tiles_w) * C0_WW;
}

// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (no wait: batch several, then tmem_ld_wait)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 3-input maximum (one FMNMX3) and saturating FMA (FFMA.SAT: result clamped to [0, 1])
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
    float d;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// The 4x4x3 input patch of window (ph,pw) of image b, straight from global memory (coordinates clamped so that
// out-of-range tile windows read valid memory; their results are discarded / multiplied by zero).
template <bool XHALF>
__device__ __forceinline__ void load_patch_global(const Conv0Params& p, int b, int ph, int pw, float (&v)[3][4][4]) {
    const int phc = min(ph, p.PH - 1), pwc = min(pw, p.PW - 1);
    if (XHALF) {                                      // float16 images (the reference's stored dtype), widened exactly
        const __half* src = reinterpret_cast<const __half*>(p.x) + ((int64_t)b * 3 * p.IH + 2 * phc) * p.IW + 2 * pwc;
        const bool al4 = (p.IW & 1) == 0;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const __half* q = src + ((int64_t)ci * p.IH + r) * p.IW;
                if (al4) {
                    const float2 a = __half22float2(__ldg(reinterpret_cast<const __half2*>(q)));
                    const float2 c = __half22float2(__ldg(reinterpret_cast<const __half2*>(q) + 1));
                    v[ci][r][0] = a.x; v[ci][r][1] = a.y; v[ci][r][2] = c.x; v[ci][r][3] = c.y;
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c) v[ci][r][c] = __half2float(__ldg(q + c));
                }
            }
        return;
    }
    const float* src = reinterpret_cast<const float*>(p.x) + ((int64_t)b * 3 * p.IH + 2 * phc) * p.IW + 2 * pwc;
    const bool al8 = (p.IW & 1) == 0;                 // even row pitch: every (row, 2*pw) address is 8-byte aligned
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float* q = src + ((int64_t)ci * p.IH + r) * p.IW;
            if (al8) {
                const float2 a = __ldg(reinterpret_cast<const float2*>(q)), c = __ldg(reinterpret_cast<const float2*>(q) + 1);
                v[ci][r][0] = a.x; v[ci][r][1] = a.y; v[ci][r][2] = c.x; v[ci][r][3] = c.y;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) v[ci][r][c] = __ldg(q + c);
            }
        }
}
// The same patch from the TMA-staged input region of the tile: fp32 [3][18 rows][36 cols], window (wr, wc) of the tile
constexpr int C0_XROWS = 2 * C0_WH + 2, C0_XCOLS = 2 * C0_WW + 4;      // 18 x 36 (34 used; 36 keeps rows 16-byte multiples)
constexpr int C0_XBYTES = 3 * C0_XROWS * C0_XCOLS * 4;                 // 7776
constexpr int C0_XSTAGE = 8192;
constexpr int C0_XCOLS_H = 40;                                          // fp16 regions: 40 columns = 80-byte rows (TMA boxes need 16-byte multiples)
constexpr int C0_XBYTES_H = 3 * C0_XROWS * C0_XCOLS_H * 2;              // 4320
template <bool XHALF>
__device__ __forceinline__ void load_patch_staged(const uint8_t* xs, int wr, int wc, float (&v)[3][4][4]) {
    if (XHALF) {
        const __half* base = reinterpret_cast<const __half*>(xs) + (2 * wr) * C0_XCOLS_H + 2 * wc;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const __half2* q = reinterpret_cast<const __half2*>(base + (ci * C0_XROWS + r) * C0_XCOLS_H);
                const float2 a = __half22float2(q[0]), c = __half22float2(q[1]);
                v[ci][r][0] = a.x; v[ci][r][1] = a.y; v[ci][r][2] = c.x; v[ci][r][3] = c.y;
            }
        return;
    }
    const float* base = reinterpret_cast<const float*>(xs) + (2 * wr) * C0_XCOLS + 2 * wc;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float2* q = reinterpret_cast<const float2*>(base + (ci * C0_XROWS + r) * C0_XCOLS);
            const float2 a = q[0], c = q[1];
            v[ci][r][0] = a.x; v[ci][r][1] = a.y; v[ci][r][2] = c.x; v[ci][r][3] = c.y;
        }
}
// Writes the four im2col rows of a window into `tiles` = two 16 KB tiles, row m.
__device__ __forceinline__ void store_patch_rows(uint8_t* tiles, int m, const float (&v)[3][4][4]) {
    const int sw = m & 7;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int dy = e >> 1, dx = e & 1;
        float k[32];
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) k[ci * 9 + kh * 3 + kw] = v[ci][dy + kh][dx + kw];
        k[27] = 1.f;                                   // backward: bias-gradient column; forward: the weights hold the bias there,
        k[28] = 1.f;                                   // split into a bf16 head (k = 27) and tail (k = 28)
#pragma unroll
        for (int i = 29; i < 32; ++i) k[i] = 0.f;
        uint8_t* rowp = tiles + (e >> 1) * C0_TILE_BYTES + m * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 u;
            u.x = pack2(k[8 * j], k[8 * j + 1]); u.y = pack2(k[8 * j + 2], k[8 * j + 3]);
            u.z = pack2(k[8 * j + 4], k[8 * j + 5]); u.w = pack2(k[8 * j + 6], k[8 * j + 7]);
            *reinterpret_cast<uint4*>(rowp + (((4 * (e & 1) + j) ^ sw) << 4)) = u;
        }
    }
}
// ------------------------------------------------------------------------------------------ forward
// warps 0-15 epilogue (TMEM lane quarter = w & 3, channel quarter = w >> 2), warps 16-19 builders, warp 20 MMA, warp 21
// TMA producer.  STAGED (image row pitch a multiple of 16 bytes): the producer streams each tile's input region
// [3][18][36] through a shared-memory ring with TMA (zero fill outside the image), so the builders issue no global loads
// at all -- with direct loads they stall on the load/store-unit queue (48 scattered loads per window).
//
// Why 16 epilogue warps (r02 ncu source view of the 8-warp form): the epilogue warps were never idle -- 36 % of their
// samples sat in the tile header (two integer divisions by run-time values, spilled addresses), 42 % in the max-pool
// arithmetic, 22 % in the store tail, all of it dependent chains that two warps per scheduler cannot overlap -- while
// the eight builder warps spent their time waiting for a free stage.  Now one builder group feeds four epilogue warps
// per scheduler, tile coordinates come from multiply-high by host-computed reciprocals, and the output leaves through a
// per-quarter staging tile so that every store instruction writes whole 128-byte lines.
constexpr int C0F_EPI_WARPS = 16;
constexpr int C0F_THREADS = (C0F_EPI_WARPS + 4 + 2) * 32;              // 704
constexpr int C0F_XSTAGES = 6;
constexpr int C0F_STAGES = 3;
constexpr int C0F_STAGE_BYTES = 2 * C0_TILE_BYTES;
constexpr int C0F_STG_QUARTER = 32 * 128 + 32 * 64;                   // pooled [32 windows][128 B] + mask [32 windows][64 B]
constexpr int C0F_SMEM = 64 * 128 + C0F_STAGES * C0F_STAGE_BYTES + C0F_XSTAGES * C0_XSTAGE + 2 * 4 * C0F_STG_QUARTER + 1024 + 1024;

struct Conv0FwdMaps { CUtensorMap x, pooled, mask; };       // input regions (load), pooled output and arg-max mask (stores)

template <bool STAGED, bool XHALF>
__global__ void __maxnreg__(80) conv0_fwd_tc_kernel(const __grid_constant__ Conv0FwdMaps maps, Conv0Params p) {
    const CUtensorMap& tma_x = maps.x;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* sw_tile = smem;                                   // [64 co][128 B], k-range 0..31 used
    uint8_t* sa = smem + 64 * 128;                             // C0F_STAGES x 32 KB
    uint8_t* xs = sa + C0F_STAGES * C0F_STAGE_BYTES;           // C0F_XSTAGES x 8 KB staged input regions
    uint8_t* stg = xs + C0F_XSTAGES * C0_XSTAGE;               // [2 sets][4 quarters] output staging
    uint64_t* a_full = reinterpret_cast<uint64_t*>(stg + 4 * 2 * C0F_STG_QUARTER);
    uint64_t* a_empty = a_full + C0F_STAGES;
    uint64_t* tmem_full = a_empty + C0F_STAGES;               // [2]
    uint64_t* tmem_empty = tmem_full + 2;                      // [2]
    uint64_t* x_full = tmem_empty + 2;                         // [C0F_XSTAGES]
    uint64_t* x_empty = x_full + C0F_XSTAGES;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(x_empty + C0F_XSTAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = p.B * p.tiles_h * p.tiles_w;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C0F_STAGES; ++i) { mbar_init(&a_full[i], 128); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], C0F_EPI_WARPS / 2); }
        for (int i = 0; i < C0F_XSTAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 128); }
        if (STAGED) tma_prefetch_desc(&tma_x);
        tma_prefetch_desc(&maps.pooled);
        tma_prefetch_desc(&maps.mask);
        fence_barrier_init();
    }
    if (warp == C0F_EPI_WARPS + 4) tmem_alloc(tmem_base_smem, 512);
    pdl_wait();                                                // the weights below were written by the previous kernel (Adam)
    if (threadIdx.x < 64) {                                    // weight tile: row = co, 32 k-values: 27 weights, the bias as a bf16
        const int co = threadIdx.x;                            // head + tail against the patch columns that hold 1.0, zeros
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = k < C0_K ? p.w[co * C0_K + k] : 0.f;
        const float bias = p.bias[co];
        v[27] = __bfloat162float(__float2bfloat16_rn(bias));   // the accumulator then carries conv + bias (error <= 2^-17 |bias|)
        v[28] = bias - v[27];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (j < 4) { u.x = pack2(v[8 * j], v[8 * j + 1]); u.y = pack2(v[8 * j + 2], v[8 * j + 3]);
                         u.z = pack2(v[8 * j + 4], v[8 * j + 5]); u.w = pack2(v[8 * j + 6], v[8 * j + 7]); }
            *reinterpret_cast<uint4*>(sw_tile + co * 128 + ((j ^ (co & 7)) << 4)) = u;
        }
        fence_proxy_async();
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    if (warp == C0F_EPI_WARPS + 5) {
        // ---- TMA producer of the staged input regions (one thread), tiles in consumption order
        if (STAGED && lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                int b, ph0, pw0;
                conv0_tile_pos(p, tile, b, ph0, pw0);
                const int q = it % C0F_XSTAGES;
                mbar_wait(&x_empty[q], ((it / C0F_XSTAGES) & 1) ^ 1);
                mbar_expect_tx(&x_full[q], XHALF ? C0_XBYTES_H : C0_XBYTES);
                tma_load_4d(xs + q * C0_XSTAGE, &tma_x, &x_full[q], 2 * pw0, 2 * ph0, 0, b);
            }
        }
    } else if (warp >= C0F_EPI_WARPS && warp < C0F_EPI_WARPS + 4) {
        // ---- builders: thread -> window of the tile
        const int t = threadIdx.x - C0F_EPI_WARPS * 32;
        const int wr = t >> 4, wc = t & 15;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int s = it % C0F_STAGES;
            float v[3][4][4];
            if (STAGED) {
                const int q = it % C0F_XSTAGES;
                mbar_wait(&x_full[q], (it / C0F_XSTAGES) & 1);
                load_patch_staged<XHALF>(xs + q * C0_XSTAGE, wr, wc, v);
                mbar_arrive(&x_empty[q]);                      // the values are in registers: the region can be refilled
            } else {
                int b, ph0, pw0;
                conv0_tile_pos(p, tile, b, ph0, pw0);
                load_patch_global<XHALF>(p, b, ph0 + wr, pw0 + wc, v);
            }
            mbar_wait(&a_empty[s], ((it / C0F_STAGES) & 1) ^ 1);
            store_patch_rows(sa + s * C0F_STAGE_BYTES, t, v);
            fence_proxy_async();
            mbar_arrive(&a_full[s]);
        }
    } else if (warp == C0F_EPI_WARPS + 4) {
        // ---- MMA issuer: all lanes converged, one elected lane issues (see tc_common.cuh)
        constexpr uint32_t idesc = idesc_bf16(128, 64);
        const uint32_t elected = elect_one();
        const uint64_t a_desc0 = smem_desc_k_sw128(smem_u32(sa)), w_desc = smem_desc_k_sw128(smem_u32(sw_tile));
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const uint32_t acc = it & 1, use = it >> 1;
            const int s = it % C0F_STAGES;
            mbar_wait(&tmem_empty[acc], (use & 1) ^ 1);
            mbar_wait(&a_full[s], (it / C0F_STAGES) & 1);
            tcgen05_fence_after();
            const uint64_t ad = a_desc0 + (uint64_t)(s * (C0F_STAGE_BYTES >> 4));
#pragma unroll
            for (uint32_t e = 0; e < 4; ++e) {
                const uint64_t ae = ad + (uint64_t)((e >> 1) * (C0_TILE_BYTES >> 4) + (e & 1) * 4);     // +64 B for odd elements
#pragma unroll
                for (uint32_t k = 0; k < 2; ++k)
                    umma_issue<1>(tmem_base + acc * 256 + e * 64, ae + 2 * k, w_desc + 2 * k, idesc, k, elected);
            }
            umma_commit_issue<1>(&a_empty[s], elected);
            umma_commit_issue<1>(&tmem_full[acc], elected);
        }
    } else {
        // ---- epilogue: two SETS of eight warps take alternate tiles (set = TMEM accumulator), thread = window (TMEM lane)
        // x 32 channels; 2x2 max-pool = max over the four element accumulators.
        // The accumulators already hold conv + bias (bias columns of the weight tile).  Per channel two 3-input FMNMX
        // (max of the four elements and of 0: ReLU folded in) are the only work on the half-rate ALU pipe, which the
        // compare / select form of this epilogue kept 82 % busy (ncu, r02h).  The arg-max is ARITHMETIC on the FMA pipe:
        // s_e = sat(2^100 (max - a_e)) is 0 for a maximal element and 1 otherwise, id = s0 (1 + s1 (1 + s2 (1 + s3))) is
        // the first maximal element -- and 4 when no element reaches 0, the ReLU-dead code -- in packed FMUL2 / FADD2 /
        // FFMA2 over channel pairs, with the 2^23 magic constant folded into the last FFMA2 so that one PRMT gathers
        // the id bytes of four channels.  The 32 windows x 64 channels of a lane quarter (two tile rows) are staged in
        // the shared-memory layout of a 128-byte- (pooled) / 64-byte- (mask) swizzled TMA box and leave with two bulk
        // tensor stores per quarter: no store instructions, no address arithmetic, no bounds checks (TMA clips).
        const int set = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
        uint8_t* sb = stg + (set * 4 + quarter) * C0F_STG_QUARTER;
        const uint32_t bar_id = 1 + set * 4 + quarter;            // named barrier of the two warps (channel halves) of a lane quarter
        // staging writes: pooled piece pi (16 B = 8 channels) of window `lane` at 16-byte slot pi ^ (lane & 7) of its row;
        // mask piece pm (16 channels) of the window at slot pm ^ ((lane >> 1) & 3) of its 64-byte row (conflict-free both ways)
        uint8_t* o_wr = sb + lane * 128;
        const uint32_t o_sw = lane & 7;
        uint8_t* m_wr = sb + 32 * 128 + lane * 64;
        const uint32_t m_sw = (lane >> 1) & 3;
        const bool issuer = half == 0 && lane == 0;               // issues the quarter's two tensor stores
        constexpr float HUGE_ = 0x1p100f;
        uint32_t it = set;
        for (int tile = blockIdx.x + set * gridDim.x; tile < ntiles; tile += 2 * gridDim.x, it += 2) {
            const uint32_t use = it >> 1;                          // acc == set
            int b, ph0, pw0;
            conv0_tile_pos(p, tile, b, ph0, pw0);
            mbar_wait_relaxed(&tmem_full[set], use & 1);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + set * 256 + ((uint32_t)(quarter * 32) << 16) + half * 32;
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the previous stores have read the staging tile
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            uint32_t mw[4];
#pragma unroll
            for (int c0 = 0; c0 < 32; c0 += 8) {
                float v0[8], v1[8], v2[8], v3[8];
                tmem_ld_32x8(taddr + c0, v0);
                tmem_ld_32x8(taddr + 64 + c0, v1);
                tmem_ld_32x8(taddr + 128 + c0, v2);
                tmem_ld_32x8(taddr + 192 + c0, v3);
                tmem_ld_wait();
                uint32_t ow[4];
#pragma unroll
                for (int j = 0; j < 8; j += 4) {
                    float pm[2];
#pragma unroll
                    for (int u = 0; u < 4; u += 2) {
                        const int c = j + u;
                        float2 mx;                                // max over the window and 0 (ReLU folded in; the bias came with the MMA)
                        mx.x = fmax3(fmax3(v0[c], v1[c], v2[c]), v3[c], 0.f);
                        mx.y = fmax3(fmax3(v0[c + 1], v1[c + 1], v2[c + 1]), v3[c + 1], 0.f);
                        const float2 mh = __fmul2_rn(mx, make_float2(HUGE_, HUGE_));
                        const float2 s0 = make_float2(fma_sat(v0[c], -HUGE_, mh.x), fma_sat(v0[c + 1], -HUGE_, mh.y));
                        const float2 s1 = make_float2(fma_sat(v1[c], -HUGE_, mh.x), fma_sat(v1[c + 1], -HUGE_, mh.y));
                        const float2 s2 = make_float2(fma_sat(v2[c], -HUGE_, mh.x), fma_sat(v2[c + 1], -HUGE_, mh.y));
                        const float2 s3 = make_float2(fma_sat(v3[c], -HUGE_, mh.x), fma_sat(v3[c + 1], -HUGE_, mh.y));
                        const float2 one = make_float2(1.f, 1.f);
                        float2 id = __fadd2_rn(s3, one);
                        id = __ffma2_rn(s2, id, one);
                        id = __ffma2_rn(s1, id, one);
                        id = __ffma2_rn(s0, id, make_float2(8388608.f, 0.f));          // (2^23 + id_even, id_odd)
                        pm[u >> 1] = fmaf(id.y, 256.f, id.x);                         // low bytes: id_even, id_odd
                        ow[c >> 1] = pack2(mx.x, mx.y);                               // exactly 0 when ReLU-dead
                    }
                    mw[((c0 & 8) + j) >> 2] = __byte_perm(__float_as_uint(pm[0]), __float_as_uint(pm[1]), 0x5410);
                }
                *reinterpret_cast<uint4*>(o_wr + (((uint32_t)(half * 4 + (c0 >> 3)) ^ o_sw) << 4)) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                if (c0 & 8)
                    *reinterpret_cast<uint4*>(m_wr + (((uint32_t)(half * 2 + (c0 >> 4)) ^ m_sw) << 4)) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
            }
            fence_proxy_async();                                   // staged bytes -> visible to the TMA (async proxy)
            tcgen05_fence_before();                                // every TMEM read of this tile has completed (wait::ld above)
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[set]);          // the accumulator may be overwritten while the stores run
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");          // both channel halves of the 32 windows are staged
            if (issuer) {
                tma_store_4d(&maps.pooled, sb, 0, pw0, ph0 + 2 * quarter, b);
                tma_store_4d(&maps.mask, sb + 32 * 128, 0, pw0, ph0 + 2 * quarter, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // all stores complete before the CTA exits
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == C0F_EPI_WARPS + 4) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// ------------------------------------------------------------------------------------------ backward (weight + bias)
// Unit of work: a HALF tile of 4 x 16 windows.  warps 0-3 final flush, warps 4-11 two builder groups of 128 threads (group
// g builds half tiles g, g+2, ... into its own TWO operand stages), warp 12 MMA, warp 13 TMA producer of the staged input
// regions (STAGED, as in the forward).
//
// What the r02 ncu source views of the earlier forms showed, and what this form does about it:
//  * 128-window tiles with one 96 KB stage per group: a group could not start a tile before the MMAs of its previous one
//    had drained (29 % of the builder samples sat in that wait) and the MMA warp waited for operands 47 % of the time.
//    Half tiles make the stage 48 KB, so every group owns two and never waits for the tensor core.
//  * The pooled gradient and the mask arrived by global loads issued at the top of the tile that needs them: four to five
//    long-scoreboard stall cycles per issued instruction.  They are now loaded into a second register set one whole tile
//    ahead (two sets x 24 registers, the loop is unrolled by two so that both are plain arrays).
//  * The kernel moves about 2 KB of shared memory per window (operand stores, MMA operand reads): M = 64 MMAs read no
//    unused second A block, and the mask expansion is one HSET2 (ALU pipe) + one HFMA2 (FMA pipe) per channel pair and
//    window element instead of HSET2 + LOP3 on the ALU pipe alone.
constexpr int C0B_WH = 4;                                     // window rows of a half tile (C0_WW columns)
constexpr int C0B_WIN = C0B_WH * C0_WW;                       // 64 windows
constexpr int C0B_TILE = C0B_WIN * 128;                       // one operand tile: 64 windows x 128 B
constexpr int C0B_THREADS = 14 * 32;
constexpr int C0B_STAGE_BYTES = 6 * C0B_TILE;                 // 4 masked-gradient tiles (A) + 2 patch tiles (B) = 48 KB
constexpr int C0B_NSTAGE = 4;                                 // two per builder group
constexpr int C0B_XSTAGES = 4;                                // a multiple of the builder-group count ON PURPOSE: a slot's mbarriers then
                                                              // always have the same consumer.  If slots alternated between the groups, a group
                                                              // could reach a slot one full phase early (before the previous occupant's data
                                                              // landed) and a parity wait cannot tell "phase n done" from "phase n-1 not done".
constexpr int C0B_XROWS = 2 * C0B_WH + 2;                     // 10 input rows
constexpr int C0B_XBYTES = 3 * C0B_XROWS * C0_XCOLS * 4, C0B_XBYTES_H = 3 * C0B_XROWS * C0_XCOLS_H * 2;      // 4320 / 2400
constexpr int C0B_XSLOT = 4608;
constexpr int C0B_SMEM = C0B_NSTAGE * C0B_STAGE_BYTES + C0B_XSTAGES * C0B_XSLOT + 1024 + 256;

// the three input rows dyh .. dyh+2 (+ 2*wr) of a window's patch, all channels: v[ci][r][c], from a staged region
template <bool XHALF>
__device__ __forceinline__ void load_rows_staged(const uint8_t* xs, int wr, int wc, int dyh, float (&v)[3][3][4]) {
    if (XHALF) {
        const __half* base = reinterpret_cast<const __half*>(xs) + (2 * wr + dyh) * C0_XCOLS_H + 2 * wc;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const __half2* q = reinterpret_cast<const __half2*>(base + (ci * C0B_XROWS + r) * C0_XCOLS_H);
                const float2 a = __half22float2(q[0]), c = __half22float2(q[1]);
                v[ci][r][0] = a.x; v[ci][r][1] = a.y; v[ci][r][2] = c.x; v[ci][r][3] = c.y;
            }
        return;
    }
    const float* base = reinterpret_cast<const float*>(xs) + (2 * wr + dyh) * C0_XCOLS + 2 * wc;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float2* q = reinterpret_cast<const float2*>(base + (ci * C0B_XROWS + r) * C0_XCOLS);
            const float2 a = q[0], c = q[1];
            v[ci][r][0] = a.x; v[ci][r][1] = a.y; v[ci][r][2] = c.x; v[ci][r][3] = c.y;
        }
}
// the same rows straight from global memory (image rows that TMA cannot address); coordinates clamped, the gradient of
// an out-of-range window is zero
template <bool XHALF>
__device__ __forceinline__ void load_rows_global(const Conv0Params& p, int b, int ph, int pw, int dyh, float (&v)[3][3][4]) {
    const int phc = min(ph, p.PH - 1), pwc = min(pw, p.PW - 1);
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int64_t off = (((int64_t)b * 3 + ci) * p.IH + 2 * phc + dyh + r) * p.IW + 2 * pwc;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                v[ci][r][c] = XHALF ? __half2float(__ldg(reinterpret_cast<const __half*>(p.x) + off + c))
                                    : __ldg(reinterpret_cast<const float*>(p.x) + off + c);
        }
}

// im2col row of element column DX from the three patch rows of its element row: k = ci*9 + kh*3 + kw <- v[ci][kh][DX + kw],
// k = 27 holds 1.0 (bias-gradient column), the rest zero; packed to 4 x 16 bytes of bf16
template <int DX>
__device__ __forceinline__ void pack_patch_row(const float (&v)[3][3][4], uint4 (&u)[4]) {
    float k[32];
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) k[ci * 9 + kh * 3 + kw] = v[ci][kh][DX + kw];
    k[27] = 1.f;
#pragma unroll
    for (int c = 28; c < 32; ++c) k[c] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        u[j].x = pack2(k[8 * j], k[8 * j + 1]); u[j].y = pack2(k[8 * j + 2], k[8 * j + 3]);
        u[j].z = pack2(k[8 * j + 4], k[8 * j + 5]); u[j].w = pack2(k[8 * j + 6], k[8 * j + 7]);
    }
}

template <bool STAGED, bool XHALF>
__global__ void __launch_bounds__(C0B_THREADS, 1) conv0_bwd_tc_kernel(const __grid_constant__ CUtensorMap tma_x, Conv0Params p) {
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* xs = smem + C0B_NSTAGE * C0B_STAGE_BYTES;         // C0B_XSTAGES staged input regions
    uint64_t* full = reinterpret_cast<uint64_t*>(xs + C0B_XSTAGES * C0B_XSLOT);
    uint64_t* empty = full + C0B_NSTAGE;
    uint64_t* tmem_full = empty + C0B_NSTAGE;
    uint64_t* x_full = tmem_full + 1;                          // [C0B_XSTAGES]
    uint64_t* x_empty = x_full + C0B_XSTAGES;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(x_empty + C0B_XSTAGES);
    constexpr uint32_t TMEM_COLS = 32;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total = p.B * p.tiles_h * p.tiles_w;             // half tiles
    const int t_begin = blockIdx.x * p.tiles_per_cta;
    const int t_end = min(total, t_begin + p.tiles_per_cta);
    const int nt = max(0, t_end - t_begin);

    if (threadIdx.x == 0) {
        for (int i = 0; i < C0B_NSTAGE; ++i) { mbar_init(&full[i], 128); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        for (int i = 0; i < C0B_XSTAGES; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 128); }
        if (STAGED) tma_prefetch_desc(&tma_x);
        fence_barrier_init();
    }
    if (warp == 12) tmem_alloc(tmem_base_smem, TMEM_COLS);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();

    if (warp == 13) {
        if (STAGED && lane == 0) {
            for (int i = 0; i < nt; ++i) {
                int b, ph0, pw0;
                conv0_tile_pos<C0B_WH>(p, t_begin + i, b, ph0, pw0);
                const int q = i % C0B_XSTAGES;
                mbar_wait(&x_empty[q], ((i / C0B_XSTAGES) & 1) ^ 1);
                mbar_expect_tx(&x_full[q], XHALF ? C0B_XBYTES_H : C0B_XBYTES);
                tma_load_4d(xs + q * C0B_XSLOT, &tma_x, &x_full[q], 2 * pw0, 2 * ph0, 0, b);
            }
        }
    } else if (warp >= 4 && warp < 12) {
        const int g = (warp - 4) >> 2, t = (threadIdx.x - 128) & 127;
        // gradient / mask: in load j (= window row j of the half tile) thread t takes 16-byte chunk t & 7 of window t >> 3:
        // consecutive lanes read consecutive addresses (four windows = 512 contiguous bytes per warp and instruction) and
        // write the eight swizzled chunks of one 128-byte operand row.
        // patch rows: thread = (window t & 63, element row dyh = t >> 6) builds the two im2col rows e = 2 dyh, 2 dyh + 1.
        const int uw = t >> 3, jc = t & 7;
        const int pwin = t & 63, dyh = t >> 6;
        const int pwr = pwin >> 4, pwc = pwin & 15;
        // loads of half tile i into a register set; out-of-image windows read a clamped (valid) address and are zeroed at use
        auto issue = [&](int i, uint4 (&d)[4], uint2 (&mk)[4]) {
            int b, ph0, pw0;
            conv0_tile_pos<C0B_WH>(p, t_begin + i, b, ph0, pw0);
            const int pw = min(pw0 + uw, p.PW - 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ph = min(ph0 + j, p.PH - 1);
                const uint32_t off = (((uint32_t)b * p.PH + ph) * p.PW + pw) * 64u + 8u * jc;      // element index fits 32 bits (checked at launch)
                d[j] = __ldcs(reinterpret_cast<const uint4*>(p.dpool + off));
                mk[j] = __ldcs(reinterpret_cast<const uint2*>(p.bmask + off));
            }
        };
        // expansion of half tile i (the n-th of this group) from a register set into stage 2 g + (n & 1)
        auto build = [&](int i, uint32_t n, const uint4 (&d)[4], const uint2 (&mk)[4]) {
            const uint32_t s = 2 * g + (n & 1), use = n >> 1;
            uint8_t* stage = smem + s * C0B_STAGE_BYTES;
            int b, ph0, pw0;
            conv0_tile_pos<C0B_WH>(p, t_begin + i, b, ph0, pw0);
            float v[3][3][4];
            if (STAGED) {
                const int q = i % C0B_XSTAGES;
                mbar_wait(&x_full[q], (i / C0B_XSTAGES) & 1);
                load_rows_staged<XHALF>(xs + q * C0B_XSLOT, pwr, pwc, dyh, v);
                mbar_arrive(&x_empty[q]);                      // the values are in registers: the region can be refilled
            } else {
                load_rows_global<XHALF>(p, b, ph0 + pwr, pw0 + pwc, dyh, v);
            }
            mbar_wait(&empty[s], (use & 1) ^ 1);
            // A_e[window][co] = mask == e ? dpool : 0   (four 128-byte rows per window, 128B-swizzled).  The mask bytes are
            // widened to the bf16 patterns 0x3F00 | id (five distinct normal numbers), compared with the element's pattern
            // to 1.0 / 0.0 (HSET2, ALU pipe) and multiplied into the gradient pair (HFMA2, FMA pipe).
            const bool colok = pw0 + uw < p.PW;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int w = 16 * j + uw;
                uint8_t* chunkp = stage + w * 128 + ((jc ^ (w & 7)) << 4);
                const bool ok = colok && ph0 + j < p.PH;
                uint32_t mw[4];
                asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(mw[0]) : "r"(mk[j].x), "r"(0x3F3F3F3Fu));
                asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(mw[1]) : "r"(mk[j].x), "r"(0x3F3F3F3Fu));
                asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(mw[2]) : "r"(mk[j].y), "r"(0x3F3F3F3Fu));
                asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(mw[3]) : "r"(mk[j].y), "r"(0x3F3F3F3Fu));
                const uint32_t dv[4] = {ok ? d[j].x : 0u, ok ? d[j].y : 0u, ok ? d[j].z : 0u, ok ? d[j].w : 0u};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t want = 0x3F003F00u | (0x00010001u * e);
                    uint32_t o[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t sel;
                        asm("set.eq.bf16x2.bf16x2 %0, %1, %2;" : "=r"(sel) : "r"(mw[k]), "r"(want));
                        asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(o[k]) : "r"(dv[k]), "r"(sel));
                    }
                    *reinterpret_cast<uint4*>(chunkp + e * C0B_TILE) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            // B: the two im2col rows of (window, dyh) = one 128-byte row of patch tile dyh
            {
                uint8_t* rowp = stage + (4 + dyh) * C0B_TILE + pwin * 128;
                const int sw = pwin & 7;
                uint4 u0[4], u1[4];
                pack_patch_row<0>(v, u0);
                pack_patch_row<1>(v, u1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    *reinterpret_cast<uint4*>(rowp + ((j ^ sw) << 4)) = u0[j];
                    *reinterpret_cast<uint4*>(rowp + (((4 + j) ^ sw) << 4)) = u1[j];
                }
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
        };
        uint4 dA[4], dB[4]; uint2 mkA[4], mkB[4];
        if (g < nt) issue(g, dA, mkA);
        if (g + 2 < nt) issue(g + 2, dB, mkB);
        uint32_t n = 0;
        for (int i = g; i < nt; i += 4) {
            build(i, n, dA, mkA); ++n;
            if (i + 4 < nt) issue(i + 4, dA, mkA);             // after the proxy fence of the tile just built: it would wait for them
            if (i + 2 < nt) {
                build(i + 2, n, dB, mkB); ++n;
                if (i + 6 < nt) issue(i + 6, dB, mkB);
            }
        }
    } else if (warp == 12) {
        // D[64 co][32 k] += A_e^T B_e over the 64 windows of the half tile, e = 0..3.  M = 64 MMAs: an M = 128 instruction
        // would read a second, unused 64-channel block of A with every K step.  A: MN-major, 64 channels = one 128-byte
        // block.  Accumulator layout of M = 64: row r sits in TMEM lane 32 (r / 16) + r % 16 (16 lanes of every lane quarter).
        constexpr uint32_t idesc = idesc_bf16(64, 32, 1, 1);
        const uint32_t elected = elect_one();
        const uint64_t a_desc0 = smem_desc_mn_sw128(smem_u32(smem), 1024);
        for (int i = 0; i < nt; ++i) {
            const uint32_t g = i & 1, n = (uint32_t)i >> 1, s = 2 * g + (n & 1);
            mbar_wait(&full[s], (n >> 1) & 1);
            tcgen05_fence_after();
            const uint64_t sd = a_desc0 + (uint64_t)(s * (C0B_STAGE_BYTES >> 4));
#pragma unroll
            for (uint32_t e = 0; e < 4; ++e) {
                const uint64_t ad = sd + (uint64_t)(e * (C0B_TILE >> 4));
                const uint64_t bd = sd + (uint64_t)((4 + (e >> 1)) * (C0B_TILE >> 4) + (e & 1) * 4);
#pragma unroll
                for (uint32_t k = 0; k < C0B_WIN / 16; ++k)       // 16 windows (rows) per MMA: 2048 bytes further
                    umma_issue<1>(tmem_base, ad + k * (2048 >> 4), bd + k * (2048 >> 4), idesc, (i > 0 || e > 0 || k > 0) ? 1u : 0u, elected);
            }
            umma_commit_issue<1>(&empty[s], elected);
        }
        umma_commit_issue<1>(tmem_full, elected);
    } else if (warp < 4 && nt > 0) {
        mbar_wait(tmem_full, 0);
        tcgen05_fence_after();
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
        if (lane < 16) {
            const int co = warp * 16 + lane;
#pragma unroll
            for (int k = 0; k < C0_K; ++k) atomicAdd(p.dw + co * C0_K + k, v[k]);
            atomicAdd(p.db + co, v[27]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 12) { tcgen05_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

}  // namespace tc

using namespace tc;

static int sm_count() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

static uint32_t magic_u32(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((((uint64_t)1 << 32) + d - 1) / d); }
static int conv0_set_magics(Conv0Params& p) {
    const uint64_t tpi = (uint64_t)p.tiles_h * p.tiles_w;
    VQA_REQUIRE((uint64_t)p.B * tpi * tpi < ((uint64_t)1 << 32), "tc conv0: %d images of %llu tiles exceed the tile index range", p.B, (unsigned long long)tpi);
    VQA_REQUIRE((uint64_t)p.B * p.PH * p.PW * 64 < ((uint64_t)1 << 32), "tc conv0: output of %d x %d x %d x 64 elements exceeds the 32-bit index range", p.B, p.PH, p.PW);
    p.magic_img = magic_u32((uint32_t)tpi);
    p.magic_w = magic_u32((uint32_t)p.tiles_w);
    return 0;
}

// Tensor map of the NCHW network input for the staged (TMA) path; false when the layout does not allow it
static bool conv0_input_map(CUtensorMap* tx, const void* x, bool half, int B, int IH, int IW, int* err, int box_rows = C0_XROWS) {
    const int esz = half ? 2 : 4;
    *err = 0;
    if ((IW * esz) % 16 != 0 || ((uintptr_t)x & 15) != 0) return false;      // TMA needs 16-byte aligned rows
    const uint64_t dims[4] = {(uint64_t)IW, (uint64_t)IH, 3, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)IW * esz, (uint64_t)IH * IW * esz, (uint64_t)3 * IH * IW * esz};
    const uint32_t box[4] = {(uint32_t)(half ? C0_XCOLS_H : C0_XCOLS), (uint32_t)box_rows, 3, 1};
    *err = make_tmap(tx, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, CU_TENSOR_MAP_SWIZZLE_NONE,
                     x, 4, dims, str, box);
    return *err == 0;
}

template <typename Maps, void (*Kern)(const Maps, Conv0Params)>
static int conv0_launch(int grid, int threads, int smem, cudaStream_t st, const Maps& tx, const Conv0Params& p) {
    static bool attr_set = false;               // one flag per kernel instantiation
    if (!attr_set) { VQA_CUDA(cudaFuncSetAttribute(Kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
    VQA_CUDA(vqa_launch_pdl(Kern, dim3(grid), dim3(threads), smem, st, tx, p));
    return 0;
}

// x [B,3,IH,IW] NCHW, x_dtype VQA_F32 or VQA_F16; w [64,3,3,3] fp32; out/mask [B,PH,PW,64] (bf16 / uint8)
extern "C" int vqa_tc_conv0_relu_pool_fwd_x(const void* x, int x_dtype, const float* w, const float* bias, void* out, uint8_t* mask,
                                            int B, int IH, int IW, int Cin, int Cout, void* stream) {
    VQA_REQUIRE(Cin == 3 && Cout == 64, "tc conv0 fwd: only Cin=3, Cout=64 (got %d, %d); use vqa_conv_relu_pool_fwd", Cin, Cout);
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4, "tc conv0 fwd: bad dims");
    VQA_REQUIRE(x_dtype == VQA_F32 || x_dtype == VQA_F16, "tc conv0 fwd: input dtype %d (fp32 or fp16)", x_dtype);
    const bool half = x_dtype == VQA_F16;
    Conv0Params p{};
    p.x = x; p.B = B; p.IH = IH; p.IW = IW;
    p.PH = (IH - 2) / 2; p.PW = (IW - 2) / 2;
    p.tiles_h = (p.PH + C0_WH - 1) / C0_WH; p.tiles_w = (p.PW + C0_WW - 1) / C0_WW;
    p.w = w; p.bias = bias; p.pooled = (bf16*)out; p.mask = mask;
    const int smem = C0F_SMEM;
    const int ntiles = B * p.tiles_h * p.tiles_w;
    if (int rc = conv0_set_magics(p)) return rc;
    const int sms = sm_count();
    const int grid = ntiles < sms ? ntiles : sms;
    Conv0FwdMaps maps{};
    int err = 0;
    const bool staged = conv0_input_map(&maps.x, x, half, B, IH, IW, &err);
    if (err) return err;
    {   // output boxes: 64 channels x 16 windows x 2 rows = the staging tile of one lane quarter, swizzled like its rows
        const uint64_t dims[4] = {64, (uint64_t)p.PW, (uint64_t)p.PH, (uint64_t)B};
        const uint32_t box[4] = {64, C0_WW, 2, 1};
        const uint64_t so[3] = {128, (uint64_t)p.PW * 128, (uint64_t)p.PH * p.PW * 128};
        if (int e = make_tmap(&maps.pooled, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B, out, 4, dims, so, box)) return e;
        const uint64_t sm[3] = {64, (uint64_t)p.PW * 64, (uint64_t)p.PH * p.PW * 64};
        if (int e = make_tmap(&maps.mask, CU_TENSOR_MAP_DATA_TYPE_UINT8, CU_TENSOR_MAP_SWIZZLE_64B, mask, 4, dims, sm, box)) return e;
    }
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (staged) rc = half ? conv0_launch<Conv0FwdMaps, conv0_fwd_tc_kernel<true, true>>(grid, C0F_THREADS, smem, st, maps, p)
                          : conv0_launch<Conv0FwdMaps, conv0_fwd_tc_kernel<true, false>>(grid, C0F_THREADS, smem, st, maps, p);
    else rc = half ? conv0_launch<Conv0FwdMaps, conv0_fwd_tc_kernel<false, true>>(grid, C0F_THREADS, smem, st, maps, p)
                   : conv0_launch<Conv0FwdMaps, conv0_fwd_tc_kernel<false, false>>(grid, C0F_THREADS, smem, st, maps, p);
    if (rc) return rc;
    VQA_CHECK_LAUNCH("conv0_fwd_tc");
    return 0;
}

extern "C" int vqa_tc_conv0_relu_pool_fwd(const float* x, const float* w, const float* bias, void* out, uint8_t* mask,
                                          int B, int IH, int IW, int Cin, int Cout, void* stream) {
    return vqa_tc_conv0_relu_pool_fwd_x(x, VQA_F32, w, bias, out, mask, B, IH, IW, Cin, Cout, stream);
}

// x [B,3,IH,IW] NCHW (fp32 / fp16); dpool [B,PH,PW,64] bf16 = gradient w.r.t. the pooled output; mask [B,PH,PW,64] from the
// forward; dw [64,3,3,3] and db [64] fp32 (both overwritten)
extern "C" int vqa_tc_conv0_bwd_weight_bias_x(const void* x, int x_dtype, const void* dpool, const uint8_t* mask, float* dw, float* db,
                                              int B, int IH, int IW, int Cin, int Cout, void* stream) {
    VQA_REQUIRE(Cin == 3 && Cout == 64, "tc conv0 backward: only Cin=3, Cout=64 (got %d, %d)", Cin, Cout);
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4, "tc conv0 backward: bad dims");
    VQA_REQUIRE(x_dtype == VQA_F32 || x_dtype == VQA_F16, "tc conv0 backward: input dtype %d (fp32 or fp16)", x_dtype);
    const bool half = x_dtype == VQA_F16;
    cudaStream_t st = (cudaStream_t)stream;
    VQA_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * 64 * C0_K, st));
    VQA_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * 64, st));
    Conv0Params p{};
    p.x = x; p.B = B; p.IH = IH; p.IW = IW;
    p.PH = (IH - 2) / 2; p.PW = (IW - 2) / 2;
    p.tiles_h = (p.PH + C0B_WH - 1) / C0B_WH; p.tiles_w = (p.PW + C0_WW - 1) / C0_WW;       // half tiles of 4 x 16 windows
    p.dpool = (const bf16*)dpool; p.bmask = mask; p.dw = dw; p.db = db;
    if (int rc = conv0_set_magics(p)) return rc;
    const int total = B * p.tiles_h * p.tiles_w;
    const int sms = sm_count();
    int ctas = total < sms ? total : sms;
    p.tiles_per_cta = (total + ctas - 1) / ctas;
    ctas = (total + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const int smem = C0B_SMEM;
    CUtensorMap tx{};
    int err = 0;
    const bool staged = conv0_input_map(&tx, x, half, B, IH, IW, &err, C0B_XROWS);
    if (err) return err;
    int rc;
    if (staged) rc = half ? conv0_launch<CUtensorMap, conv0_bwd_tc_kernel<true, true>>(ctas, C0B_THREADS, smem, st, tx, p)
                          : conv0_launch<CUtensorMap, conv0_bwd_tc_kernel<true, false>>(ctas, C0B_THREADS, smem, st, tx, p);
    else rc = half ? conv0_launch<CUtensorMap, conv0_bwd_tc_kernel<false, true>>(ctas, C0B_THREADS, smem, st, tx, p)
                   : conv0_launch<CUtensorMap, conv0_bwd_tc_kernel<false, false>>(ctas, C0B_THREADS, smem, st, tx, p);
    if (rc) return rc;
    VQA_CHECK_LAUNCH("conv0_bwd_tc");
    return 0;
}

extern "C" int vqa_tc_conv0_bwd_weight_bias(const float* x, const void* dpool, const uint8_t* mask, float* dw, float* db,
                                            int B, int IH, int IW, int Cin, int Cout, void* stream) {
    return vqa_tc_conv0_bwd_weight_bias_x(x, VQA_F32, dpool, mask, dw, db, B, IH, IW, Cin, Cout, stream);
}
