// 3x3 / stride-1 convolutions of the image encoder on tcgen05 tensor cores (models/model.py:72-84).
//
// Implicit GEMM without im2col and without re-reading the input per filter tap: one CTA tile = 16 x 8 spatial
// positions (M = 128) x all output channels (N = BN <= 256).  For a 64-channel slice ONE 4-D TMA box
// [64 ch, 10 w, 18 h, 1 image] brings the tile plus its halo into shared memory as 180 rows x 128 bytes (128-byte
// swizzle); the A operand of filter tap (kh,kw) is the same buffer with the start address advanced by
// (kh*10 + kw) rows and an 8-row-group stride of 10 rows (1280 B) -- the swizzle is a function of the absolute
// shared-memory address (tools/umma_probe.cu), so nine taps cost one load.  Out-of-bounds rows/cols are zero-filled
// by TMA (no padding copies).  The B operand is the packed weight [N][tap][C] (K-major); when all of it fits in
// shared memory (conv1: 144 KB) it is loaded once per CTA and stays resident, otherwise its 64-wide k-slabs are
// streamed through a second mbarrier ring.
//
// NCTA = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) works on two tiles at once; each CTA loads its own halo
// tile and HALF of the weight rows, the pair's tensor cores share both halves, so per-SM shared-memory traffic of
// the B operand halves.  The leader CTA's single MMA thread issues for both; mbarrier completions are multicast.
//
// Persistent kernel: grid = #SMs, tiles round-robin; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 =
// epilogue (two warps per TMEM lane quarter, alternating 32-column chunks).  Two TMEM accumulators (2 x BN columns)
// so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Epilogues:
//   POOL  (forward)  bias + ReLU + 2x2 max-pool + arg-max mask: window partners are lanes l, l^1, l^8, l^9 of
//                    one warp (tile rows are 8 wide); a 2-step butterfly leaves each lane 8 of every 32 channels.
//   STORE (dgrad)    plain bf16 store of the 128 x BN tile (gradient w.r.t. the layer input).
//   UNPOOL (dgrad)   the layer input is the pooled output of the layer below: the tile is written straight into that
//                    layer's UN-POOLED gradient (value at the arg-max element of each 2x2 window, zeros elsewhere) and
//                    its bias gradient is reduced with a transposing butterfly -- vqa_unpool_bf16 disappears.
#include "tc_common.cuh"
#include <atomic>

namespace tc {

constexpr int TILE_H = 16, TILE_W = 8;
constexpr int HALO_H = TILE_H + 2, HALO_W = TILE_W + 2;
constexpr int HALO_BYTES = HALO_H * HALO_W * 128;            // 23040: one 64-channel slice of tile + halo
constexpr int A_STAGE_BYTES = 23552;                          // padded to a multiple of 1024
// Epilogue warps: 8 (two per TMEM lane quarter, alternating 32-column chunks).  (16 for the POOL epilogue measured no
// faster once the arg-max rode in the value bits: conv1 forward is then bound by the MMAs' shared-memory reads.)
// The UNPOOL epilogue (dgrad + max-pool backward + bias gradient of the layer below) is the longest per tile and was what
// the MMAs waited for (tensor pipe 65 %, issue slots 27 %): 16 warps, one 32-column chunk each at BN = 128.
__host__ __device__ constexpr int conv_epi_warps(int epi) { return epi == 2 ? 16 : 8; }
__host__ __device__ constexpr int conv_threads(int epi) { return 64 + conv_epi_warps(epi) * 32; }
constexpr int MAX_A_STAGES = 8, MAX_B_STAGES = 8;
constexpr int SMEM_LIMIT = 232448 - 1024;                     // 227 KB minus alignment slack
enum { EPI_POOL = 0, EPI_STORE = 1, EPI_UNPOOL = 2 };

struct ConvParams {
    int B, tiles_h, tiles_w;       // tile grid per image
    int chunks;                    // 64-channel slices of the reduction channel dimension
    int sign;                      // +1: forward taps (h+kh, w+kw);  -1: data-gradient taps (h-kh, w-kw)
    int valid_h, valid_w;          // extent of valid output positions (conv-output for POOL, input for STORE)
    int N;                         // output channels (== BN)
    int a_stages, b_stages;        // ring depths (b_stages unused when the weights are resident)
    // POOL
    const float* bias; bf16* pooled; uint8_t* mask; int PH, PW;
    // STORE
    bf16* dx;
    // UNPOOL: dgrad fused with the max-pool backward (+ bias gradient) of the layer below: instead of dx it writes that
    // layer's un-pooled gradient udy [B, 2*valid_h, 2*valid_w, N] using its pooling mask umask [B, valid_h, valid_w, N]
    const uint8_t* umask; bf16* udy; float* udb;
};

// MT = tiles per CTA and round.  MT = 2 (streamed weights, BN <= 128): both tiles' MMAs share every weight slab, which
// halves the TMA traffic of the B operand into shared memory and the mbarrier waits per MMA.
template <int BN, int EPI, int NCTA, bool RESIDENT, int MT = 1>
__global__ void __launch_bounds__(conv_threads(EPI), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, ConvParams p) {
    pdl_trigger();
    constexpr int BN_CTA = BN / NCTA;                 // weight rows held by this CTA
    constexpr int B_SLAB = BN_CTA * 128;              // one 64-wide k-slab of them
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const int nkb = 9 * p.chunks;
    uint8_t* sa = smem;
    uint8_t* sb = smem + p.a_stages * A_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sb + (RESIDENT ? nkb : p.b_stages) * B_SLAB);
    uint64_t* fullA = bars;
    uint64_t* emptyA = fullA + MAX_A_STAGES;
    uint64_t* fullB = emptyA + MAX_A_STAGES;
    uint64_t* emptyB = fullB + MAX_B_STAGES;
    uint64_t* tmem_full = emptyB + MAX_B_STAGES;     // [2]
    uint64_t* tmem_empty = tmem_full + 2;            // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* bias_s = reinterpret_cast<float*>(bars + 64);            // POOL: [BN] conv bias, read as broadcast LDS.128

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = NCTA == 2 ? cluster_ctarank() : 0;
    const bool leader = rank == 0;
    const int tiles_per_img = p.tiles_h * p.tiles_w;
    const int ntiles = p.B * tiles_per_img;
    const int nrounds = (ntiles + NCTA * MT - 1) / (NCTA * MT);     // MT tiles per CTA of the pair per round
    const int nclusters = gridDim.x / NCTA, cluster_id = blockIdx.x / NCTA;
    constexpr uint32_t TMEM_COLS = 2 * MT * BN;                       // 128, 256 or 512: all powers of two
    static_assert(TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "accumulators must fit the 512 TMEM columns");
    static_assert(MT == 1 || NCTA == 1, "MT > 1 is implemented for single-CTA MMAs only");

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b);
        for (int i = 0; i < MAX_A_STAGES; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
        for (int i = 0; i < MAX_B_STAGES; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], conv_epi_warps(EPI) * NCTA); }
        fence_barrier_init();
    }
    if (warp == 1) { if (NCTA == 2) tmem_alloc_pair(tmem_base_smem, TMEM_COLS); else tmem_alloc(tmem_base_smem, TMEM_COLS); }
    tcgen05_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();                                       // everything above overlapped the previous kernel's tail
    if (EPI == EPI_POOL) {
        for (int i = threadIdx.x; i < BN; i += conv_threads(EPI)) bias_s[i] = p.bias[i];
        __syncthreads();
    }

    if (warp == 0) {
        // ===== TMA producer (one per CTA; byte counts go to the leader's barriers) =====
        if (lane == 0) {
            if (RESIDENT) {
                if (leader) mbar_expect_tx(&fullB[0], (uint32_t)(NCTA * nkb * B_SLAB));
                for (int kb = 0; kb < nkb; ++kb) {
                    if (NCTA == 2) tma_load_2d_pair(sb + kb * B_SLAB, &tma_b, &fullB[0], kb * 64, (int)rank * BN_CTA);
                    else tma_load_2d(sb + kb * B_SLAB, &tma_b, &fullB[0], kb * 64, 0);
                }
            }
            uint32_t ita = 0, itb = 0;
            for (int round = cluster_id; round < nrounds; round += nclusters) {
                for (int cc = 0; cc < p.chunks; ++cc) {
                    for (int i = 0; i < MT; ++i, ++ita) {
                        const int tile = (round * MT + i) * NCTA + (int)rank;
                        int b = p.B, h0 = 0, w0 = 0;                   // b == B: everything out of bounds -> zero fill
                        if (tile < ntiles) {
                            b = tile / tiles_per_img;
                            const int r = tile - b * tiles_per_img;
                            h0 = (r / p.tiles_w) * TILE_H; w0 = (r % p.tiles_w) * TILE_W;
                        }
                        const int hh = p.sign > 0 ? h0 : h0 - 2, ww = p.sign > 0 ? w0 : w0 - 2;
                        const int s = ita % p.a_stages;
                        const uint32_t ph = (ita / p.a_stages) & 1;
                        mbar_wait(&emptyA[s], ph ^ 1);
                        if (leader) mbar_expect_tx(&fullA[s], NCTA * HALO_BYTES);
                        if (NCTA == 2) tma_load_4d_pair(sa + s * A_STAGE_BYTES, &tma_a, &fullA[s], cc * 64, ww, hh, b);
                        else tma_load_4d(sa + s * A_STAGE_BYTES, &tma_a, &fullA[s], cc * 64, ww, hh, b);
                    }
                    if (!RESIDENT) {
                        for (int tap = 0; tap < 9; ++tap, ++itb) {
                            const int t = itb % p.b_stages;
                            const uint32_t pb = (itb / p.b_stages) & 1;
                            const int kb = tap * p.chunks + cc;
                            mbar_wait(&emptyB[t], pb ^ 1);
                            if (leader) mbar_expect_tx(&fullB[t], NCTA * B_SLAB);
                            if (NCTA == 2) tma_load_2d_pair(sb + t * B_SLAB, &tma_b, &fullB[t], kb * 64, (int)rank * BN_CTA);
                            else tma_load_2d(sb + t * B_SLAB, &tma_b, &fullB[t], kb * 64, 0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the leader CTA's warp 1, all lanes converged, one elected lane issues =====
        if (leader) {
            constexpr uint32_t idesc = idesc_bf16(128 * NCTA, BN);
            const uint32_t elected = elect_one();
            if (RESIDENT) { mbar_wait(&fullB[0], 0); tcgen05_fence_after(); }
            const uint64_t b_desc0 = smem_desc_k_sw128(smem_u32(sb));
            uint32_t ita = 0, itb = 0, tcount = 0;
            for (int round = cluster_id; round < nrounds; round += nclusters, ++tcount) {
                const uint32_t acc = tcount & 1, use = tcount >> 1;
                mbar_wait(&tmem_empty[acc], (use & 1) ^ 1);      // both CTAs' epilogues have drained this accumulator
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + acc * (MT * BN);
                for (int cc = 0; cc < p.chunks; ++cc, ita += MT) {
                    // tap (kh,kw) = the halo tile advanced by whole rows: +8 descriptor units (128 B) per row
                    uint64_t a_desc0[MT];
#pragma unroll
                    for (int i = 0; i < MT; ++i) {
                        const int s = (ita + i) % p.a_stages;
                        mbar_wait(&fullA[s], ((ita + i) / p.a_stages) & 1);
                        a_desc0[i] = smem_desc_k_sw128_shifted(smem_u32(sa + s * A_STAGE_BYTES), HALO_W * 128);
                    }
                    tcgen05_fence_after();
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
                        const int kh = tap / 3, kw = tap - kh * 3;
                        const int row = p.sign > 0 ? kh * HALO_W + kw : (2 - kh) * HALO_W + (2 - kw);
                        uint64_t bd;
                        int t = 0;
                        if (RESIDENT) bd = b_desc0 + (uint64_t)((tap * p.chunks + cc) * (B_SLAB >> 4));
                        else {
                            t = itb % p.b_stages;
                            mbar_wait(&fullB[t], (itb / p.b_stages) & 1);
                            tcgen05_fence_after();
                            bd = b_desc0 + (uint64_t)(t * (B_SLAB >> 4));
                            ++itb;
                        }
#pragma unroll
                        for (int i = 0; i < MT; ++i)
                            umma_issue_k64<NCTA>(d_tmem + i * BN, a_desc0[i] + (uint64_t)(row * 8), bd, idesc,
                                                 (cc > 0 || tap > 0) ? 1u : 0u, elected);
                        if (!RESIDENT) umma_commit_issue<NCTA>(&emptyB[t], elected);
                    }
#pragma unroll
                    for (int i = 0; i < MT; ++i) umma_commit_issue<NCTA>(&emptyA[(ita + i) % p.a_stages], elected);
                }
                umma_commit_issue<NCTA>(&tmem_full[acc], elected);
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32*(w%4)..+31; the two warps of a quarter alternate column chunks =====
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        float dbsum[4] = {0.f, 0.f, 0.f, 0.f};          // UNPOOL: bias-gradient partials, see the epilogue for the lane -> channel map
        // UNPOOL: the pooling-mask bytes of this lane's position do not depend on the MMAs, so they are fetched one tile
        // ahead (their DRAM latency would otherwise sit on the epilogue's critical path four times per round)
        constexpr int CSTEP = 32 * (conv_epi_warps(EPI) / 4);                  // columns covered by one pass of all column groups
        constexpr int NMK = EPI == EPI_UNPOOL ? (BN + CSTEP - 1) / CSTEP : 1;
        uint32_t mnext[NMK][8];
        auto load_masks = [&](int tile) {
            const bool live = tile < ntiles;
            const int b = live ? tile / tiles_per_img : 0, r = live ? tile - b * tiles_per_img : 0;
            const int h = (r / p.tiles_w) * TILE_H + 4 * quarter + (lane >> 3), w = (r % p.tiles_w) * TILE_W + (lane & 7);
            const bool ok = live && h < p.valid_h && w < p.valid_w;
            const uint8_t* mp = p.umask + (((int64_t)b * p.valid_h + h) * p.valid_w + w) * p.N + half * 32;
#pragma unroll
            for (int k = 0; k < NMK; ++k) {
                uint4 m0 = make_uint4(0x04040404u, 0x04040404u, 0x04040404u, 0x04040404u), m1 = m0;   // 4 = dead / outside
                if (ok && half * 32 + CSTEP * k < BN) {
                    m0 = __ldcs(reinterpret_cast<const uint4*>(mp + CSTEP * k));
                    m1 = __ldcs(reinterpret_cast<const uint4*>(mp + CSTEP * k) + 1);
                }
                mnext[k][0] = m0.x; mnext[k][1] = m0.y; mnext[k][2] = m0.z; mnext[k][3] = m0.w;
                mnext[k][4] = m1.x; mnext[k][5] = m1.y; mnext[k][6] = m1.z; mnext[k][7] = m1.w;
            }
        };
        if (EPI == EPI_UNPOOL) load_masks((cluster_id * MT) * NCTA + (int)rank);
        uint32_t tcount = 0;
        for (int round = cluster_id; round < nrounds; round += nclusters, ++tcount) {
            const uint32_t acc = tcount & 1, use = tcount >> 1;
            mbar_wait(&tmem_full[acc], use & 1);
            tcgen05_fence_after();
#pragma unroll 1
            for (int ti = 0; ti < MT; ++ti) {
            const int tile = (round * MT + ti) * NCTA + (int)rank;
            const bool live = tile < ntiles;
            const int b = live ? tile / tiles_per_img : 0, r = live ? tile - b * tiles_per_img : 0;
            const int h0 = (r / p.tiles_w) * TILE_H, w0 = (r % p.tiles_w) * TILE_W;
            const uint32_t taddr = tmem_base + (acc * MT + ti) * BN + ((uint32_t)(quarter * 32) << 16);
            const int rr = 4 * quarter + (lane >> 3), cc = lane & 7;         // position inside the 16x8 tile
            if (EPI == EPI_STORE) {
                // 4 x 4 transpose of the 16-byte pieces inside each lane quad (positions w..w+3 of one row): lane q stores
                // piece q of the four positions, 64 contiguous bytes per position (8 lines per instruction, not 32)
                const int q = lane & 3;
                const int h = h0 + rr, wq = w0 + (cc & 4);
                const bool okh = live && h < p.valid_h;
                bf16* obase = p.dx + (((int64_t)b * p.valid_h + h) * p.valid_w + wq) * p.N + 8 * q;
#pragma unroll 1
                for (int c0 = half * 32; c0 < BN; c0 += 64) {
                    float v[32];
                    tmem_ld_32x32(taddr + c0, v);
                    uint32_t G[4][4];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        G[j >> 2][j & 3] = *reinterpret_cast<const uint32_t*>(&h2);
                    }
#pragma unroll
                    for (int step = 0; step < 2; ++step) {
                        const int off = 1 << step;
                        const bool up = (q & off) != 0;
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            const int lo = step == 0 ? 2 * m : m, hi = lo + off;
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const uint32_t r = __shfl_xor_sync(0xffffffffu, up ? G[lo][t] : G[hi][t], off);
                                if (up) G[lo][t] = r; else G[hi][t] = r;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (okh && wq + j < p.valid_w)
                            *reinterpret_cast<uint4*>(obase + (int64_t)j * p.N + c0) = make_uint4(G[j][0], G[j][1], G[j][2], G[j][3]);
                }
            } else if (EPI == EPI_UNPOOL) {
                // A lane owns one position (TMEM lane) and 32 channels = four 16-byte pieces.  A 4 x 4 transpose inside each
                // lane quad (positions w..w+3 of one row) leaves lane q with piece q of all four positions, so that one
                // store instruction writes 64 contiguous bytes per position (8 lines instead of 32 per instruction).
                const int q = lane & 3;
                const int h = h0 + rr, wq = w0 + (cc & 4);                 // the quad's row and first column
                const bool okh = live && h < p.valid_h;
                uint32_t mcur[NMK][8];
#pragma unroll
                for (int k = 0; k < NMK; ++k)
#pragma unroll
                    for (int j = 0; j < 8; ++j) mcur[k][j] = mnext[k][j];
                load_masks(ti + 1 < MT ? tile + NCTA : ((round + nclusters) * MT) * NCTA + (int)rank);
                bf16* obase = p.udy + (((int64_t)b * 2 * p.valid_h + 2 * h) * (2 * p.valid_w) + 2 * wq) * p.N + half * 32 + 8 * q;
#pragma unroll
                for (int k = 0; k < NMK; ++k) {
                    const int c0 = half * 32 + CSTEP * k;
                    if (c0 >= BN) break;                                     // warp-uniform: more column groups than 32-column chunks
                    float v[32];
                    tmem_ld_32x32(taddr + c0, v);
                    uint32_t G[4][4], M[4][2];          // [piece -> position after the transpose][words]
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                        G[j >> 2][j & 3] = *reinterpret_cast<const uint32_t*>(&h2);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) M[j >> 1][j & 1] = mcur[k][j];
#pragma unroll
                    for (int step = 0; step < 2; ++step) {
                        const int off = 1 << step;
                        const bool up = (q & off) != 0;
#pragma unroll
                        for (int m = 0; m < 2; ++m) {
                            const int lo = step == 0 ? 2 * m : m, hi = lo + off;   // slot pair exchanged in this step
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const uint32_t r = __shfl_xor_sync(0xffffffffu, up ? G[lo][t] : G[hi][t], off);
                                if (up) G[lo][t] = r; else G[hi][t] = r;
                            }
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
                                const uint32_t r = __shfl_xor_sync(0xffffffffu, up ? M[lo][t] : M[hi][t], off);
                                if (up) M[lo][t] = r; else M[hi][t] = r;
                            }
                        }
                    }
                    // now G[j] / M[j] = channels c0 + 8q .. + 7 of position (h, wq + j)
                    float bs[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) bs[i] = 0.f;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        // the 8 mask bytes (0..3 = arg-max element, 4 = ReLU-dead) widened once to fp16 patterns 0x3C00 | id
                        // (normal numbers): one HSET2 per channel pair then gives a 0xFFFF / 0 select mask on the
                        // half-precision pipe (was xor / sub / shift + two PRMT per pair and element on the integer pipe)
                        uint32_t mw[4];
                        asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(mw[0]) : "r"(M[j][0]), "r"(0x3C3C3C3Cu));
                        asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(mw[1]) : "r"(M[j][0]), "r"(0x3C3C3C3Cu));
                        asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(mw[2]) : "r"(M[j][1]), "r"(0x3C3C3C3Cu));
                        asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(mw[3]) : "r"(M[j][1]), "r"(0x3C3C3C3Cu));
                        if (okh && wq + j < p.valid_w) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const uint32_t want = 0x3C003C00u | (0x00010001u * e);
                                uint32_t o[4];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    uint32_t sel;
                                    asm("set.eq.u32.f16x2 %0, %1, %2;" : "=r"(sel) : "r"(mw[i]), "r"(want));
                                    o[i] = G[j][i] & sel;
                                }
                                __stcs(reinterpret_cast<uint4*>(obase + ((int64_t)(e >> 1) * (2 * p.valid_w) + 2 * j + (e & 1)) * p.N + CSTEP * k),
                                       make_uint4(o[0], o[1], o[2], o[3]));
                            }
                        }
                        // bias gradient of the layer below: (bf16) gradients whose ReLU was alive (mask id != 4; the
                        // prefetch fills 4 for positions outside the image)
                        uint32_t a[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint32_t alive;
                            asm("set.ne.u32.f16x2 %0, %1, %2;" : "=r"(alive) : "r"(mw[i]), "r"(0x3C043C04u));
                            a[i] = G[j][i] & alive;
                        }
                        bs[0] += __uint_as_float(a[0] << 16); bs[1] += __uint_as_float(a[0] & 0xffff0000u);
                        bs[2] += __uint_as_float(a[1] << 16); bs[3] += __uint_as_float(a[1] & 0xffff0000u);
                        bs[4] += __uint_as_float(a[2] << 16); bs[5] += __uint_as_float(a[2] & 0xffff0000u);
                        bs[6] += __uint_as_float(a[3] << 16); bs[7] += __uint_as_float(a[3] & 0xffff0000u);
                    }
                    // transposing reduction over the 8 lanes that share q: lane ends with channel 8q + 4*b4 + 2*b3 + b2
                    {
                        int n = 8;
#pragma unroll
                        for (int off = 16; off >= 4; off >>= 1, n >>= 1) {
                            const bool up = (lane & off) != 0;
#pragma unroll
                            for (int i = 0; i < n / 2; ++i) {
                                const float keep = up ? bs[n / 2 + i] : bs[i];
                                const float send = up ? bs[i] : bs[n / 2 + i];
                                bs[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                            }
                        }
                    }
                    dbsum[k] += bs[0];
                }
            } else {
                // window: pooled row = h0/2 + 2*quarter + (lane>>4), pooled col = w0/2 + cc/2; element e = dy*2+dx
                const int ph_ = (h0 >> 1) + 2 * quarter + (lane >> 4), pw_ = (w0 >> 1) + (cc >> 1);
                const bool ok = live && ph_ < p.PH && pw_ < p.PW;
                const int bit0 = lane & 1, bitY = (lane >> 3) & 1;
                const int64_t obase = (((int64_t)b * p.PH + ph_) * p.PW + pw_) * p.N;
                // Max-pool with the arg-max carried IN the value: after the bias add, the two low mantissa bits of every
                // fp32 are replaced by 3 - e (e = this lane's element id inside its window), so a plain fmaxf picks the
                // maximum, ties go to the lowest e for positive values (torch's max_pool2d order; non-positive maxima
                // are ReLU-dead and need no index), and the winner's id rides along -- no index selects, no index
                // shuffles.  Cost: values are truncated by < 2^-21 relative before the bf16 rounding.
                const uint32_t ekey = 3u - (uint32_t)(2 * bitY + bit0);
#pragma unroll 1
                for (int c0 = half * 32; c0 < BN; c0 += 32 * (conv_epi_warps(EPI_POOL) / 4)) {     // half = 0..3 here
                    float v[32];
                    tmem_ld_32x32(taddr + c0, v);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bq = *reinterpret_cast<const float4*>(bias_s + c0 + j);
                        v[j] = __uint_as_float((__float_as_uint(v[j] + bq.x) & ~3u) | ekey);
                        v[j + 1] = __uint_as_float((__float_as_uint(v[j + 1] + bq.y) & ~3u) | ekey);
                        v[j + 2] = __uint_as_float((__float_as_uint(v[j + 2] + bq.z) & ~3u) | ekey);
                        v[j + 3] = __uint_as_float((__float_as_uint(v[j + 3] + bq.w) & ~3u) | ekey);
                    }
                    // step 1 (partner lane^1, dx): keep 16 channels: [0,16) if bit0==0 else [16,32)
                    float k1[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float mine = bit0 ? v[16 + j] : v[j];
                        const float send = bit0 ? v[j] : v[16 + j];
                        k1[j] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 1));
                    }
                    // step 2 (partner lane^8, dy): keep 8 channels: first 8 if bitY==0 else last 8
                    float k2[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float mine = bitY ? k1[8 + j] : k1[j];
                        const float send = bitY ? k1[j] : k1[8 + j];
                        k2[j] = fmaxf(mine, __shfl_xor_sync(0xffffffffu, send, 8));
                    }
                    if (ok) {
                        const int nb = c0 + 16 * bit0 + 8 * bitY;
                        uint4 u; uint2 mk;
                        uint32_t mw[2] = {0u, 0u};
                        float o[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t kb_ = __float_as_uint(k2[j]);
                            const float x = __uint_as_float(kb_ & ~3u);
                            const bool alive = x > 0.f;
                            o[j] = alive ? x : 0.f;
                            mw[j >> 2] |= (alive ? 3u - (kb_ & 3u) : 4u) << (8 * (j & 3));
                        }
                        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                        for (int t = 0; t < 4; ++t) hh[t] = __floats2bfloat162_rn(o[2 * t], o[2 * t + 1]);
                        mk.x = mw[0]; mk.y = mw[1];
                        *reinterpret_cast<uint4*>(p.pooled + obase + nb) = u;
                        *reinterpret_cast<uint2*>(p.mask + obase + nb) = mk;
                    }
                }
            }
            }   // tiles of the round
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { if (NCTA == 2) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]); }
        }
        if (EPI == EPI_UNPOOL) {
#pragma unroll
            for (int k = 0; k < NMK; ++k)
                if (half * 32 + CSTEP * k < BN)
                    atomicAdd(p.udb + half * 32 + CSTEP * k + 8 * (lane & 3) + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1),
                              dbsum[k]);
        }
    }

    tcgen05_fence_before();
    if (NCTA == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        if (NCTA == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// cta_group selection: 1 = single-CTA MMAs, 2 = CTA pairs.  Changed only by vqa_tc_conv_set_cta_group (tests/bench).
static std::atomic<int> g_conv_cta_group{0};   // 0 = per-shape choice (launch_conv), 1 / 2 = forced; atomic: callers may race a set against launches

template <int BN, int EPI, int NCTA, bool RESIDENT, int MT = 1>
static int launch_conv_cfg(const CUtensorMap& ta, const CUtensorMap& tb, ConvParams p, int smem_bytes, cudaStream_t st) {
    auto kern = conv_tc_kernel<BN, EPI, NCTA, RESIDENT, MT>;
    static int attr_bytes = 0;
    if (attr_bytes < smem_bytes) {
        VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_bytes = smem_bytes;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = p.B * p.tiles_h * p.tiles_w;
    const int nrounds = (ntiles + NCTA * MT - 1) / (NCTA * MT);
    int grid = (nrounds < sms / NCTA ? nrounds : sms / NCTA) * NCTA;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(conv_threads(EPI)); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (NCTA == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = NCTA; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (vqa_pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    VQA_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
    VQA_CHECK_LAUNCH("conv_tc");
    return 0;
}

// Chooses resident vs streamed weights and the ring depths from the shared-memory budget, then launches.
template <int BN, int EPI>
static int launch_conv(const CUtensorMap& ta, const CUtensorMap& tb1, const CUtensorMap& tb2, ConvParams p, cudaStream_t st) {
    // CTA pairs halve the B-operand (weight) reads from shared memory.  Measured on B200 at the config.yaml shapes they win
    // where the weights are streamed for a 256-wide tile (conv2 forward: 0.353 -> 0.333 ms) and for the 64-wide data
    // gradient (conv1 dgrad: 0.418 -> 0.406 ms), and lose where one CTA already keeps the weights resident for a 128-wide
    // tile (conv1 forward) or shares each streamed slab between two tiles (MT = 2: conv2 dgrad).
    const int forced = g_conv_cta_group.load(std::memory_order_relaxed);
    const int ncta = forced != 0 ? forced
                   : ((EPI == EPI_POOL && BN == 256) || (EPI == EPI_STORE && BN == 64)) ? 2 : 1;
    const int nkb = 9 * p.chunks;
    const int slab = (BN / ncta) * 128;
    const int bars = 512 + 1024;                   // mbarriers + the staged bias
    const int resident_bytes = nkb * slab;
    const bool resident = resident_bytes + 2 * A_STAGE_BYTES + bars <= SMEM_LIMIT;
    int smem;
    if (resident) {
        int a = (SMEM_LIMIT - bars - resident_bytes) / A_STAGE_BYTES;
        p.a_stages = a > 4 ? 4 : a; p.b_stages = 1;
        smem = resident_bytes + p.a_stages * A_STAGE_BYTES + bars + 1024;
    } else {
        const bool mt2 = ncta == 1 && BN <= 128 && EPI != EPI_POOL;          // two tiles share every streamed weight slab
        p.a_stages = mt2 ? 4 : (p.chunks >= 2 ? 3 : 2);
        int bs = (SMEM_LIMIT - bars - p.a_stages * A_STAGE_BYTES) / slab;
        p.b_stages = bs > MAX_B_STAGES ? MAX_B_STAGES : bs;
        VQA_REQUIRE(p.b_stages >= 2, "tc conv: weight slab does not fit in shared memory");
        smem = p.b_stages * slab + p.a_stages * A_STAGE_BYTES + bars + 1024;
    }
    if (ncta == 2) {
        if (resident) return launch_conv_cfg<BN, EPI, 2, true>(ta, tb2, p, smem, st);
        return launch_conv_cfg<BN, EPI, 2, false>(ta, tb2, p, smem, st);
    }
    if (resident) return launch_conv_cfg<BN, EPI, 1, true>(ta, tb1, p, smem, st);
    if constexpr (BN <= 128 && EPI != EPI_POOL) return launch_conv_cfg<BN, EPI, 1, false, 2>(ta, tb1, p, smem, st);
    else return launch_conv_cfg<BN, EPI, 1, false>(ta, tb1, p, smem, st);
}

// activation tensor map: NHWC bf16 [B, H, W, C] with box [64, 10, 18, 1] (tile + halo)
static int act_tmap(CUtensorMap* m, const void* base, int B, int H, int W, int C) {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    const uint32_t box[4] = {64, HALO_W, HALO_H, 1};
    return make_tmap_bf16(m, base, 4, dims, str, box);
}
// weight tensor map: [N][K] bf16, box = 64 k-columns x `rows` weight rows (all of them, or one CTA's half)
static int weight_tmap(CUtensorMap* m, const void* base, int N, int K, int rows) {
    const uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t str[1] = {(uint64_t)K * 2};
    const uint32_t box[2] = {64, (uint32_t)rows};
    return make_tmap_bf16(m, base, 2, dims, str, box);
}

}  // namespace tc

using namespace tc;

extern "C" int vqa_tc_conv_set_cta_group(int cta_group) {
    VQA_REQUIRE(cta_group >= 0 && cta_group <= 2, "conv cta_group must be 0 (per-shape choice), 1 or 2");
    g_conv_cta_group.store(cta_group, std::memory_order_relaxed);
    return 0;
}

// x [B,IH,IW,Cin] bf16 NHWC; wp [Cout][9*Cin] bf16 (tap-major, channel-minor); out/mask [B,PH,PW,Cout]
extern "C" int vqa_tc_conv3x3_relu_pool_fwd(const void* x, const void* wp, const float* bias, void* out, uint8_t* mask,
                                            int B, int IH, int IW, int Cin, int Cout, void* stream) {
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4, "tc conv fwd: bad dims");
    VQA_REQUIRE(Cin % 64 == 0, "tc conv fwd: Cin=%d must be a multiple of 64 (use vqa_conv_relu_pool_fwd)", Cin);
    VQA_REQUIRE(Cout == 64 || Cout == 128 || Cout == 256, "tc conv fwd: Cout=%d must be 64, 128 or 256", Cout);
    const int OH = IH - 2, OW = IW - 2, PH = OH / 2, PW = OW / 2;
    CUtensorMap ta, tb1, tb2;
    if (int e = act_tmap(&ta, x, B, IH, IW, Cin)) return e;
    if (int e = weight_tmap(&tb1, wp, Cout, 9 * Cin, Cout)) return e;
    if (int e = weight_tmap(&tb2, wp, Cout, 9 * Cin, Cout / 2)) return e;
    ConvParams p{};
    p.B = B; p.tiles_h = (2 * PH + TILE_H - 1) / TILE_H; p.tiles_w = (2 * PW + TILE_W - 1) / TILE_W;
    p.chunks = Cin / 64; p.sign = 1; p.valid_h = 2 * PH; p.valid_w = 2 * PW; p.N = Cout;
    p.bias = bias; p.pooled = (bf16*)out; p.mask = mask; p.PH = PH; p.PW = PW;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cout == 64) return launch_conv<64, EPI_POOL>(ta, tb1, tb2, p, st);
    if (Cout == 128) return launch_conv<128, EPI_POOL>(ta, tb1, tb2, p, st);
    return launch_conv<256, EPI_POOL>(ta, tb1, tb2, p, st);
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
    CUtensorMap ta, tb1, tb2;
    if (int e = act_tmap(&ta, dy, B, OHp, OWp, Cout)) return e;
    if (int e = weight_tmap(&tb1, wd, Cin, 9 * Cout, Cin)) return e;
    if (int e = weight_tmap(&tb2, wd, Cin, 9 * Cout, Cin / 2)) return e;
    ConvParams p{};
    p.B = B; p.tiles_h = (IH + TILE_H - 1) / TILE_H; p.tiles_w = (IW + TILE_W - 1) / TILE_W;
    p.chunks = Cout / 64; p.sign = -1; p.valid_h = IH; p.valid_w = IW; p.N = Cin;
    p.dx = (bf16*)dx;
    cudaStream_t st = (cudaStream_t)stream;
    if (Cin == 64) return launch_conv<64, EPI_STORE>(ta, tb1, tb2, p, st);
    if (Cin == 128) return launch_conv<128, EPI_STORE>(ta, tb1, tb2, p, st);
    return launch_conv<256, EPI_STORE>(ta, tb1, tb2, p, st);
}

// Data gradient fused with the max-pool backward and bias gradient of the layer BELOW (whose pooled output is this
// layer's input): mask_below [B,IH,IW,Cin] is that layer's arg-max mask; dy_below [B,2IH,2IW,Cin] receives its un-pooled
// gradient and db_below [Cin] its bias gradient (overwritten).  Replaces vqa_tc_conv3x3_bwd_data + vqa_unpool_bf16.
extern "C" int vqa_tc_conv3x3_bwd_data_unpool(const void* dy, const void* wd, const uint8_t* mask_below, void* dy_below,
                                              float* db_below, int B, int IH, int IW, int Cin, int Cout, void* stream) {
    VQA_REQUIRE(B > 0 && IH >= 4 && IW >= 4 && mask_below && dy_below && db_below, "tc conv dgrad+unpool: bad arguments");
    VQA_REQUIRE(Cout % 64 == 0, "tc conv dgrad+unpool: Cout=%d must be a multiple of 64", Cout);
    VQA_REQUIRE(Cin == 64 || Cin == 128 || Cin == 256, "tc conv dgrad+unpool: Cin=%d must be 64, 128 or 256", Cin);
    const int OHp = ((IH - 2) / 2) * 2, OWp = ((IW - 2) / 2) * 2;
    cudaStream_t st = (cudaStream_t)stream;
    VQA_CUDA(cudaMemsetAsync(db_below, 0, sizeof(float) * Cin, st));
    CUtensorMap ta, tb1, tb2;
    if (int e = act_tmap(&ta, dy, B, OHp, OWp, Cout)) return e;
    if (int e = weight_tmap(&tb1, wd, Cin, 9 * Cout, Cin)) return e;
    if (int e = weight_tmap(&tb2, wd, Cin, 9 * Cout, Cin / 2)) return e;
    ConvParams p{};
    p.B = B; p.tiles_h = (IH + TILE_H - 1) / TILE_H; p.tiles_w = (IW + TILE_W - 1) / TILE_W;
    p.chunks = Cout / 64; p.sign = -1; p.valid_h = IH; p.valid_w = IW; p.N = Cin;
    p.umask = mask_below; p.udy = (bf16*)dy_below; p.udb = db_below;
    if (Cin == 64) return launch_conv<64, EPI_UNPOOL>(ta, tb1, tb2, p, st);
    if (Cin == 128) return launch_conv<128, EPI_UNPOOL>(ta, tb1, tb2, p, st);
    return launch_conv<256, EPI_UNPOOL>(ta, tb1, tb2, p, st);
}

// ------------------------------------------------------------------------------------------
// weight packing (fp32 OIHW -> the two bf16 K-major layouts) and un-pooling of the gradient
// ------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, bf16* __restrict__ wp, bf16* __restrict__ wd,
                                        int Cout, int Cin) {
    pdl_trigger();
    pdl_wait();
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
    VQA_CUDA(vqa_launch_pdl(pack_conv_weight_kernel, dim3((unsigned)ceil_div64(n, 256)), dim3(256), 0, (cudaStream_t)stream, w, (bf16*)wp, (bf16*)wd, Cout, Cin));
    VQA_CHECK_LAUNCH("pack_conv_weight");
    return 0;
}

// dy[b, 2ph+dy, 2pw+dx, c] = (mask[b,ph,pw,c] == dy*2+dx) ? dpool[b,ph,pw,c] : 0      (8 channels per thread)
// and, in the same pass over (dpool, mask), the bias gradient db[c] = sum over positions of dpool where the ReLU was
// alive (mask < 4).  Grid-stride over positions with a fixed channel chunk per thread, one block reduction + C atomics.
__global__ void __launch_bounds__(256)
unpool_bf16_kernel(const bf16* __restrict__ dpool, const uint8_t* __restrict__ mask, bf16* __restrict__ dy,
                   float* __restrict__ db, int64_t npos, int PH, int PW, int C) {
    pdl_trigger();
    pdl_wait();
    __shared__ float red[256][9];
    const int c8 = C >> 3;
    const int64_t total = npos * c8;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pos = i / c8;
        const int c0 = (int)(i - pos * c8) * 8;
        const int pw = (int)(pos % PW);
        const int64_t t = pos / PW;
        const int ph = (int)(t % PH);
        const int64_t b = t / PH;
        const uint4 g = __ldcs(reinterpret_cast<const uint4*>(dpool + pos * C + c0));
        const uint2 m = __ldcs(reinterpret_cast<const uint2*>(mask + pos * C + c0));
        const uint16_t* gs = reinterpret_cast<const uint16_t*>(&g);
        const uint8_t* ms = reinterpret_cast<const uint8_t*>(&m);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            uint4 o;
            uint16_t* os = reinterpret_cast<uint16_t*>(&o);
#pragma unroll
            for (int j = 0; j < 8; ++j) os[j] = ms[j] == e ? gs[j] : (uint16_t)0;
            const int64_t opos = ((b * 2 * PH + 2 * ph + (e >> 1)) * (2 * PW) + 2 * pw + (e & 1));
            __stcs(reinterpret_cast<uint4*>(dy + opos * C + c0), o);
        }
        if (db) {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (ms[j] < 4) acc[j] += __uint_as_float((uint32_t)gs[j] << 16);
        }
    }
    if (db) {
        // threads with equal (threadIdx.x % c8) own the same 8 channels (blockDim and the grid stride are multiples of c8)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
        __syncthreads();
        if (threadIdx.x < C) {
            const int chunk = threadIdx.x >> 3, j = threadIdx.x & 7;
            float s = 0.f;
            for (int t = chunk; t < 256; t += c8) s += red[t][j];
            atomicAdd(db + threadIdx.x, s);
        }
    }
}

extern "C" int vqa_unpool_bf16(const void* dpool, const uint8_t* mask, void* dy, float* db, int B, int PH, int PW, int C, void* stream) {
    VQA_REQUIRE(B > 0 && PH > 0 && PW > 0 && C % 8 == 0, "unpool: bad dims (C must be a multiple of 8)");
    VQA_REQUIRE(db == nullptr || (C <= 256 && 256 % (C / 8) == 0), "unpool: the fused bias gradient needs C in {8,16,...,256} with 256 %% (C/8) == 0");
    cudaStream_t st = (cudaStream_t)stream;
    if (db) VQA_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * C, st));
    const int64_t npos = (int64_t)B * PH * PW;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = ceil_div64(npos * (C / 8), 256);
    const unsigned grid = (unsigned)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    VQA_CUDA(vqa_launch_pdl(unpool_bf16_kernel, dim3(grid), dim3(256), 0, st, (const bf16*)dpool, mask, (bf16*)dy, db, npos, PH, PW, C));
    VQA_CHECK_LAUNCH("unpool_bf16");
    return 0;
}
