// Persistent LSTM recurrence (forward) for the question encoder (nn.LSTM, models/model.py:164) on tcgen05.
//
// One cooperative launch runs ALL time steps of ALL directions.  The recurrent weights never leave the chip:
// CTA j of direction d owns hidden units [16j, 16j+16), i.e. the 64 rows {gate*H + 16j + u} of W_hh[d]
// (64 x H bf16 = 128 KB at H = 1024) loaded ONCE by TMA into shared memory as the K-major B operand, rows
// ordered u*4+gate so that a thread's accumulator row holds the four gates of each of its 16 units.
// Per step: D[b, 64] = h_{s-1}[b, :] W_slice^T with h streamed from L2 by TMA (4-stage ring), accumulators in
// TMEM (one 128-row tile per 128 samples), then the fused cell epilogue: + x-projection, sigmoid/tanh,
// c/h update, variable-length masking, h_s written back (bf16) for the next step.
// Between steps the 64 CTAs of a direction meet at a global-memory barrier (release: __threadfence + atomicAdd,
// acquire: ld.acquire.gpu + fence.proxy.async before the next TMA reads h_s).  The two directions never wait
// for each other.  Co-residency of all CTAs is guaranteed by the cooperative launch.
// CS = 4 (opt-in, see vqa_tc_lstm_fwd): the CTAs run as clusters of four of the same direction.  Every CTA needs ALL of h_{s-1} each step, so
// without sharing the 64 CTAs of a direction pull 64 x 512 KB through L2 per step (the kernel was L2-bandwidth-bound);
// in a cluster each CTA fetches a quarter of every h tile and TMA-multicasts it to its three peers, a slot is
// released to all four producers by a multicast tcgen05.commit.
#include "tc_common.cuh"
#include <cstdlib>

namespace tc {

constexpr int LSTM_THREADS = 320;       // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (8 warps = 2 m-tiles x 4 quarters)
constexpr int LSTM_STAGES = 4;
constexpr int LSTM_A_BYTES = 128 * 64 * 2;

struct LstmParams {
    bf16* gx;            // [dirs][T][B][4H]  in: x W_ih^T + b ; out: activated gates (saved for backward)
    float* cs;           // [dirs][T][B][H]
    bf16* hs;            // [dirs][T+1][B][H]  slot 0 = zeros, slot s+1 = h after step s
    bf16* qf;            // [B][dirs*H]
    const int64_t* len;  // [B]  (row order of the tensors: the length-sorted lengths when `order` is given)
    const int* order;    // optional [B]: row b of gx/cs/hs is sample order[b] (vqa_length_order); qf is written in sample order
    unsigned int* sync;  // [dirs] zero-initialised counters
    int T, B, H, dirs, ctas_per_dir, mtiles;
    int Bp, b0;          // batch pitch of the tensors and first sequence of this launch (B = sequences in this launch, <= 256)
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int CS>
__global__ void __launch_bounds__(LSTM_THREADS, 1)
lstm_persistent_fwd_kernel(const __grid_constant__ CUtensorMap tma_h, const __grid_constant__ CUtensorMap tma_w, LstmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    const int kblocks = p.H / 64;
    uint8_t* sw = smem;                                       // kblocks x [64 rows][128 B]
    uint8_t* sa = smem + kblocks * 8192;                      // LSTM_STAGES x 16 KB
    uint64_t* full = reinterpret_cast<uint64_t*>(sa + LSTM_STAGES * LSTM_A_BYTES);
    uint64_t* empty = full + LSTM_STAGES;
    uint64_t* w_full = empty + LSTM_STAGES;
    uint64_t* tmem_full = w_full + 1;                          // [2] one per m-tile
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_full + 2);
    int* tile_len = reinterpret_cast<int*>(tmem_base_smem + 2);    // [2] longest sequence of each 128-row tile

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.x / p.ctas_per_dir, j = blockIdx.x - dir * p.ctas_per_dir;
    const int T = p.T, B = p.B, H = p.H;

    // A 128-row tile takes part in step s only while one of its sequences is still running (s < its longest length):
    // with the batch in descending length order (vqa_length_order) the short half of the batch leaves the recurrent
    // GEMM, the h ingest and the step barrier early.  Correct for any row order (unsorted: every tile runs to ~T).
    if (threadIdx.x < 2) tile_len[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x < 256 && (int)threadIdx.x < B) {
        const int64_t l = p.len[p.b0 + threadIdx.x];
        atomicMax(&tile_len[threadIdx.x >> 7], l < 0 ? 0 : (l > T ? T : (int)l));
    }

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_h); tma_prefetch_desc(&tma_w);
        for (int i = 0; i < LSTM_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CS); }
        mbar_init(w_full, 1);
        mbar_init(&tmem_full[0], 1); mbar_init(&tmem_full[1], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_smem, 128);
    tcgen05_fence_before();
    if (CS > 1) cluster_sync_all(); else __syncthreads();      // peers' barriers exist before anything is multicast
    tcgen05_fence_after();
    const uint32_t crank = CS > 1 ? cluster_ctarank() : 0;
    constexpr uint16_t CMASK = (uint16_t)((1u << CS) - 1);
    constexpr int SLICE_ROWS = 128 / CS;
    const uint32_t tmem_base = *tmem_base_smem;
    const int tlen0 = tile_len[0], tlen1 = p.mtiles > 1 ? tile_len[1] : 0;

    if (warp == 0) {
        if (lane == 0) {
            // resident recurrent weights: rows [64j, 64j+64) of the packed matrix of this direction
            mbar_expect_tx(w_full, (uint32_t)kblocks * 8192u);
            for (int kb = 0; kb < kblocks; ++kb) tma_load_3d(sw + kb * 8192, &tma_w, w_full, kb * 64, j * 64, dir);
            uint32_t it = 0;
            const unsigned int* cnt = p.sync + dir;
            const unsigned int per_tile = (unsigned int)p.ctas_per_dir * 4u;       // arriving warps per active tile and step
            unsigned int target = 0;                                               // arrivals of all previous steps
            for (int s = 0; s < T; ++s) {
                const int nact = (s < tlen0) + (s < tlen1);
                if (nact == 0) break;                          // lengths only shrink the active set: nothing left
                if (s > 0) {                                   // h_{s-1} of every CTA of this direction is in L2
                    uint32_t spins = 0;
                    while (ld_acquire_gpu(cnt) < target) {
                        if (++spins > (1u << 24)) { printf("vqa_b200: lstm step barrier timeout (cta %d step %d)\n", blockIdx.x, s); __trap(); }
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                target += per_tile * (unsigned int)nact;
                for (int mt = 0; mt < p.mtiles; ++mt) {
                    if (s >= (mt == 0 ? tlen0 : tlen1)) continue;
                    for (int kb = 0; kb < kblocks; ++kb, ++it) {
                        const int st = it % LSTM_STAGES;
                        const uint32_t ph = (it / LSTM_STAGES) & 1;
                        mbar_wait(&empty[st], ph ^ 1);            // all CS consumers have released the slot
                        mbar_expect_tx(&full[st], LSTM_A_BYTES);
                        if (CS > 1)                               // my quarter of the tile, to every CTA of the cluster
                            tma_load_3d_mcast(sa + st * LSTM_A_BYTES + crank * SLICE_ROWS * 128, &tma_h, &full[st], kb * 64,
                                              s * p.Bp + p.b0 + mt * 128 + (int)crank * SLICE_ROWS, dir, CMASK);
                        else
                            tma_load_3d(sa + st * LSTM_A_BYTES, &tma_h, &full[st], kb * 64, s * p.Bp + p.b0 + mt * 128, dir);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // all lanes converged, one elected lane issues (see tc_common.cuh)
        constexpr uint32_t idesc = idesc_bf16(128, 64);
        const uint32_t elected = elect_one();
        mbar_wait(w_full, 0);
        const uint64_t a_desc0 = smem_desc_k_sw128(smem_u32(sa)), b_desc0 = smem_desc_k_sw128(smem_u32(sw));
        uint32_t it = 0;
        for (int s = 0; s < T; ++s)
            for (int mt = 0; mt < p.mtiles; ++mt) {
                if (s >= (mt == 0 ? tlen0 : tlen1)) continue;  // same skips as the producer: the ring positions stay in step
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int st = it % LSTM_STAGES;
                    mbar_wait(&full[st], (it / LSTM_STAGES) & 1);
                    tcgen05_fence_after();
                    umma_issue_k64<1>(tmem_base + mt * 64, a_desc0 + (uint64_t)(st * (LSTM_A_BYTES >> 4)),
                                      b_desc0 + (uint64_t)(kb * (8192 >> 4)), idesc, kb > 0 ? 1u : 0u, elected);
                    if (CS > 1) umma_commit_mcast_issue(&empty[st], CMASK, elected);
                    else umma_commit_issue<1>(&empty[st], elected);
                }
                umma_commit_issue<1>(&tmem_full[mt], elected);
            }
    } else {
        // ---- epilogue: warp -> (m-tile, lane quarter); thread -> sample b; 16 hidden units x 4 gates in registers
        const int ew = warp - 2;
        const int mt = ew >> 2, quarter = warp & 3;
        const int b = mt * 128 + quarter * 32 + lane;
        const bool active_tile = mt < p.mtiles;
        const int tlen = mt == 0 ? tlen0 : tlen1;
        const bool row_ok = active_tile && b < B;
        const int len = row_ok ? (int)p.len[p.b0 + b] : 0;
        const int u0 = j * 16;
        float c_state[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) c_state[u] = 0.f;
        float h_state[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) h_state[u] = 0.f;
        for (int s = 0; s < T; ++s) {
            if (active_tile) {
                const int64_t row = ((int64_t)dir * T + s) * p.Bp + p.b0 + b;
                const bool step_on = row_ok && s < len;
                // a tile past its longest sequence has no MMAs to wait for and nobody waits for it: its warps only copy the
                // frozen state forward (h for the weight gradient's operand, c, the final cell) and never signal
                const bool tile_on = s < tlen;
                bf16* g = p.gx + row * 4 * H + u0;
                // x-projection (+biases) of this thread's 16 units, 4 gates: 8 x 16 B -- independent of the recurrent
                // MMAs, so in flight while this warp waits for them
                uint4 xq[4][2];
                if (step_on) {
#pragma unroll
                    for (int gate = 0; gate < 4; ++gate) {
                        xq[gate][0] = __ldcs(reinterpret_cast<const uint4*>(g + (int64_t)gate * H));
                        xq[gate][1] = __ldcs(reinterpret_cast<const uint4*>(g + (int64_t)gate * H + 8));
                    }
                }
                if (tile_on) {
                    mbar_wait(&tmem_full[mt], s & 1);          // one commit per active step; active steps are 0 .. tlen-1
                    tcgen05_fence_after();
                }
                uint4 oq[4][2];
#pragma unroll
                for (int half = 0; half < 2; ++half) {          // units 8*half .. 8*half+7 <-> accumulator columns 32*half ..
                    if (!tile_on) break;
                    float v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + mt * 64 + half * 32, v);
                    if (step_on) {
#pragma unroll
                        for (int uu = 0; uu < 8; ++uu) {
                            const int u = half * 8 + uu;
                            float xg[4];
#pragma unroll
                            for (int gate = 0; gate < 4; ++gate)
                                xg[gate] = __bfloat162float(reinterpret_cast<const bf16*>(&xq[gate][half])[uu]);
                            // MUFU.TANH (2^-11 relative error, below the bf16 rounding of the stored gates): the precise
                            // tanhf / expf forms were ~1000 instructions per thread and step on the serial critical path
                            const float gi = sigmoid_approx(v[uu * 4 + 0] + xg[0]);
                            const float gf = sigmoid_approx(v[uu * 4 + 1] + xg[1]);
                            const float gg = tanh_approx(v[uu * 4 + 2] + xg[2]);
                            const float go = sigmoid_approx(v[uu * 4 + 3] + xg[3]);
                            c_state[u] = gf * c_state[u] + gi * gg;
                            h_state[u] = go * tanh_approx(c_state[u]);
                            reinterpret_cast<bf16*>(&oq[0][half])[uu] = __float2bfloat16_rn(gi);
                            reinterpret_cast<bf16*>(&oq[1][half])[uu] = __float2bfloat16_rn(gf);
                            reinterpret_cast<bf16*>(&oq[2][half])[uu] = __float2bfloat16_rn(gg);
                            reinterpret_cast<bf16*>(&oq[3][half])[uu] = __float2bfloat16_rn(go);
                        }
                    }
                }
                if (row_ok) {
                    // h_s first: it is what every CTA of the direction waits for
                    bf16* hdst = p.hs + (((int64_t)dir * (T + 1) + s + 1) * p.Bp + p.b0 + b) * H + u0;
                    uint4 q[2];
                    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
                    for (int t = 0; t < 8; ++t) hh[t] = __floats2bfloat162_rn(h_state[2 * t], h_state[2 * t + 1]);
                    *reinterpret_cast<uint4*>(hdst) = q[0];
                    *reinterpret_cast<uint4*>(hdst + 8) = q[1];
                }
                if (tile_on) {
                    tcgen05_fence_before();
                    // release this warp's h_s stores to the other CTAs of the direction (warps of unused m-tiles stay out)
                    __syncwarp();
                    if (lane == 0) { asm volatile("fence.proxy.async;" ::: "memory"); __threadfence(); atomicAdd(p.sync + dir, 1u); }
                }
                // everything only the backward pass reads goes out after the signal, off the step-to-step critical path
                if (step_on) {
#pragma unroll
                    for (int gate = 0; gate < 4; ++gate) {
                        *reinterpret_cast<uint4*>(g + (int64_t)gate * H) = oq[gate][0];
                        *reinterpret_cast<uint4*>(g + (int64_t)gate * H + 8) = oq[gate][1];
                    }
                }
                if (row_ok) {
                    // state after step s (frozen rows copy forward)
                    float* cdst = p.cs + row * H + u0;
#pragma unroll
                    for (int t = 0; t < 4; ++t)
                        *reinterpret_cast<float4*>(cdst + 4 * t) = make_float4(c_state[4 * t], c_state[4 * t + 1], c_state[4 * t + 2], c_state[4 * t + 3]);
                    if (s == T - 1) {
                        const int sample = p.order ? p.order[p.b0 + b] : p.b0 + b;
                        bf16* qdst = p.qf + (int64_t)sample * p.dirs * H + (int64_t)dir * H + u0;
                        uint4 q[2];
                        __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
                        for (int t = 0; t < 8; ++t) hh[t] = __floats2bfloat162_rn(c_state[2 * t], c_state[2 * t + 1]);
                        *reinterpret_cast<uint4*>(qdst) = q[0];
                        *reinterpret_cast<uint4*>(qdst + 8) = q[1];
                    }
                }
            }
            if (!active_tile) break;
        }
    }
    tcgen05_fence_before();
    if (CS > 1) cluster_sync_all(); else __syncthreads();      // no CTA leaves while peers may still signal its barriers
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// ------------------------------------------------------------------------------------------------------------------
// Persistent LSTM recurrence, BACKWARD: one cooperative launch for all T steps and both directions instead of a
// pointwise kernel + split-K GEMM launch pair per step.  Per step s = T-1 .. 0 every CTA
//   phase P  computes its share of the gate gradients dg_s from (dh, dc, saved gates, cell states)  [grid-stride, 8 units per item]
//   signal   row_cnt[dir, m]: dg_s rows of this row tile complete (waited on by the TMA producers of that row tile)
//   phase G  (s > 0) runs one split-K tile of  dh_{s-1}[b, :] += dg_s[b, k-range] W_hh[k-range, :]  on tcgen05:
//            A = dg_s K-major by TMA, B = the bf16 W_hh as stored [4H, H] (MN-major), fp32 vector reductions into dh
//   signal   tile_cnt[tile]: all partial sums of this dh block are in (waited on by the tile's own nsplit CTAs)
// The mbarrier ring, the TMEM accumulator and the tensor maps live across steps; the two synchronisations per step are
// release (threadfence + atomicAdd) / acquire (ld.acquire.gpu) on per-tile / per-row-tile counters, never grid-wide.
// ------------------------------------------------------------------------------------------------------------------
constexpr int LB_THREADS = 512;          // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue; all sixteen run the pointwise phase
                                         // (148 x 512 threads >= the 65536 items of a step at B = 256, H = 1024: one round)
constexpr int LB_STAGES = 6;
constexpr int LB_A_BYTES = 128 * 64 * 2, LB_B_BYTES = 128 * 64 * 2;

struct LstmBwdParams {
    const bf16* gates; const float* cs; float* dh; float* dc; const bf16* dc_init; bf16* dg; const int64_t* len;
    const int* order;                    // optional: row b of the LSTM tensors is sample order[b] (dc_init is in sample order)
    unsigned int* sync;
    int T, B, H, dirs;
    int mt, nt, nsplit, kbps;            // GEMM tiling: tiles = mt * nt * dirs, each split into nsplit k-ranges of kbps k-blocks
};

// release / acquire on monotonically growing counters in global memory (all CTAs are co-resident: cooperative launch)
__device__ __forceinline__ void cta_signal(unsigned int* cnt) {          // after the CTA's writes / reductions
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(cnt, 1u); }
}
__device__ __forceinline__ void spin_until(const unsigned int* cnt, unsigned int target) {
    uint32_t spins = 0;
    while (ld_acquire_gpu(cnt) < target) {
        if (++spins > (1u << 26)) { printf("vqa_b200: lstm backward wait timeout (cta %d)\n", blockIdx.x); __trap(); }
    }
}

// (A variant that kept a [512 x 128] block of W_hh resident per CTA and streamed only dg measured slower, 0.41 vs 0.33 ms:
// the ring that is left beside 128 KB of weights holds too few bytes in flight, and the partial sums double.)
__global__ void __launch_bounds__(LB_THREADS, 1)
lstm_persistent_bwd_kernel(const __grid_constant__ CUtensorMap tma_dg, const __grid_constant__ CUtensorMap tma_w, LstmBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sa = smem;
    uint8_t* sb = smem + LB_STAGES * LB_A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + LB_STAGES * (LB_A_BYTES + LB_B_BYTES));
    uint64_t* empty = full + LB_STAGES;
    uint64_t* tmem_full = empty + LB_STAGES;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);
    int* tile_len_s = reinterpret_cast<int*>(tmem_base_smem + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = p.T, B = p.B, H = p.H;
    if (threadIdx.x == 0) *tile_len_s = 0;
    __syncthreads();
    {   // longest sequence of this CTA's 128-row tile: steps s >= tile_len have no gate gradient in the tile (see below)
        const int midx_ = ((int)blockIdx.x / p.nsplit / p.nt) % p.mt;
        const int b = midx_ * 128 + (int)threadIdx.x;
        if (threadIdx.x < 128 && b < B) {
            const int64_t l = p.len[b];
            atomicMax(tile_len_s, l < 0 ? 0 : (l > T ? T : (int)l));
        }
    }
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tma_dg); tma_prefetch_desc(&tma_w);
        for (int i = 0; i < LB_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_smem, 128);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;

    // this CTA's GEMM tile (fixed for all steps)
    const int ntile_ctas = p.mt * p.nt * p.dirs * p.nsplit;
    const bool has_tile = (int)blockIdx.x < ntile_ctas;
    int t_ = blockIdx.x;
    const int split = t_ % p.nsplit; t_ /= p.nsplit;
    const int nidx = t_ % p.nt; t_ /= p.nt;
    const int midx = t_ % p.mt; t_ /= p.mt;
    const int dir = has_tile ? t_ : 0;
    const int total_kb = 4 * H / 64;
    const int kb_begin = split * p.kbps;
    const int nkb = has_tile ? max(0, min(total_kb, kb_begin + p.kbps) - kb_begin) : 0;
    const int m0 = midx * 128, n0 = nidx * 128;

    // Synchronisation is point to point, not grid-wide.  The nsplit CTAs of an output tile (dir, m, n) are also the ones
    // that run the pointwise phase for that tile's 128 x 128 (sample, unit) block, so
    //   tile_cnt[tile]  (+1 per CTA after its reductions)  gates the next pointwise phase of those same nsplit CTAs, and
    //   row_cnt[dir, m] (+1 per CTA after its dg stores)   gates the A-operand loads of the nt * nsplit CTAs of that row tile.
    const int tile_id = (dir * p.mt + midx) * p.nt + nidx;
    unsigned int* tile_cnt = p.sync + tile_id;
    unsigned int* row_cnt = p.sync + p.mt * p.nt * p.dirs + dir * p.mt + midx;
    const int h8 = H >> 3;
    uint32_t it = 0;                                        // ring position, continues across steps (producer and MMA warp)
    const uint32_t elected = elect_one();
    constexpr uint32_t idesc = idesc_bf16(128, 128, 0, 1);
    const uint64_t a_desc0 = smem_desc_k_sw128(smem_u32(sa));
    const uint64_t b_desc0 = smem_desc_mn_sw128(smem_u32(sb), 8192);
    // Steps s >= tile_len: every row of this CTA's tile is frozen, so dg_s is zero there and dh_{s-1} gets no
    // contribution: only the pointwise phase runs (zero gate gradients, dc passes through, dc_init enters at s = T-1);
    // it touches this thread's own items only, so those steps need neither of the two synchronisations.  The counters
    // therefore count ACTIVE steps.  All CTAs that share a counter share the row tile and skip the same steps.  With the
    // batch in descending length order (vqa_length_order) the short half of the batch skips about half of its steps.
    const int tile_len = *tile_len_s;

    for (int s = T - 1; s >= 0; --s) {
        if (!has_tile) break;
        if (s >= tile_len) {
            for (int li = split * LB_THREADS + threadIdx.x; li < 128 * 16; li += p.nsplit * LB_THREADS) {
                const int b = m0 + (li >> 4);
                if (b < B)
                    lstm_bwd_pointwise_item8<true>(((int64_t)dir * B + b) * h8 + (n0 >> 3) + (li & 15), p.gates, p.cs, p.dh, p.dc,
                                                   s == T - 1 ? p.dc_init : nullptr, p.dg, p.len, s, T, B, H, p.dirs, p.order);
            }
            continue;
        }
        const unsigned int done = (unsigned int)(tile_len - 1 - s);  // active steps completed so far
        // ---- phase P: this CTA's share (1 / nsplit) of the tile's 128 rows x 16 groups of 8 units
        if (done > 0) {
            if (threadIdx.x == 0) spin_until(tile_cnt, (unsigned int)p.nsplit * done);   // dh of this block is complete
            __syncthreads();
        }
        for (int li = split * LB_THREADS + threadIdx.x; li < 128 * 16; li += p.nsplit * LB_THREADS) {
            const int b = m0 + (li >> 4);
            if (b < B)
                lstm_bwd_pointwise_item8<true>(((int64_t)dir * B + b) * h8 + (n0 >> 3) + (li & 15), p.gates, p.cs, p.dh, p.dc,
                                               s == T - 1 ? p.dc_init : nullptr, p.dg, p.len, s, T, B, H, p.dirs, p.order);
        }
        if (s == 0) break;
        // the weight halves of the first ring-full of k-blocks do not depend on this step's dg: in flight across the barrier
        const int npre = min(nkb, LB_STAGES);
        if (warp == 0 && lane == 0) {
            for (int i = 0; i < npre; ++i) {
                const uint32_t iti = it + i;
                const int st = iti % LB_STAGES;
                mbar_wait(&empty[st], ((iti / LB_STAGES) & 1) ^ 1);
                mbar_expect_tx(&full[st], LB_A_BYTES + LB_B_BYTES);
                const int kc = (kb_begin + i) * 64;
                tma_load_3d(sb + st * LB_B_BYTES, &tma_w, &full[st], n0, kc, dir);
                tma_load_3d(sb + st * LB_B_BYTES + 8192, &tma_w, &full[st], n0 + 64, kc, dir);
            }
        }
        cta_signal(row_cnt);
        // ---- phase G
        if (nkb > 0) {
            if (warp == 0) {
                if (lane == 0) {
                    spin_until(row_cnt, (unsigned int)(p.nt * p.nsplit) * (done + 1));     // dg_s rows of this row tile are written
                    asm volatile("fence.proxy.async;" ::: "memory");          // ... through the generic proxy
                    const int row0 = (dir * T + s) * B + m0;
                    for (int i = 0; i < nkb; ++i, ++it) {
                        const int st = it % LB_STAGES;
                        const int kc = (kb_begin + i) * 64;
                        if (i >= npre) {
                            mbar_wait(&empty[st], ((it / LB_STAGES) & 1) ^ 1);
                            mbar_expect_tx(&full[st], LB_A_BYTES + LB_B_BYTES);
                            tma_load_3d(sb + st * LB_B_BYTES, &tma_w, &full[st], n0, kc, dir);
                            tma_load_3d(sb + st * LB_B_BYTES + 8192, &tma_w, &full[st], n0 + 64, kc, dir);
                        }
                        tma_load_2d(sa + st * LB_A_BYTES, &tma_dg, &full[st], kc, row0);
                    }
                }
            } else if (warp == 1) {
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int st = it % LB_STAGES;
                    mbar_wait(&full[st], (it / LB_STAGES) & 1);
                    tcgen05_fence_after();
                    const uint64_t ad = a_desc0 + (uint64_t)(st * (LB_A_BYTES >> 4)), bd = b_desc0 + (uint64_t)(st * (LB_B_BYTES >> 4));
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k)
                        umma_issue<1>(tmem_base, ad + 2 * k, bd + k * (2048 >> 4), idesc, (i > 0 || k > 0) ? 1u : 0u, elected);
                    umma_commit_issue<1>(&empty[st], elected);
                }
                umma_commit_issue<1>(tmem_full, elected);
            } else if (warp < 6) {
                const int quarter = warp & 3;
                const int m = m0 + quarter * 32 + lane;
                mbar_wait(tmem_full, (uint32_t)(done & 1u));
                tcgen05_fence_after();
                float* orow = p.dh + ((int64_t)dir * B + m) * H + n0;
#pragma unroll 1
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                    if (m < B) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) red_add_f32x4(orow + c0 + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
                tcgen05_fence_before();
            }
        }
        cta_signal(tile_cnt);
        tcgen05_fence_after();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 128); }
}

// w_hh fp32 [4H][H] (gate-stacked i,f,g,o) -> bf16 [H/16][64][H], row u*4+gate of block j = source row gate*H + 16j + u
__global__ void pack_lstm_whh_kernel(const float* __restrict__ w, bf16* __restrict__ wp, int H) {
    pdl_trigger();
    pdl_wait();
    const int64_t i8 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // 8 consecutive k per thread (H % 16 == 0)
    const int h8 = H >> 3;
    if (i8 >= (int64_t)4 * H * h8) return;
    const int k = (int)(i8 % h8) * 8;
    const int r = (int)(i8 / h8);               // packed row
    const int jblk = r >> 6, loc = r & 63, u = loc >> 2, gate = loc & 3;
    const float4* src = reinterpret_cast<const float4*>(w + ((int64_t)gate * H + jblk * 16 + u) * H + k);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    uint4 o;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(a.x, a.y); o.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(a.z, a.w); o.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(b.x, b.y); o.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(b.z, b.w); o.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(wp + (int64_t)r * H + k) = o;
}

}  // namespace tc

using namespace tc;

static int g_lstm_cluster_ok = -1;
// 4 if the persistent LSTM runs as clusters of four with TMA multicast, 1 if the driver rejected that launch, 0 if unknown
extern "C" int vqa_tc_lstm_cluster_size(void) { return g_lstm_cluster_ok < 0 ? 0 : (g_lstm_cluster_ok ? 4 : 1); }

extern "C" int vqa_pack_lstm_whh(const float* w_hh, void* wp, int H, void* stream) {
    VQA_REQUIRE(w_hh && wp && H > 0 && H % 16 == 0, "pack_lstm_whh: bad arguments");
    const int64_t n = (int64_t)4 * H * H;
    VQA_REQUIRE(((uintptr_t)w_hh & 15) == 0 && ((uintptr_t)wp & 15) == 0, "pack_lstm_whh: pointers must be 16-byte aligned");
    VQA_CUDA(vqa_launch_pdl(pack_lstm_whh_kernel, dim3((unsigned)ceil_div64(n / 8, 256)), dim3(256), 0, (cudaStream_t)stream, w_hh, (bf16*)wp, H));
    VQA_CHECK_LAUNCH("pack_lstm_whh");
    return 0;
}

// gx [dirs][T][B][4H] bf16, cs [dirs][T][B][H] fp32, hs [dirs][T+1][B][H] bf16 (slot 0 must be zero),
// qf [B][dirs*H] bf16, wp [dirs][4H][H] bf16 from vqa_pack_lstm_whh, sync: dirs zeroed uint32 counters.
// `order` (optional, from vqa_length_order): the rows of gx / cs / hs are the samples in descending length order
// (row j = sample order[j]) and q_len is the matching len_sorted; qf is still written in SAMPLE order.  The kernels are
// correct for any row order -- the order only decides how early whole 128-row tiles drop out of the recurrence.
extern "C" int vqa_tc_lstm_fwd_ordered(void* gx, float* cs_, void* hs, void* qf, const void* wp, const int64_t* q_len, const int* order,
                                       unsigned int* sync, int T, int B, int H, int dirs, void* stream) {
    VQA_REQUIRE(T > 0 && B > 0 && (dirs == 1 || dirs == 2), "tc lstm: bad dims");
    VQA_REQUIRE(H % 64 == 0 && H >= 64 && H <= 1024, "tc lstm: hidden size %d must be a multiple of 64 and <= 1024 (weights resident in shared memory)", H);
    int dev = 0, sms = 148, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    VQA_REQUIRE(coop, "tc lstm: device does not support cooperative launch");
    const int ctas_per_dir = H / 16;
    VQA_REQUIRE(ctas_per_dir * dirs <= sms, "tc lstm: %d CTAs needed but only %d SMs", ctas_per_dir * dirs, sms);
    cudaStream_t st = (cudaStream_t)stream;

    // Clusters of 4 (TMA multicast of h) are OPT-IN (environment VQA_LSTM_CLUSTER=4): they cut the L2 traffic 4x but
    // measured no faster on B200 (the per-SM ingest of h bounds the step), and Nsight Compute cannot replay a
    // cooperative cluster launch (LaunchFailed), which would break profiling of the whole step.
    int& cluster_ok = g_lstm_cluster_ok;           // -1 unknown, 0 rejected / not requested, 1 works
    if (cluster_ok < 0) {
        const char* env = getenv("VQA_LSTM_CLUSTER");
        if (!(env && atoi(env) == 4)) cluster_ok = 0;
    }
    const int cs = (cluster_ok != 0 && ctas_per_dir % 4 == 0) ? 4 : 1;
    CUtensorMap th, tw;
    {
        const uint64_t dims[3] = {(uint64_t)H, (uint64_t)(T + 1) * B, (uint64_t)dirs};
        const uint64_t str[2] = {(uint64_t)H * 2, (uint64_t)(T + 1) * B * H * 2};
        const uint32_t box[3] = {64, (uint32_t)(128 / cs), 1};
        if (int e = make_tmap_bf16(&th, hs, 3, dims, str, box)) return e;
    }
    {
        const uint64_t dims[3] = {(uint64_t)H, (uint64_t)4 * H, (uint64_t)dirs};
        const uint64_t str[2] = {(uint64_t)H * 2, (uint64_t)4 * H * H * 2};
        const uint32_t box[3] = {64, 64, 1};
        if (int e = make_tmap_bf16(&tw, wp, 3, dims, str, box)) return e;
    }
    const int smem = (H / 64) * 8192 + LSTM_STAGES * LSTM_A_BYTES + 1024 + 256;
    static int attr_smem1 = 0, attr_smem4 = 0;
    if (attr_smem1 < smem) {
        VQA_CUDA(cudaFuncSetAttribute(lstm_persistent_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem1 = smem;
    }
    if (attr_smem4 < smem) {
        VQA_CUDA(cudaFuncSetAttribute(lstm_persistent_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem4 = smem;
    }
    // at most 256 sequences (two 128-row accumulator tiles) per cooperative launch: larger batches run chunk by chunk
    for (int b0 = 0; b0 < B; b0 += 256) {
        LstmParams p{};
        p.gx = (bf16*)gx; p.cs = cs_; p.hs = (bf16*)hs; p.qf = (bf16*)qf; p.len = q_len; p.order = order; p.sync = sync;
        p.T = T; p.B = B - b0 < 256 ? B - b0 : 256; p.Bp = B; p.b0 = b0;
        p.H = H; p.dirs = dirs; p.ctas_per_dir = ctas_per_dir; p.mtiles = (p.B + 127) / 128;
        VQA_CUDA(cudaMemsetAsync(sync, 0, sizeof(unsigned int) * dirs, st));
        vqa_count_launch();
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(ctas_per_dir * dirs); cfg.blockDim = dim3(LSTM_THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 4; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
        cfg.attrs = attr;
        if (cluster_ok != 0 && ctas_per_dir % 4 == 0) {
            cfg.numAttrs = 2;
            cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_persistent_fwd_kernel<4>, th, tw, p);
            if (e == cudaSuccess) { cluster_ok = 1; continue; }
            (void)cudaGetLastError();
            if (cluster_ok == 1) { vqa_set_error("lstm_persistent_fwd (clusters): %s", cudaGetErrorString(e)); return (int)e; }
            cluster_ok = 0;                        // first attempt rejected: rebuild the h map with full-tile boxes, single CTAs
            const uint64_t dims[3] = {(uint64_t)H, (uint64_t)(T + 1) * B, (uint64_t)dirs};
            const uint64_t str[2] = {(uint64_t)H * 2, (uint64_t)(T + 1) * B * H * 2};
            const uint32_t box[3] = {64, 128, 1};
            if (int e2 = make_tmap_bf16(&th, hs, 3, dims, str, box)) return e2;
        }
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, lstm_persistent_fwd_kernel<1>, th, tw, p);
        if (e != cudaSuccess) { vqa_set_error("lstm_persistent_fwd: %s", cudaGetErrorString(e)); return (int)e; }
    }
    return 0;
}

extern "C" int vqa_tc_lstm_fwd(void* gx, float* cs_, void* hs, void* qf, const void* wp, const int64_t* q_len,
                               unsigned int* sync, int T, int B, int H, int dirs, void* stream) {
    return vqa_tc_lstm_fwd_ordered(gx, cs_, hs, qf, wp, q_len, nullptr, sync, T, B, H, dirs, stream);
}

// Backward recurrence in one cooperative launch (see lstm_persistent_bwd_kernel).  gates / dg [dirs][T][B][4H] bf16,
// cs [dirs][T][B][H] fp32, dh [dirs][B][H] fp32 ZEROED by the caller (gradient w.r.t. the final hidden state would be
// added here), dc [dirs][B][H] fp32 scratch, dc_init [B][dirs*H] bf16 = gradient w.r.t. the final cell state,
// whh [dirs][4H][H] bf16 (the recurrent weights as stored), sync: scratch of (tiles + dirs * ceil(B/128)) uint32 (<= 256).  Same results as T calls of
// vqa_lstm_step_bwd_pointwise interleaved with T-1 split-K vqa_tc_gemm calls (up to fp32 summation order).
// `order` as for vqa_tc_lstm_fwd_ordered: rows in descending length order, q_len = len_sorted, dc_init in SAMPLE order.
extern "C" int vqa_tc_lstm_bwd_ordered(const void* gates, const float* cs_, float* dh, float* dc, const void* dc_init, void* dg,
                                       const void* whh, const int64_t* q_len, const int* order, unsigned int* sync, int T, int B, int H,
                                       int dirs, void* stream) {
    VQA_REQUIRE(T > 0 && B > 0 && (dirs == 1 || dirs == 2), "tc lstm bwd: bad dims");
    VQA_REQUIRE(H % 128 == 0, "tc lstm bwd: hidden size %d must be a multiple of 128", H);
    int dev = 0, sms = 148, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    VQA_REQUIRE(coop, "tc lstm bwd: device does not support cooperative launch");
    LstmBwdParams p{};
    p.gates = (const bf16*)gates; p.cs = cs_; p.dh = dh; p.dc = dc; p.dc_init = (const bf16*)dc_init; p.dg = (bf16*)dg;
    p.len = q_len; p.order = order; p.sync = sync; p.T = T; p.B = B; p.H = H; p.dirs = dirs;
    p.mt = (B + 127) / 128; p.nt = H / 128;
    const int tiles = p.mt * p.nt * dirs, total_kb = 4 * H / 64;
    VQA_REQUIRE(tiles <= sms, "tc lstm bwd: %d output tiles but only %d SMs (use the per-step path)", tiles, sms);
    int nsplit = sms / tiles;
    if (nsplit > total_kb / 4) nsplit = total_kb / 4;          // at least 4 k-blocks per split
    if (nsplit < 1) nsplit = 1;
    p.kbps = (total_kb + nsplit - 1) / nsplit;
    p.nsplit = (total_kb + p.kbps - 1) / p.kbps;
    CUtensorMap tdg, tw;
    {
        const uint64_t dims[2] = {(uint64_t)4 * H, (uint64_t)dirs * T * B};
        const uint64_t str[1] = {(uint64_t)4 * H * 2};
        const uint32_t box[2] = {64, 128};
        if (int e = make_tmap_bf16(&tdg, dg, 2, dims, str, box)) return e;
    }
    {
        const uint64_t dims[3] = {(uint64_t)H, (uint64_t)4 * H, (uint64_t)dirs};
        const uint64_t str[2] = {(uint64_t)H * 2, (uint64_t)4 * H * H * 2};
        const uint32_t box[3] = {64, 64, 1};
        if (int e = make_tmap_bf16(&tw, whh, 3, dims, str, box)) return e;
    }
    const int smem = LB_STAGES * (LB_A_BYTES + LB_B_BYTES) + 1024 + 256;
    static int attr_bytes = 0;
    if (attr_bytes < smem) {
        VQA_CUDA(cudaFuncSetAttribute(lstm_persistent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_bytes = smem;
    }
    cudaStream_t st = (cudaStream_t)stream;
    VQA_CUDA(cudaMemsetAsync(sync, 0, sizeof(unsigned int) * (size_t)(tiles + dirs * p.mt), st));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(tiles * p.nsplit); cfg.blockDim = dim3(LB_THREADS); cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    VQA_CUDA(cudaLaunchKernelEx(&cfg, lstm_persistent_bwd_kernel, tdg, tw, p));
    VQA_CHECK_LAUNCH("lstm_persistent_bwd");
    return 0;
}

extern "C" int vqa_tc_lstm_bwd(const void* gates, const float* cs_, float* dh, float* dc, const void* dc_init, void* dg,
                               const void* whh, const int64_t* q_len, unsigned int* sync, int T, int B, int H, int dirs,
                               void* stream) {
    return vqa_tc_lstm_bwd_ordered(gates, cs_, dh, dc, dc_init, dg, whh, q_len, nullptr, sync, T, B, H, dirs, stream);
}
