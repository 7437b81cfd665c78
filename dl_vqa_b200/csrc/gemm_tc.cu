// bf16 tensor-core GEMM for sm_100a: TMA -> 128B-swizzled shared memory -> tcgen05.mma (fp32 accumulators in
// TMEM) -> tcgen05.ld epilogue.  C[z][m,n] = sum_k A[z][m,k] * B[z][n,k]  (both operands K-major bf16).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (one TMEM lane quarter each).  One 128 x BN output tile per CTA, BK = 64,
// STAGES-deep mbarrier ring.  Split-K (gridDim.z = nbatch * nsplit) accumulates with fp32 atomics.
//
// Used for: the hoisted LSTM input projection, attention.v_conv / q_lin, classifier.lin1 / lin2 and all of
// their data/weight gradients (operands pre-transposed by vqa_transpose_bf16 where the reduction index is
// not the contiguous one).
#include "tc_common.cuh"
#include <mutex>

namespace tc {

// ------------------------------------------------------------------------------------------ host tensor maps
PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    });
    return fn;
}

int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides) {
    PFN_encodeTiled enc = get_encode_tiled();
    VQA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    VQA_REQUIRE(((uintptr_t)base & 15) == 0, "TMA: base pointer %p is not 16-byte aligned", base);
    cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        VQA_REQUIRE((gstr[i] & 15) == 0, "TMA: stride %llu of dim %d is not a multiple of 16 bytes",
                    (unsigned long long)gstr[i], i + 1);
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

int make_tmap(CUtensorMap* out, CUtensorMapDataType dtype, CUtensorMapSwizzle swizzle, const void* base, int rank,
              const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box) {
    PFN_encodeTiled enc = get_encode_tiled();
    VQA_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    VQA_REQUIRE(((uintptr_t)base & 15) == 0, "TMA: base pointer %p is not 16-byte aligned", base);
    cuuint64_t gdim[5]; cuuint64_t gstr[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) {
        gstr[i] = strides_bytes[i];
        VQA_REQUIRE((gstr[i] & 15) == 0, "TMA: stride %llu of dim %d is not a multiple of 16 bytes", (unsigned long long)gstr[i], i + 1);
    }
    CUresult r = enc(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

// ------------------------------------------------------------------------------------------ kernel
constexpr int BM = 128, BK = 64, GEMM_THREADS = 192;

struct GemmEpilogue {
    void* out; int out_bf16; int64_t ldc, c_sb;       // out_bf16: 0 = fp32, 1 = bf16, 2 = fp16 (16-bit outputs share every store path)
    const float* bias; const float* bias2; int64_t bias_sb;
    int relu, atomic, use_dropout;
    uint32_t site; Dropout drop;
};

// ST = ring depth.  6 stages (one CTA per SM) for few, long tiles; 3 stages (96 KB at BN = 128, two or three CTAs
// per SM, each with its own TMEM columns) when there are many short tiles, so that one CTA's prologue / epilogue
// overlaps another CTA's MMAs (attention.v_conv: 10816 tiles of only 4 k-blocks each).
template <int BN, int ST>
struct GemmSmem {
    static constexpr int STAGES = ST;
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Epilogue of one accumulator row chunk: v[0..31] = columns nb..nb+31 of output row m (batch `batch`): bias / ReLU /
// dropout / conversion / store (or fp32 atomics for split-K).
// Called by all 32 lanes of the warp (lane = TMEM lane = output row); `row_ok` masks the stores of rows beyond M.
// `bias_lane` = bias[nb + lane] + bias2[nb + lane] (0 beyond N): one coalesced load per chunk, fetched by the caller
// before it waits for the accumulator, and broadcast here with shuffles -- 32 (x2) same-address loads per lane after
// the TMEM wait doubled the time of the LSTM input projection.
// two fp32 -> one packed 16-bit pair: bf16x2, or f16x2 (saturating to the largest finite value: a later consumer adds in
// fp16 and must never see an infinity)
__device__ __forceinline__ uint32_t gemm_pack16(float lo, float hi, bool f16) {
    uint32_t r;
    if (f16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

__device__ __forceinline__ float gemm_bias_lane(const float* bias, const float* bias2, int n, int N) {
    float b = 0.f;
    if (n < N) {
        if (bias) b += __ldg(bias + n);
        if (bias2) b += __ldg(bias2 + n);
    }
    return b;
}

__device__ __forceinline__ void gemm_store_chunk(const GemmEpilogue& ep, float (&v)[32], int batch, int m, int nb, int N,
                                                 bool has_bias, float bias_lane, uint32_t dkey, int M) {
    const bool row_ok = m < M;
    const int lane = threadIdx.x & 31;
    if (ep.atomic) {
        if (!row_ok) return;
        float* o = (float*)ep.out + (int64_t)batch * ep.c_sb + (int64_t)m * ep.ldc + nb;
        if (nb + 32 <= N && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {       // 8 vector reductions instead of 32 scalar ones
#pragma unroll
            for (int j = 0; j < 32; j += 4) red_add_f32x4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nb + j < N) atomicAdd(o + j, v[j]);
        }
        return;
    }
    if (has_bias) {                                          // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_lane, j);
    }
    if (ep.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (ep.use_dropout) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            // element index m*N + n; pairs may straddle an odd boundary, so take single multipliers
            const uint64_t idx = (uint64_t)m * N + nb + j;
            if ((idx & 1) == 0) {
                float a, b; dropout_mult2(ep.drop, dkey, idx, a, b); v[j] *= a; v[j + 1] *= b;
            } else {
                v[j] *= dropout_mult(ep.drop, ep.site, idx); v[j + 1] *= dropout_mult(ep.drop, ep.site, idx + 1);
            }
        }
    }
    const bool full_chunk = nb + 32 <= N;
    if (ep.out_bf16) {
        bf16* ob = (bf16*)ep.out + (int64_t)batch * ep.c_sb + nb;
        // warp-uniform: every row's 64-byte chunk is 16-byte aligned
        if (full_chunk && (ep.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(ob) & 15) == 0) {
            // A lane holds one row's 64 bytes as four 16-byte pieces.  4 x 4 transpose inside each lane quad: lane q ends
            // with piece q of the quad's four rows, so one store instruction writes 64 contiguous bytes per row
            // (8 lines per instruction instead of 32 -- the epilogue of the store-heavy GEMMs is LSU-wavefront bound).
            uint32_t G[4][4];
            const bool f16 = ep.out_bf16 == 2;               // warp-uniform
#pragma unroll
            for (int j = 0; j < 16; ++j) G[j >> 2][j & 3] = gemm_pack16(v[2 * j], v[2 * j + 1], f16);
            const int q = lane & 3;
#pragma unroll
            for (int step = 0; step < 2; ++step) {
                const int off = 1 << step;
                const bool up = (q & off) != 0;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int lo = step == 0 ? 2 * i : i, hi = lo + off;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const uint32_t r = __shfl_xor_sync(0xffffffffu, up ? G[lo][t] : G[hi][t], off);
                        if (up) G[lo][t] = r; else G[hi][t] = r;
                    }
                }
            }
            const int mq = m - q;                            // first row of the quad
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (mq + j < M)
                    *reinterpret_cast<uint4*>(ob + (int64_t)(mq + j) * ep.ldc + 8 * q) = make_uint4(G[j][0], G[j][1], G[j][2], G[j][3]);
        } else if (row_ok) {
            bf16* o = ob + (int64_t)m * ep.ldc;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (nb + j < N) {
                    if (ep.out_bf16 == 2) reinterpret_cast<__half*>(o)[j] = __float2half_rn(fminf(fmaxf(v[j], -65504.f), 65504.f));
                    else o[j] = __float2bfloat16_rn(v[j]);
                }
        }
    } else if (row_ok) {
        float* o = (float*)ep.out + (int64_t)batch * ep.c_sb + (int64_t)m * ep.ldc + nb;
        if (full_chunk && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nb + j < N) o[j] = v[j];
        }
    }
}

// Split-K epilogue only (fp32 vector reductions into a zeroed output): the ATOMIC instantiation of gemm_tc_kernel
// carries nothing else -- these kernels run ~10 us (one LSTM backward step), code size is start-up latency.
__device__ __forceinline__ void gemm_red_chunk(const GemmEpilogue& ep, float (&v)[32], int batch, int m, int nb, int N, int M) {
    if (m >= M) return;
    float* o = (float*)ep.out + (int64_t)batch * ep.c_sb + (int64_t)m * ep.ldc + nb;
    if (nb + 32 <= N && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) red_add_f32x4(o + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) if (nb + j < N) atomicAdd(o + j, v[j]);
    }
}

// Lean epilogue of the persistent kernel's FAST instantiation: bf16 output, optional bias, whole 32-column chunks,
// 16-byte aligned rows -- no ReLU / dropout / fp32 / ragged / atomic paths in the instruction stream (the generic
// epilogue's run-time flag tests and dead branches cost the store-heavy GEMMs ~20% in instruction-fetch stalls).
__device__ __forceinline__ void gemm_store_chunk_bf16(const GemmEpilogue& ep, float (&v)[32], int batch, int m, int nb,
                                                      bool has_bias, float bias_lane, int M) {
    const int lane = threadIdx.x & 31;
    if (has_bias) {                                          // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bias_lane, j);
    }
    uint32_t G[4][4];
    if (ep.out_bf16 == 2) {                                  // warp-uniform; two straight-line blocks, no per-element select
#pragma unroll
        for (int j = 0; j < 16; ++j) G[j >> 2][j & 3] = gemm_pack16(v[2 * j], v[2 * j + 1], true);
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) G[j >> 2][j & 3] = gemm_pack16(v[2 * j], v[2 * j + 1], false);
    }
    const int q = lane & 3;
#pragma unroll
    for (int step = 0; step < 2; ++step) {                   // quad transpose, see gemm_store_chunk
        const int off = 1 << step;
        const bool up = (q & off) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int lo = step == 0 ? 2 * i : i, hi = lo + off;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const uint32_t r = __shfl_xor_sync(0xffffffffu, up ? G[lo][t] : G[hi][t], off);
                if (up) G[lo][t] = r; else G[hi][t] = r;
            }
        }
    }
    bf16* ob = (bf16*)ep.out + (int64_t)batch * ep.c_sb + (int64_t)(m - q) * ep.ldc + nb + 8 * q;
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (m - q + j < M) *reinterpret_cast<uint4*>(ob + (int64_t)j * ep.ldc) = make_uint4(G[j][0], G[j][1], G[j][2], G[j][3]);
}

// MAJ = 0: A [M,K], B [N,K] with K contiguous (K-major operands, one TMA box per operand per stage).
// MAJ = 1: A [K,M], B [K,N] with M / N contiguous (MN-major operands: the weight-gradient form dW = dY^T X
//          consumed without transposes; one TMA box per 64-wide column block per stage).
// MAJ = 2: A [M,K] K-major, B [K,N] MN-major: the data-gradient form dX = dY W with W [N_red, K_out] read as stored
//          (no transposed weight copy).
template <int BN, int MAJ, int ST, bool ATOMIC>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               GemmEpilogue ep, int M, int N, int K, int nsplit, int kblocks_per_split, const int* __restrict__ klist) {
    using S = GemmSmem<BN, ST>;
    constexpr bool AMN = MAJ == 1, BMN = MAJ >= 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* sa = smem;
    uint8_t* sb = smem + S::STAGES * S::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::STAGES * (S::A_BYTES + S::B_BYTES));
    uint64_t* empty = full + S::STAGES;
    uint64_t* tmem_full = empty + S::STAGES;
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_full + 1);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int batch = blockIdx.z / nsplit, split = blockIdx.z - batch * nsplit;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    int total_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b);
        for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_smem, BN);      // BN fp32 columns x 128 lanes
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();                                         // operands / output of the previous kernel in the stream

    // Reduction-block list (vqa_tc_gemm_kblocks): klist[0] = n, klist[1..n] = the 64-row blocks of K that are not known to
    // be all zero (written by an earlier kernel of the stream, hence read after pdl_wait).  Every role derives the same
    // [kb_begin, kb_begin + nkb) of the LIST from klist[0]; only the producer looks at the entries.
    if (klist != nullptr) {
        total_kb = min(total_kb, max(0, klist[0]));
        kblocks_per_split = (total_kb + nsplit - 1) / nsplit;
    }
    const int kb_begin = split * kblocks_per_split;
    const int kb_end = min(total_kb, kb_begin + kblocks_per_split);
    const int nkb = max(0, kb_end - kb_begin);

    if (warp == 0) {
        auto load_stage = [&](int i, int kb) {          // lane 0 only
            const int s = i % S::STAGES;
            const uint32_t ph = (i / S::STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            mbar_expect_tx(&full[s], S::A_BYTES + S::B_BYTES);
            const int kc = kb * BK;
            if (!AMN) tma_load_3d(sa + s * S::A_BYTES, &tma_a, &full[s], kc, m0, batch);
            else {
#pragma unroll
                for (int blk = 0; blk < BM / 64; ++blk)
                    tma_load_3d(sa + s * S::A_BYTES + blk * 8192, &tma_a, &full[s], m0 + blk * 64, kc, batch);
            }
            if (!BMN) tma_load_3d(sb + s * S::B_BYTES, &tma_b, &full[s], kc, n0, batch);
            else {
#pragma unroll
                for (int blk = 0; blk < BN / 64; ++blk)
                    tma_load_3d(sb + s * S::B_BYTES + blk * 8192, &tma_b, &full[s], n0 + blk * 64, kc, batch);
            }
        };
        if (klist == nullptr) {
            if (lane == 0)
                for (int i = 0; i < nkb; ++i) load_stage(i, kb_begin + i);
        } else {
            // 32 list entries per coalesced load (the next 32 already in flight), handed to lane 0 by shuffle: a dependent
            // global load per k-block in a one-thread producer would cost more than the MMAs of the block
            const int* e = klist + 1 + kb_begin;
            int cur = lane < nkb ? e[lane] : 0;
            for (int base = 0; base < nkb; base += 32) {
                const int nxt = base + 32 + lane < nkb ? e[base + 32 + lane] : 0;
                const int cnt = min(32, nkb - base);
                for (int j = 0; j < cnt; ++j) {
                    const int kb = __shfl_sync(0xffffffffu, cur, j);
                    if (lane == 0) load_stage(base + j, kb);
                    __syncwarp();
                }
                cur = nxt;
            }
        }
    } else if (warp == 1) {
        // all lanes converged, one elected lane issues (see tc_common.cuh)
        constexpr uint32_t idesc = idesc_bf16(BM, BN, AMN ? 1 : 0, BMN ? 1 : 0);
        const uint32_t elected = elect_one();
        const uint64_t a_desc0 = AMN ? smem_desc_mn_sw128(smem_u32(sa), 8192) : smem_desc_k_sw128(smem_u32(sa));
        const uint64_t b_desc0 = BMN ? smem_desc_mn_sw128(smem_u32(sb), 8192) : smem_desc_k_sw128(smem_u32(sb));
        constexpr uint32_t KSTEP_A = AMN ? (2048 >> 4) : (32 >> 4);   // 16 reduction elements further, in 16-byte units
        constexpr uint32_t KSTEP_B = BMN ? (2048 >> 4) : (32 >> 4);
        for (int i = 0; i < nkb; ++i) {
            const int s = i % S::STAGES;
            mbar_wait(&full[s], (i / S::STAGES) & 1);
            tcgen05_fence_after();
            const uint64_t ad = a_desc0 + (uint64_t)(s * (S::A_BYTES >> 4)), bd = b_desc0 + (uint64_t)(s * (S::B_BYTES >> 4));
#pragma unroll
            for (uint32_t k = 0; k < BK / 16; ++k)
                umma_issue<1>(tmem_base, ad + k * KSTEP_A, bd + k * KSTEP_B, idesc, (i > 0 || k > 0) ? 1u : 0u, elected);
            umma_commit_issue<1>(&empty[s], elected);      // frees the smem stage when these MMAs retire
        }
        umma_commit_issue<1>(tmem_full, elected);          // accumulator complete
    } else {
        // ---- epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (hardware restriction)
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int m = m0 + row;
        if (nkb > 0) { mbar_wait(tmem_full, 0); tcgen05_fence_after(); }
        const float* bias = ep.bias ? ep.bias + (int64_t)batch * ep.bias_sb : nullptr;
        const float* bias2 = ep.bias2 ? ep.bias2 + (int64_t)batch * ep.bias_sb : nullptr;
        const uint32_t dkey = ep.use_dropout ? dropout_key(ep.drop, ep.site) : 0;
        const bool has_bias = !ep.atomic && (bias || bias2);
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= N) break;                     // warp-uniform
            float v[32];
            if (nkb > 0) tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
            }
            if (ATOMIC) gemm_red_chunk(ep, v, batch, m, n0 + c0, N, M);
            else gemm_store_chunk(ep, v, batch, m, n0 + c0, N, has_bias, has_bias ? gemm_bias_lane(bias, bias2, n0 + c0 + lane, N) : 0.f,
                                  dkey, M);
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, BN); }
}

// ------------------------------------------------------------------------------------------ persistent variant
// K-major operands, no split-K: grid = #SMs, output tiles round-robin (n fastest, so concurrently running CTAs share the
// A rows in L2), one continuous TMA ring across tiles, two TMEM accumulators so that the epilogue of tile i (8 warps)
// overlaps the MMAs of tile i+1.  Used when a GEMM has many tiles (attention.v_conv: 10816 tiles of 4 k-blocks, the
// LSTM input projection): per-CTA setup, pipeline fill and the store-heavy epilogue then vanish from the critical path.
constexpr int PGEMM_THREADS = 64 + 8 * 32;
template <int BN>
struct PGemmSmem {
    static constexpr int STAGES = BN == 256 ? 4 : 6;
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;
    static constexpr int BYTES = STAGES * (A_BYTES + B_BYTES) + 1024 + 256;
};

template <int BN, bool BMN, bool FAST>
__global__ void __launch_bounds__(PGEMM_THREADS, 1)
gemm_tc_persistent_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                          GemmEpilogue ep, int M, int N, int K, int nbatch) {
    using S = PGemmSmem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);      // pointer arithmetic keeps the shared address space (LDS/STS, not generic LD/ST)
    uint8_t* sa = smem;
    uint8_t* sb = smem + S::STAGES * S::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::STAGES * (S::A_BYTES + S::B_BYTES));
    uint64_t* empty = full + S::STAGES;
    uint64_t* tmem_full = empty + S::STAGES;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;          // [2]
    uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = (M + BM - 1) / BM, nt = (N + BN - 1) / BN;
    const int ntiles = mt * nt * nbatch;
    const int nkb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a); tma_prefetch_desc(&tma_b);
        for (int i = 0; i < S::STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_base_smem, 2 * BN);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_base_smem;
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int batch = tile / (mt * nt), r = tile - batch * (mt * nt);
                const int m0 = (r / nt) * BM, n0 = (r % nt) * BN;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % S::STAGES;
                    mbar_wait(&empty[s], ((it / S::STAGES) & 1) ^ 1);
                    mbar_expect_tx(&full[s], S::A_BYTES + S::B_BYTES);
                    tma_load_3d(sa + s * S::A_BYTES, &tma_a, &full[s], kb * BK, m0, batch);
                    if (!BMN) tma_load_3d(sb + s * S::B_BYTES, &tma_b, &full[s], kb * BK, n0, batch);
                    else {
#pragma unroll
                        for (int blk = 0; blk < BN / 64; ++blk)
                            tma_load_3d(sb + s * S::B_BYTES + blk * 8192, &tma_b, &full[s], n0 + blk * 64, kb * BK, batch);
                    }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = idesc_bf16(BM, BN, 0, BMN ? 1 : 0);
        constexpr uint32_t KSTEP_B = BMN ? (2048 >> 4) : (32 >> 4);
        const uint32_t elected = elect_one();
        const uint64_t a_desc0 = smem_desc_k_sw128(smem_u32(sa));
        const uint64_t b_desc0 = BMN ? smem_desc_mn_sw128(smem_u32(sb), 8192) : smem_desc_k_sw128(smem_u32(sb));
        uint32_t it = 0, tcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const uint32_t acc = tcount & 1, use = tcount >> 1;
            mbar_wait(&tmem_empty[acc], (use & 1) ^ 1);
            tcgen05_fence_after();
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const int s = it % S::STAGES;
                mbar_wait(&full[s], (it / S::STAGES) & 1);
                tcgen05_fence_after();
                const uint64_t ad = a_desc0 + (uint64_t)(s * (S::A_BYTES >> 4)), bd = b_desc0 + (uint64_t)(s * (S::B_BYTES >> 4));
#pragma unroll
                for (uint32_t k = 0; k < BK / 16; ++k)
                    umma_issue<1>(tmem_base + acc * BN, ad + 2 * k, bd + k * KSTEP_B, idesc, (kb > 0 || k > 0) ? 1u : 0u, elected);
                umma_commit_issue<1>(&empty[s], elected);
            }
            umma_commit_issue<1>(&tmem_full[acc], elected);
        }
    } else {
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t dkey = ep.use_dropout ? dropout_key(ep.drop, ep.site) : 0;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const uint32_t acc = tcount & 1, use = tcount >> 1;
            const int batch = tile / (mt * nt), r = tile - batch * (mt * nt);
            const int m0 = (r / nt) * BM, n0 = (r % nt) * BN;
            const int m = m0 + quarter * 32 + lane;
            const float* bias = ep.bias ? ep.bias + (int64_t)batch * ep.bias_sb : nullptr;
            const float* bias2 = ep.bias2 ? ep.bias2 + (int64_t)batch * ep.bias_sb : nullptr;
            const bool has_bias = bias || bias2;
            float bl[BN / 64];                               // this lane's column bias of each of the warp's chunks
#pragma unroll
            for (int k = 0; k < BN / 64; ++k) bl[k] = has_bias ? gemm_bias_lane(bias, bias2, n0 + half * 32 + 64 * k + lane, N) : 0.f;
            mbar_wait(&tmem_full[acc], use & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int k = 0; k < BN / 64; ++k) {
                const int c0 = half * 32 + 64 * k;
                if (n0 + c0 < N) {                           // warp-uniform
                    float v[32];
                    tmem_ld_32x32(tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
                    if (FAST) gemm_store_chunk_bf16(ep, v, batch, m, n0 + c0, has_bias, bl[k], M);
                    else gemm_store_chunk(ep, v, batch, m, n0 + c0, N, has_bias, bl[k], dkey, M);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) { tcgen05_fence_after(); tmem_dealloc(tmem_base, 2 * BN); }
}

template <int BN, bool BMN, bool FAST>
static int launch_gemm_persistent_f(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                                    int nbatch, int sms, cudaStream_t st) {
    auto kern = gemm_tc_persistent_kernel<BN, BMN, FAST>;
    static bool attr_set = false;
    if (!attr_set) {
        VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PGemmSmem<BN>::BYTES));
        attr_set = true;
    }
    VQA_CUDA(vqa_launch_pdl(kern, dim3(sms), dim3(PGEMM_THREADS), PGemmSmem<BN>::BYTES, st, ta, tb, ep, M, N, K, nbatch));
    VQA_CHECK_LAUNCH("gemm_tc_persistent");
    return 0;
}

template <int BN, bool BMN>
static int launch_gemm_persistent(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                                  int nbatch, int sms, cudaStream_t st) {
    const bool fast = ep.out_bf16 && !ep.relu && !ep.use_dropout && !ep.atomic && N % 32 == 0 && (ep.ldc & 7) == 0 &&
                      (ep.c_sb & 7) == 0 && ((uintptr_t)ep.out & 15) == 0;
    if (fast) return launch_gemm_persistent_f<BN, BMN, true>(ta, tb, ep, M, N, K, nbatch, sms, st);
    return launch_gemm_persistent_f<BN, BMN, false>(ta, tb, ep, M, N, K, nbatch, sms, st);
}

template <int BN, int MAJ, int ST, bool ATOMIC>
static int launch_gemm_sta(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                           dim3 grid, int nsplit, int kbps, const int* klist, cudaStream_t st) {
    auto kern = gemm_tc_kernel<BN, MAJ, ST, ATOMIC>;
    static bool attr_set = false;
    if (!attr_set) {
        VQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<BN, ST>::BYTES));
        attr_set = true;
    }
    VQA_CUDA(vqa_launch_pdl(kern, grid, dim3(GEMM_THREADS), GemmSmem<BN, ST>::BYTES, st, ta, tb, ep, M, N, K, nsplit, kbps, klist));
    VQA_CHECK_LAUNCH("gemm_tc");
    return 0;
}

template <int BN, int MAJ, int ST>
static int launch_gemm_st(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                          dim3 grid, int nsplit, int kbps, const int* klist, cudaStream_t st) {
    if (ep.atomic) return launch_gemm_sta<BN, MAJ, ST, true>(ta, tb, ep, M, N, K, grid, nsplit, kbps, klist, st);
    return launch_gemm_sta<BN, MAJ, ST, false>(ta, tb, ep, M, N, K, grid, nsplit, kbps, klist, st);
}

template <int BN, int MAJ>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpilogue& ep, int M, int N, int K,
                       int nbatch, int nsplit, const int* klist, cudaStream_t st) {
    const int total_kb = (K + BK - 1) / BK;
    if (nsplit < 1) nsplit = 1;
    if (nsplit > total_kb) nsplit = total_kb > 0 ? total_kb : 1;
    int kbps = total_kb > 0 ? (total_kb + nsplit - 1) / nsplit : 1;
    nsplit = total_kb > 0 ? (total_kb + kbps - 1) / kbps : 1;
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, nbatch * nsplit);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ctas = (int64_t)grid.x * grid.y * grid.z;
    if (MAJ != 1 && nsplit == 1 && !ep.atomic && total_kb > 0 && ctas >= 4 * (int64_t)sms)
        return launch_gemm_persistent<BN, MAJ == 2>(ta, tb, ep, M, N, K, nbatch, sms, st);     // (klist: MAJ == 1 only)
    if (ctas >= 2 * (int64_t)sms) return launch_gemm_st<BN, MAJ, 3>(ta, tb, ep, M, N, K, grid, nsplit, kbps, klist, st);
    return launch_gemm_st<BN, MAJ, 6>(ta, tb, ep, M, N, K, grid, nsplit, kbps, klist, st);
}

}  // namespace tc

using namespace tc;

// C[z][m,n] = act(sum_k A[z][m,k] B[z][n,k] + bias) ; A [M,K] (row pitch lda), B [N,K] (row pitch ldb), bf16;
// with VQA_GEMM_OPERANDS_MN: A stored [K,M] (row pitch lda), B stored [K,N] (row pitch ldb).
static int tc_gemm_impl(const void* A, int64_t lda, int64_t a_sb, const void* B, int64_t ldb, int64_t b_sb,
                        void* C, int c_dtype, int64_t ldc, int64_t c_sb,
                        const float* bias, const float* bias2, int64_t bias_sb,
                        int M, int N, int K, int nbatch, int flags,
                        float p_drop, uint64_t seed, uint32_t site, const int* klist, void* stream) {
    VQA_REQUIRE(M > 0 && N > 0 && K > 0 && nbatch >= 1, "tc_gemm: bad dims M=%d N=%d K=%d nbatch=%d", M, N, K, nbatch);
    VQA_REQUIRE(A && B && C, "tc_gemm: null operand");
    VQA_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "tc_gemm: row pitches (%lld, %lld) must be multiples of 8 bf16 elements (TMA 16-byte strides)",
                (long long)lda, (long long)ldb);
    VQA_REQUIRE(nbatch == 1 || (a_sb % 8 == 0 && b_sb % 8 == 0), "tc_gemm: batch strides must be multiples of 8 elements");
    VQA_REQUIRE(c_dtype == VQA_F32 || c_dtype == VQA_BF16 || c_dtype == VQA_F16, "tc_gemm: bad output dtype");
    const bool splitk = (flags & VQA_GEMM_SPLITK) != 0;
    const bool mn = (flags & VQA_GEMM_OPERANDS_MN) != 0;
    const bool bmn = (flags & VQA_GEMM_B_MN) != 0;          // A [M,K] K-major, B stored [K,N]
    VQA_REQUIRE(!(mn && bmn), "tc_gemm: OPERANDS_MN and B_MN are mutually exclusive");
    VQA_REQUIRE(klist == nullptr || (mn && nbatch == 1 && ((uintptr_t)klist & 3) == 0),
                "tc_gemm_kblocks: a reduction-block list needs VQA_GEMM_OPERANDS_MN, nbatch == 1 and a 4-byte aligned list");
    if (splitk) {
        VQA_REQUIRE(c_dtype == VQA_F32 && !bias && !bias2 && !(flags & VQA_GEMM_RELU) && p_drop == 0.f,
                    "tc_gemm: split-K needs a zeroed fp32 output and no fused bias/relu/dropout");
    }
    VQA_REQUIRE(!(flags & VQA_GEMM_ACCUMULATE), "tc_gemm: ACCUMULATE is not supported (use SPLITK into a zeroed buffer)");
    cudaStream_t st = (cudaStream_t)stream;

    // 128-wide tiles unless that leaves most SMs idle (e.g. the per-step LSTM data gradient: M = 256): then 64-wide
    int BN = N <= 64 ? 64 : 128;
    if (BN == 128 && !splitk) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t tiles128 = (int64_t)((M + BM - 1) / BM) * ((N + 127) / 128) * nbatch;
        if (tiles128 * 2 <= sms) BN = 64;
    }
    // 256-wide tiles for the big store-heavy GEMMs (attention.v_conv and its data gradient, the LSTM input projection):
    // with N = 128 the MMA's operand reads alone saturate the 128 B/clk of shared memory; at N = 256 the A tile is read
    // once per 256 output columns.  Persistent kernel only (2 x 256 TMEM columns, 4 stages of 48 KB).  (Keeping the
    // [256 x K] weight slice resident in shared memory instead of re-streaming it from L2 per tile measured SLOWER:
    // v_conv 0.140 vs 0.126 ms -- the L2 -> SM traffic of B is not what bounds these GEMMs.)
    if (BN == 128 && !splitk && !mn && N % 256 == 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t tiles256 = (int64_t)((M + BM - 1) / BM) * (N / 256) * nbatch;
        if (tiles256 >= 4 * (int64_t)sms) BN = 256;
    }
    CUtensorMap ta, tb;
    if (!mn) {
        const uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, (uint64_t)nbatch};
        const uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)(nbatch > 1 ? a_sb : (int64_t)M * lda) * 2};
        const uint32_t box[3] = {BK, BM, 1};
        if (int e = make_tmap_bf16(&ta, A, 3, dims, str, box)) return e;
        if (!bmn) {
            const uint64_t dimsb[3] = {(uint64_t)K, (uint64_t)N, (uint64_t)nbatch};
            const uint64_t strb[2] = {(uint64_t)ldb * 2, (uint64_t)(nbatch > 1 ? b_sb : (int64_t)N * ldb) * 2};
            const uint32_t boxb[3] = {BK, (uint32_t)BN, 1};
            if (int e = make_tmap_bf16(&tb, B, 3, dimsb, strb, boxb)) return e;
        } else {
            const uint64_t dimsb[3] = {(uint64_t)N, (uint64_t)K, (uint64_t)nbatch};
            const uint64_t strb[2] = {(uint64_t)ldb * 2, (uint64_t)(nbatch > 1 ? b_sb : (int64_t)K * ldb) * 2};
            const uint32_t boxb[3] = {64, BK, 1};
            if (int e = make_tmap_bf16(&tb, B, 3, dimsb, strb, boxb)) return e;
        }
    } else {
        const uint64_t dims[3] = {(uint64_t)M, (uint64_t)K, (uint64_t)nbatch};
        const uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)(nbatch > 1 ? a_sb : (int64_t)K * lda) * 2};
        const uint32_t box[3] = {64, BK, 1};
        if (int e = make_tmap_bf16(&ta, A, 3, dims, str, box)) return e;
        const uint64_t dimsb[3] = {(uint64_t)N, (uint64_t)K, (uint64_t)nbatch};
        const uint64_t strb[2] = {(uint64_t)ldb * 2, (uint64_t)(nbatch > 1 ? b_sb : (int64_t)K * ldb) * 2};
        if (int e = make_tmap_bf16(&tb, B, 3, dimsb, strb, box)) return e;
    }
    GemmEpilogue ep{};
    ep.out = C; ep.out_bf16 = c_dtype == VQA_BF16 ? 1 : (c_dtype == VQA_F16 ? 2 : 0); ep.ldc = ldc; ep.c_sb = c_sb;
    ep.bias = bias; ep.bias2 = bias2; ep.bias_sb = bias_sb;
    ep.relu = (flags & VQA_GEMM_RELU) ? 1 : 0;
    ep.atomic = splitk ? 1 : 0;
    ep.use_dropout = p_drop > 0.f;
    ep.site = site; ep.drop = make_dropout(seed, p_drop);

    int nsplit = 1;
    if (splitk) {
        const int64_t tiles = (int64_t)((M + BM - 1) / BM) * ((N + BN - 1) / BN) * nbatch;
        nsplit = (int)(148 / tiles);                        // about one CTA per SM: every extra split adds a full tile of atomics
        const int maxs = ((K + BK - 1) / BK + 3) / 4;       // at least 4 k-blocks per split
        if (nsplit > maxs) nsplit = maxs;
        if (tiles >= 148) nsplit = 1;                       // enough tiles to fill the GPU: no split, no atomics
        if (nsplit < 1) nsplit = 1;
        if (nsplit == 1) ep.atomic = 0;                     // the output is zeroed, a plain store is equivalent
    }
    if (BN == 256) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (bmn) return launch_gemm_persistent<256, true>(ta, tb, ep, M, N, K, nbatch, sms, st);
        return launch_gemm_persistent<256, false>(ta, tb, ep, M, N, K, nbatch, sms, st);
    }
    if (mn) {
        if (BN == 64) return launch_gemm<64, 1>(ta, tb, ep, M, N, K, nbatch, nsplit, klist, st);
        return launch_gemm<128, 1>(ta, tb, ep, M, N, K, nbatch, nsplit, klist, st);
    }
    if (bmn) {
        if (BN == 64) return launch_gemm<64, 2>(ta, tb, ep, M, N, K, nbatch, nsplit, nullptr, st);
        return launch_gemm<128, 2>(ta, tb, ep, M, N, K, nbatch, nsplit, nullptr, st);
    }
    if (BN == 64) return launch_gemm<64, 0>(ta, tb, ep, M, N, K, nbatch, nsplit, nullptr, st);
    return launch_gemm<128, 0>(ta, tb, ep, M, N, K, nbatch, nsplit, nullptr, st);
}

extern "C" int vqa_tc_gemm(const void* A, int64_t lda, int64_t a_sb, const void* B, int64_t ldb, int64_t b_sb,
                           void* C, int c_dtype, int64_t ldc, int64_t c_sb,
                           const float* bias, const float* bias2, int64_t bias_sb,
                           int M, int N, int K, int nbatch, int flags,
                           float p_drop, uint64_t seed, uint32_t site, void* stream) {
    return tc_gemm_impl(A, lda, a_sb, B, ldb, b_sb, C, c_dtype, ldc, c_sb, bias, bias2, bias_sb, M, N, K, nbatch, flags,
                        p_drop, seed, site, nullptr, stream);
}

// The reduction-major form (VQA_GEMM_OPERANDS_MN, nbatch = 1) restricted to the 64-row reduction blocks named by a
// DEVICE-side list: kblocks[0] = n, kblocks[1..n] = block indices (block j = reduction rows 64j .. 64j+63), each at most
// once.  Blocks that are not listed are not read and contribute nothing -- the caller guarantees that one operand is all
// zero there (the LSTM gate gradients of (step, row) pairs past the end of their question: vqa_lstm_active_kblocks).
// K is still the dense row count (tensor-map bound).  With split-K the LIST is what is divided among the splits.
extern "C" int vqa_tc_gemm_kblocks(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                                   int M, int N, int K, int flags, const int32_t* kblocks, void* stream) {
    VQA_REQUIRE(kblocks != nullptr, "tc_gemm_kblocks: null block list");
    VQA_REQUIRE((flags & ~(VQA_GEMM_OPERANDS_MN | VQA_GEMM_SPLITK)) == 0 && (flags & VQA_GEMM_OPERANDS_MN),
                "tc_gemm_kblocks: flags must be VQA_GEMM_OPERANDS_MN, optionally with VQA_GEMM_SPLITK");
    return tc_gemm_impl(A, lda, 0, B, ldb, 0, C, VQA_F32, ldc, 0, nullptr, nullptr, 0, M, N, K, 1, flags, 0.f, 0, 0,
                        kblocks, stream);
}

// ------------------------------------------------------------------------------------------
// dst[c, r] (bf16, row pitch ldd) = (bf16) src[r, c]   -- tiled transpose with optional fp32 -> bf16 cast.
// rows x cols source with pitch lds; batched over z with strides.
// ------------------------------------------------------------------------------------------
template <typename TS>
__global__ void transpose_bf16_kernel(const TS* __restrict__ src, int64_t lds, int64_t s_sb, bf16* __restrict__ dst,
                                      int64_t ldd, int64_t d_sb, int rows, int cols) {
    pdl_trigger();
    pdl_wait();
    __shared__ float tile[32][33];
    src += (int64_t)blockIdx.z * s_sb; dst += (int64_t)blockIdx.z * d_sb;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? to_f32(src[(int64_t)r * lds + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) dst[(int64_t)c * ldd + r] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    }
}

extern "C" int vqa_transpose_bf16(const void* src, int src_dtype, int64_t lds, int64_t s_sb, void* dst, int64_t ldd,
                                  int64_t d_sb, int rows, int cols, int nbatch, void* stream) {
    VQA_REQUIRE(rows > 0 && cols > 0 && nbatch >= 1 && lds >= cols && ldd >= rows, "transpose: bad dims");
    dim3 grid((cols + 31) / 32, (rows + 31) / 32, nbatch), block(32, 8);
    VQA_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "transpose: too many rows/batches for one launch");
    cudaStream_t st = (cudaStream_t)stream;
    if (src_dtype == VQA_F32) VQA_CUDA(vqa_launch_pdl(transpose_bf16_kernel<float>, grid, block, 0, st, (const float*)src, lds, s_sb, (bf16*)dst, ldd, d_sb, rows, cols));
    else if (src_dtype == VQA_BF16) VQA_CUDA(vqa_launch_pdl(transpose_bf16_kernel<bf16>, grid, block, 0, st, (const bf16*)src, lds, s_sb, (bf16*)dst, ldd, d_sb, rows, cols));
    else VQA_REQUIRE(false, "transpose: bad dtype");
    VQA_CHECK_LAUNCH("transpose_bf16");
    return 0;
}
