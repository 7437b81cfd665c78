// Dense contractions (nn.Linear, 1x1 conv, LSTM projections and their autograd) on the fp32 SIMT core,
// plus the LSTM time step with the cell fused in the epilogue.
#include "gemm_simt.cuh"

using namespace simt;

struct GemmArgs {
    const void* A; int64_t a_sr, a_sk, a_sb;
    const void* B; int64_t b_sr, b_sk, b_sb;
    void* C; int64_t ldc, c_sb;
    const float* bias; const float* bias2; int64_t bias_sb;
    int M, N, K, nbatch, flags;
    float p; uint64_t seed; uint32_t site;
    cudaStream_t st;
};

template <typename TA, typename TB, typename TC>
static int gemm_t(const GemmArgs& a) {
    DenseLoader<TA> al{(const TA*)a.A, a.M, a.K, a.a_sr, a.a_sk, a.a_sb, 0};
    DenseLoader<TB> bl{(const TB*)a.B, a.N, a.K, a.b_sr, a.b_sk, a.b_sb, 0};
    EpStore<TC> ep{};
    ep.out = (TC*)a.C; ep.ldc = a.ldc; ep.sbatch = a.c_sb;
    ep.bias = a.bias; ep.bias2 = a.bias2; ep.bias_sbatch = a.bias_sb;
    ep.relu = (a.flags & VQA_GEMM_RELU) ? 1 : 0;
    ep.accumulate = (a.flags & VQA_GEMM_ACCUMULATE) ? 1 : 0;
    ep.use_dropout = a.p > 0.f;
    ep.site = a.site;
    ep.drop = make_dropout(a.seed, a.p);
    return launch(al, bl, ep, a.M, a.N, a.K, a.nbatch, 1, a.st, "gemm");
}

template <typename TA, typename TB>
static int gemm_splitk_t(const GemmArgs& a) {
    DenseLoader<TA> al{(const TA*)a.A, a.M, a.K, a.a_sr, a.a_sk, a.a_sb, 0};
    DenseLoader<TB> bl{(const TB*)a.B, a.N, a.K, a.b_sr, a.b_sk, a.b_sb, 0};
    EpAtomic ep{(float*)a.C, a.ldc, a.c_sb, 0, 0, 0};
    return launch(al, bl, ep, a.M, a.N, a.K, a.nbatch, pick_split(a.M, a.N, a.K, a.nbatch), a.st, "gemm_splitk");
}

extern "C" int vqa_gemm(const void* A, int a_dtype, int64_t a_sr, int64_t a_sk, int64_t a_sb,
                        const void* B, int b_dtype, int64_t b_sr, int64_t b_sk, int64_t b_sb,
                        void* C, int c_dtype, int64_t ldc, int64_t c_sb,
                        const float* bias, const float* bias2, int64_t bias_sb,
                        int M, int N, int K, int nbatch, int flags,
                        float p_drop, uint64_t seed, uint32_t site, void* stream) {
    VQA_REQUIRE(M >= 0 && N >= 0 && K >= 0 && nbatch >= 1, "gemm: bad dims M=%d N=%d K=%d nbatch=%d", M, N, K, nbatch);
    VQA_REQUIRE(A && B && C, "gemm: null operand");
    GemmArgs a{A, a_sr, a_sk, a_sb, B, b_sr, b_sk, b_sb, C, ldc, c_sb, bias, bias2, bias_sb,
               M, N, K, nbatch, flags, p_drop, seed, site, (cudaStream_t)stream};
    const int key = a_dtype * 4 + b_dtype * 2 + c_dtype;
    if (flags & VQA_GEMM_SPLITK) {
        VQA_REQUIRE(c_dtype == VQA_F32, "gemm: split-K needs an fp32 output");
        VQA_REQUIRE(!bias && !bias2 && !(flags & VQA_GEMM_RELU) && p_drop == 0.f,
                    "gemm: split-K cannot fuse bias/relu/dropout");
        switch (key >> 1) {
            case 0: return gemm_splitk_t<float, float>(a);
            case 1: return gemm_splitk_t<float, bf16>(a);
            case 2: return gemm_splitk_t<bf16, float>(a);
            case 3: return gemm_splitk_t<bf16, bf16>(a);
        }
    }
    switch (key) {
        case 0: return gemm_t<float, float, float>(a);
        case 1: return gemm_t<float, float, bf16>(a);
        case 2: return gemm_t<float, bf16, float>(a);
        case 3: return gemm_t<float, bf16, bf16>(a);
        case 4: return gemm_t<bf16, float, float>(a);
        case 5: return gemm_t<bf16, float, bf16>(a);
        case 6: return gemm_t<bf16, bf16, float>(a);
        case 7: return gemm_t<bf16, bf16, bf16>(a);
    }
    VQA_REQUIRE(false, "gemm: bad dtype codes %d %d %d", a_dtype, b_dtype, c_dtype);
    return 0;
}

// LSTM forward step: A = h_{s-1} [B,H] (row-major), B = W_hh gate-interleaved view, epilogue = cell
template <typename T>
static int lstm_step_t(void* gx, float* cs, void* hs, void* qf, const float* w_hh, int64_t w_sb,
                       const int64_t* q_len, int s, int T_, int B, int H, int dirs, cudaStream_t st) {
    // h_{s-1} for direction z lives at hs + ((z*T + s-1)*B)*H  -> batch stride T*B*H
    const T* hprev = s > 0 ? (const T*)hs + (int64_t)(s - 1) * B * H : (const T*)hs;
    DenseLoader<T> al{hprev, B, s > 0 ? H : 0, H, 1, (int64_t)T_ * B * H, 0};
    DenseLoader<float> bl{w_hh, 4 * H, s > 0 ? H : 0, H, 1, w_sb, H};
    EpLstmCell<T> ep{(T*)gx, cs, (T*)hs, (T*)qf, q_len, s, T_, B, H, dirs, 0};
    return launch(al, bl, ep, B, 4 * H, s > 0 ? H : 0, dirs, 1, st, "lstm_step_fwd");
}

extern "C" int vqa_lstm_step_fwd(void* gx, float* cs, void* hs, void* qf, const float* w_hh, int64_t w_hh_dir_stride,
                                 const int64_t* q_len, int act_dtype, int s, int T, int B, int H, int dirs,
                                 void* stream) {
    VQA_REQUIRE(s >= 0 && s < T && B > 0 && H > 0 && (dirs == 1 || dirs == 2), "lstm step: bad dims");
    cudaStream_t st = (cudaStream_t)stream;
    if (act_dtype == VQA_F32) return lstm_step_t<float>(gx, cs, hs, qf, w_hh, w_hh_dir_stride, q_len, s, T, B, H, dirs, st);
    if (act_dtype == VQA_BF16) return lstm_step_t<bf16>(gx, cs, hs, qf, w_hh, w_hh_dir_stride, q_len, s, T, B, H, dirs, st);
    VQA_REQUIRE(false, "lstm step: bad dtype");
    return 0;
}
