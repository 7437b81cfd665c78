// Image encoder, generic-shape / exact-fp32 arm: implicit-GEMM conv + bias + ReLU + 2x2 max-pool in one
// kernel (models/model.py:72-84), and its backward.  See gemm_simt.cuh for the tiling.
#include "gemm_simt.cuh"

using namespace simt;

static int conv_geom(ConvGeom& g, int B, int IH, int IW, int Cin, int Cout, int KS, int stride) {
    VQA_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && KS > 0 && stride > 0, "conv: bad dims");
    VQA_REQUIRE(IH >= KS && IW >= KS, "conv: input %dx%d smaller than kernel %d", IH, IW, KS);
    const int OH = (IH - KS) / stride + 1, OW = (IW - KS) / stride + 1;
    g.B = B; g.IH = IH; g.IW = IW; g.Cin = Cin; g.Cout = Cout; g.KS = KS; g.stride = stride;
    g.PH = OH / 2; g.PW = OW / 2;
    VQA_REQUIRE(g.PH > 0 && g.PW > 0, "conv: pooled output is empty (%dx%d conv output)", OH, OW);
    VQA_REQUIRE((int64_t)B * g.PH * g.PW * 4 < (1ll << 31), "conv: too many output positions for int32 indexing");
    return 0;
}

template <typename TIn, bool NCHW, typename TOut>
static int conv_fwd_t(const void* x, const float* w, const float* bias, void* out, uint8_t* mask,
                      const ConvGeom& g, cudaStream_t st) {
    const int M = g.B * g.PH * g.PW * 4, K = g.KS * g.KS * g.Cin;
    ConvFwdALoader<TIn, NCHW> al{(const TIn*)x, g, M, K};
    ConvWLoader bl{w, g.Cout, g.Cin, g.KS * g.KS, K};
    EpPool<TOut> ep{(TOut*)out, mask, bias};
    return launch(al, bl, ep, M, g.Cout, K, 1, 1, st, "conv_relu_pool_fwd");
}

extern "C" int vqa_conv_relu_pool_fwd(const void* x, int x_dtype, int x_nchw, const float* w, const float* bias,
                                      void* out, uint8_t* mask, int act_dtype,
                                      int B, int IH, int IW, int Cin, int Cout, int KS, int stride, void* stream) {
    ConvGeom g;
    if (int e = conv_geom(g, B, IH, IW, Cin, Cout, KS, stride)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    const int key = (x_dtype << 2) | ((x_nchw ? 1 : 0) << 1) | act_dtype;
    switch (key) {
        case (VQA_F32 << 2) | 2 | VQA_F32:  return conv_fwd_t<float, true, float>(x, w, bias, out, mask, g, st);
        case (VQA_F32 << 2) | 2 | VQA_BF16: return conv_fwd_t<float, true, bf16>(x, w, bias, out, mask, g, st);
        case (VQA_F32 << 2) | 0 | VQA_F32:  return conv_fwd_t<float, false, float>(x, w, bias, out, mask, g, st);
        case (VQA_BF16 << 2) | 0 | VQA_BF16: return conv_fwd_t<bf16, false, bf16>(x, w, bias, out, mask, g, st);
        default: VQA_REQUIRE(false, "conv fwd: unsupported dtype/layout combination (x_dtype=%d nchw=%d act=%d)",
                             x_dtype, x_nchw, act_dtype);
    }
    return 0;
}

template <typename T>
static int conv_dgrad_t(const void* dpool, const uint8_t* mask, const float* w, void* dx, const ConvGeom& g,
                        cudaStream_t st) {
    const int M = g.B * g.IH * g.IW, K = g.KS * g.KS * g.Cout;
    ConvDgradALoader<T> al{{(const T*)dpool, mask, g}, M, K};
    ConvDgradWLoader bl{w, g.Cout, g.Cin, g.KS * g.KS, K};
    EpStore<T> ep{};
    ep.out = (T*)dx; ep.ldc = g.Cin;
    return launch(al, bl, ep, M, g.Cin, K, 1, 1, st, "conv_bwd_data");
}

extern "C" int vqa_conv_bwd_data(const void* dpool, const uint8_t* mask, const float* w, void* dx, int act_dtype,
                                 int B, int IH, int IW, int Cin, int Cout, int KS, int stride, void* stream) {
    ConvGeom g;
    if (int e = conv_geom(g, B, IH, IW, Cin, Cout, KS, stride)) return e;
    VQA_REQUIRE((int64_t)B * IH * IW < (1ll << 31), "conv dgrad: too many input positions");
    cudaStream_t st = (cudaStream_t)stream;
    if (act_dtype == VQA_F32) return conv_dgrad_t<float>(dpool, mask, w, dx, g, st);
    if (act_dtype == VQA_BF16) return conv_dgrad_t<bf16>(dpool, mask, w, dx, g, st);
    VQA_REQUIRE(false, "conv dgrad: bad dtype %d", act_dtype);
    return 0;
}

template <typename TIn, bool NCHW, typename T>
static int conv_wgrad_t(const void* x, const void* dpool, const uint8_t* mask, float* dw, const ConvGeom& g,
                        cudaStream_t st) {
    const int N = g.KS * g.KS * g.Cin, K = g.B * g.PH * g.PW * 4;
    ConvWgradALoader<T> al{(const T*)dpool, mask, g.Cout, K};
    ConvWgradBLoader<TIn, NCHW> bl{(const TIn*)x, g, N, K};
    EpAtomic ep{dw, 0, 0, 1, g.Cin, g.KS * g.KS};
    const int split = pick_split(g.Cout, N, K, 1);
    return launch(al, bl, ep, g.Cout, N, K, 1, split, st, "conv_bwd_weight");
}

int vqa_colsum_impl(const void* in, int dtype, int64_t ld, const uint8_t* mask, float* out, int64_t rows, int cols,
                    cudaStream_t st);

extern "C" int vqa_conv_bwd_weight(const void* x, int x_dtype, int x_nchw, const void* dpool, const uint8_t* mask,
                                   float* dw, float* db, int act_dtype,
                                   int B, int IH, int IW, int Cin, int Cout, int KS, int stride, void* stream) {
    ConvGeom g;
    if (int e = conv_geom(g, B, IH, IW, Cin, Cout, KS, stride)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    VQA_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Cout * Cin * KS * KS, st));
    int e;
    const int key = (x_dtype << 2) | ((x_nchw ? 1 : 0) << 1) | act_dtype;
    switch (key) {
        case (VQA_F32 << 2) | 2 | VQA_F32:  e = conv_wgrad_t<float, true, float>(x, dpool, mask, dw, g, st); break;
        case (VQA_F32 << 2) | 2 | VQA_BF16: e = conv_wgrad_t<float, true, bf16>(x, dpool, mask, dw, g, st); break;
        case (VQA_F32 << 2) | 0 | VQA_F32:  e = conv_wgrad_t<float, false, float>(x, dpool, mask, dw, g, st); break;
        case (VQA_BF16 << 2) | 0 | VQA_BF16: e = conv_wgrad_t<bf16, false, bf16>(x, dpool, mask, dw, g, st); break;
        default: VQA_REQUIRE(false, "conv wgrad: unsupported dtype/layout combination"); e = 0;
    }
    if (e) return e;
    if (db) {
        VQA_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * (size_t)Cout, st));
        return vqa_colsum_impl(dpool, act_dtype, Cout, mask, db, (int64_t)B * g.PH * g.PW, Cout, st);
    }
    return 0;
}
