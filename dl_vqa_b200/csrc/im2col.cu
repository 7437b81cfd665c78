// General convolutions of the image encoder on the tensor-core arm (models/model.py:72-84 with ANY cfg stride /
// kernel size / channel list, e.g. the stride-2 encoder of config/config_eval.yaml:52-62): the layers the direct
// tcgen05 kernels (conv0_tc.cu, conv_tc.cu: 3x3, stride 1, 64-multiples) do not cover run as
//     im2col  ->  vqa_tc_gemm (tcgen05 / TMEM / TMA; bias + ReLU in its epilogue)  ->  2x2 max-pool + arg-max mask
// and, backward,
//     un-pool  ->  vqa_tc_gemm (weight gradient, reduction-major operands)  and  vqa_tc_gemm (data gradient)  ->  col2im.
// Everything in this file is the memory-bound glue around those GEMMs: 128-bit accesses along the channel dimension
// (NHWC, C % 8 == 0), one thread per 8 channels.  Column order of the patch matrix: k = (kh * KS + kw) * Cin + ci,
// zero padded to Kp (a multiple of 8, the GEMM's pitch rule); weights are re-packed to the same order.
#include "common.cuh"

namespace {

__device__ __forceinline__ float in_val(const float* x, int64_t i) { return x[i]; }
__device__ __forceinline__ float in_val(const __half* x, int64_t i) { return __half2float(x[i]); }
__device__ __forceinline__ float in_val(const bf16* x, int64_t i) { return __bfloat162float(x[i]); }

// ---- patch matrix, scalar form (any layout / channel count; used for the first layer: NCHW, Cin = 3)
template <typename T>
__global__ void im2col_scalar_kernel(const T* __restrict__ x, int nchw, bf16* __restrict__ col, int64_t total, int IH, int IW,
                                     int Cin, int KS, int stride, int OH, int OW, int Kp) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over M * Kp
    if (i >= total) return;
    const int k = (int)(i % Kp);
    const int64_t m = i / Kp;
    float v = 0.f;
    if (k < KS * KS * Cin) {
        const int ci = k % Cin, tap = k / Cin, kh = tap / KS, kw = tap - kh * KS;
        const int ow = (int)(m % OW), oh = (int)((m / OW) % OH);
        const int64_t b = m / ((int64_t)OW * OH);
        const int ih = oh * stride + kh, iw = ow * stride + kw;
        v = nchw ? in_val(x, ((b * Cin + ci) * IH + ih) * IW + iw) : in_val(x, ((b * IH + ih) * IW + iw) * Cin + ci);
    }
    col[i] = __float2bfloat16_rn(v);
}

// ---- patch matrix, vector form: NHWC bf16 input with Cin % 8 == 0 (then Kp == KS*KS*Cin)
__global__ void im2col_vec8_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int64_t total, int IH, int IW, int Cin,
                                   int KS, int stride, int OH, int OW) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over M * KS*KS * Cin/8
    if (i >= total) return;
    const int c8n = Cin >> 3;
    const int c8 = (int)(i % c8n);
    const int tap = (int)((i / c8n) % (KS * KS));
    const int64_t m = i / ((int64_t)c8n * KS * KS);
    const int kh = tap / KS, kw = tap - kh * KS;
    const int ow = (int)(m % OW), oh = (int)((m / OW) % OH);
    const int64_t b = m / ((int64_t)OW * OH);
    const uint4 v = *reinterpret_cast<const uint4*>(x + ((b * IH + oh * stride + kh) * IW + ow * stride + kw) * Cin + c8 * 8);
    *reinterpret_cast<uint4*>(col + i * 8) = v;       // i * 8 == m * Kp + tap * Cin + c8 * 8
}

// ---- weights: OIHW fp32 -> [Cout][Kp] bf16 in patch order (and the inverse for the gradient, fp32 -> fp32)
__global__ void pack_weight_kernel(const float* __restrict__ w, bf16* __restrict__ wp, int Cout, int Cin, int KS, int Kp) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over Cout * Kp
    if (i >= (int64_t)Cout * Kp) return;
    const int k = (int)(i % Kp), o = (int)(i / Kp);
    float v = 0.f;
    if (k < KS * KS * Cin) {
        const int ci = k % Cin, tap = k / Cin;
        v = w[((int64_t)o * Cin + ci) * KS * KS + tap];
    }
    wp[i] = __float2bfloat16_rn(v);
}
__global__ void unpack_weight_grad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Cout, int Cin, int KS, int Kp) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over Cout * Cin * KS * KS (OIHW index)
    const int kk = KS * KS;
    if (i >= (int64_t)Cout * Cin * kk) return;
    const int tap = (int)(i % kk), ci = (int)((i / kk) % Cin), o = (int)(i / ((int64_t)kk * Cin));
    dw[i] = dwp[(int64_t)o * Kp + tap * Cin + ci];
}

// ---- 2x2 / stride-2 max-pool (floor) of y = relu(conv + bias) [B,OH,OW,C] with the arg-max mask of the direct kernels:
// 0..3 = (dy*2+dx) of the FIRST maximum (torch's max_pool2d order), 4 = ReLU-dead (maximum is 0)
__global__ void pool2x2_fwd_kernel(const bf16* __restrict__ y, bf16* __restrict__ out, uint8_t* __restrict__ mask,
                                   int64_t total, int OH, int OW, int PH, int PW, int C) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over B * PH * PW * C/8
    if (i >= total) return;
    const int c8n = C >> 3;
    const int c8 = (int)(i % c8n);
    const int pw = (int)((i / c8n) % PW), ph = (int)((i / ((int64_t)c8n * PW)) % PH);
    const int64_t b = i / ((int64_t)c8n * PW * PH);
    float best[8];
    uint32_t id[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const uint4 u = *reinterpret_cast<const uint4*>(y + ((b * OH + 2 * ph + (e >> 1)) * OW + 2 * pw + (e & 1)) * C + c8 * 8);
        float v[8];
        unpack8(u, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (e == 0 || v[j] > best[j]) { best[j] = v[j]; id[j] = e; }
    }
    uint4 o;
    uint32_t* ow_ = reinterpret_cast<uint32_t*>(&o);
    uint32_t m0 = 0, m1 = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (!(best[j] > 0.f)) { best[j] = 0.f; id[j] = 4u; }
        if (j < 4) m0 |= id[j] << (8 * j); else m1 |= id[j] << (8 * (j - 4));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(best[2 * j], best[2 * j + 1]);
        ow_[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(out + i * 8) = o;
    *reinterpret_cast<uint2*>(mask + i * 8) = make_uint2(m0, m1);
}

// ---- max-pool backward into the conv-output grid [B,OH,OW,C] (rows / columns the floor pooling dropped get zeros)
__global__ void unpool2x2_bwd_kernel(const bf16* __restrict__ da, const uint8_t* __restrict__ mask, bf16* __restrict__ dy,
                                     int64_t total, int OH, int OW, int PH, int PW, int C) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over B * OH * OW * C/8
    if (i >= total) return;
    const int c8n = C >> 3;
    const int c8 = (int)(i % c8n);
    const int ow = (int)((i / c8n) % OW), oh = (int)((i / ((int64_t)c8n * OW)) % OH);
    const int64_t b = i / ((int64_t)c8n * OW * OH);
    uint4 o = make_uint4(0, 0, 0, 0);
    const int ph = oh >> 1, pw = ow >> 1;
    if (ph < PH && pw < PW) {
        const int64_t src = (((b * PH + ph) * PW + pw) * C + c8 * 8);
        const uint4 g = *reinterpret_cast<const uint4*>(da + src);
        const uint2 mk = *reinterpret_cast<const uint2*>(mask + src);
        const uint32_t e = (uint32_t)((oh & 1) * 2 + (ow & 1));
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
        uint32_t r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t mb = j < 2 ? mk.x >> (16 * j) : mk.y >> (16 * (j - 2));       // two mask bytes of this channel pair
            const uint32_t lo = (mb & 0xffu) == e ? 0x0000ffffu : 0u, hi = ((mb >> 8) & 0xffu) == e ? 0xffff0000u : 0u;
            r[j] = gw[j] & (lo | hi);
        }
        o = make_uint4(r[0], r[1], r[2], r[3]);
    }
    *reinterpret_cast<uint4*>(dy + i * 8) = o;
}

// ---- data gradient: dx[b,ih,iw,ci] = sum over the taps (kh,kw) whose output position exists of dcol[(b,oh,ow)][(kh,kw,ci)]
__global__ void col2im_vec8_kernel(const bf16* __restrict__ dcol, bf16* __restrict__ dx, int64_t total, int IH, int IW, int Cin,
                                   int KS, int stride, int OH, int OW, int Kp) {
    pdl_trigger();
    pdl_wait();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;        // over B * IH * IW * Cin/8
    if (i >= total) return;
    const int c8n = Cin >> 3;
    const int c8 = (int)(i % c8n);
    const int iw = (int)((i / c8n) % IW), ih = (int)((i / ((int64_t)c8n * IW)) % IH);
    const int64_t b = i / ((int64_t)c8n * IW * IH);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int kh = 0; kh < KS; ++kh) {
        const int th = ih - kh;
        if (th < 0 || th % stride != 0) continue;
        const int oh = th / stride;
        if (oh >= OH) continue;
        for (int kw = 0; kw < KS; ++kw) {
            const int tw = iw - kw;
            if (tw < 0 || tw % stride != 0) continue;
            const int ow = tw / stride;
            if (ow >= OW) continue;
            const uint4 u = *reinterpret_cast<const uint4*>(dcol + ((b * OH + oh) * OW + ow) * Kp + (kh * KS + kw) * Cin + c8 * 8);
            float v[8];
            unpack8(u, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += v[j];
        }
    }
    uint4 o;
    uint32_t* ow_ = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
        ow_[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = o;
}

unsigned grid_for(int64_t total, int threads) { return (unsigned)ceil_div64(total, threads); }

}  // namespace

static int conv_dims_ok(int B, int IH, int IW, int Cin, int KS, int stride, int* OH, int* OW) {
    VQA_REQUIRE(B > 0 && IH > 0 && IW > 0 && Cin > 0 && KS > 0 && stride > 0 && IH >= KS && IW >= KS, "im2col conv: bad dims");
    *OH = (IH - KS) / stride + 1;
    *OW = (IW - KS) / stride + 1;
    return 0;
}

extern "C" int vqa_im2col(const void* x, int x_dtype, int nchw, void* col, int B, int IH, int IW, int Cin, int KS, int stride,
                          int Kp, void* stream) {
    int OH, OW;
    if (int e = conv_dims_ok(B, IH, IW, Cin, KS, stride, &OH, &OW)) return e;
    VQA_REQUIRE(x && col && Kp >= KS * KS * Cin && Kp % 8 == 0, "im2col: Kp=%d must be a multiple of 8 and >= %d", Kp, KS * KS * Cin);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t M = (int64_t)B * OH * OW;
    if (x_dtype == VQA_BF16 && !nchw && Cin % 8 == 0 && Kp == KS * KS * Cin && ((uintptr_t)x & 15) == 0 && ((uintptr_t)col & 15) == 0) {
        const int64_t total = M * KS * KS * (Cin / 8);
        VQA_CUDA(vqa_launch_pdl(im2col_vec8_kernel, dim3(grid_for(total, 256)), dim3(256), 0, st, (const bf16*)x, (bf16*)col, total,
                                IH, IW, Cin, KS, stride, OH, OW));
    } else {
        const int64_t total = M * Kp;
        if (x_dtype == VQA_F32)
            VQA_CUDA(vqa_launch_pdl(im2col_scalar_kernel<float>, dim3(grid_for(total, 256)), dim3(256), 0, st, (const float*)x, nchw, (bf16*)col,
                                    total, IH, IW, Cin, KS, stride, OH, OW, Kp));
        else if (x_dtype == VQA_F16)
            VQA_CUDA(vqa_launch_pdl(im2col_scalar_kernel<__half>, dim3(grid_for(total, 256)), dim3(256), 0, st, (const __half*)x, nchw, (bf16*)col,
                                    total, IH, IW, Cin, KS, stride, OH, OW, Kp));
        else if (x_dtype == VQA_BF16)
            VQA_CUDA(vqa_launch_pdl(im2col_scalar_kernel<bf16>, dim3(grid_for(total, 256)), dim3(256), 0, st, (const bf16*)x, nchw, (bf16*)col,
                                    total, IH, IW, Cin, KS, stride, OH, OW, Kp));
        else VQA_REQUIRE(false, "im2col: input dtype %d", x_dtype);
    }
    VQA_CHECK_LAUNCH("im2col");
    return 0;
}

extern "C" int vqa_conv_weight_pack_im2col(const float* w, void* wp, int Cout, int Cin, int KS, int Kp, void* stream) {
    VQA_REQUIRE(w && wp && Cout > 0 && Cin > 0 && KS > 0 && Kp >= KS * KS * Cin, "conv_weight_pack_im2col: bad arguments");
    const int64_t total = (int64_t)Cout * Kp;
    VQA_CUDA(vqa_launch_pdl(pack_weight_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, w, (bf16*)wp, Cout, Cin, KS, Kp));
    VQA_CHECK_LAUNCH("conv_weight_pack_im2col");
    return 0;
}

extern "C" int vqa_conv_weight_grad_unpack_im2col(const float* dwp, float* dw, int Cout, int Cin, int KS, int Kp, void* stream) {
    VQA_REQUIRE(dwp && dw && Cout > 0 && Cin > 0 && KS > 0 && Kp >= KS * KS * Cin, "conv_weight_grad_unpack_im2col: bad arguments");
    const int64_t total = (int64_t)Cout * Cin * KS * KS;
    VQA_CUDA(vqa_launch_pdl(unpack_weight_grad_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, dwp, dw, Cout, Cin, KS, Kp));
    VQA_CHECK_LAUNCH("conv_weight_grad_unpack_im2col");
    return 0;
}

extern "C" int vqa_pool2x2_fwd(const void* y, void* out, uint8_t* mask, int B, int OH, int OW, int C, void* stream) {
    VQA_REQUIRE(y && out && mask && B > 0 && OH >= 2 && OW >= 2 && C > 0 && C % 8 == 0, "pool2x2_fwd: bad arguments (C %% 8 == 0, conv output >= 2x2)");
    const int PH = OH / 2, PW = OW / 2;
    const int64_t total = (int64_t)B * PH * PW * (C / 8);
    VQA_CUDA(vqa_launch_pdl(pool2x2_fwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)y, (bf16*)out, mask,
                            total, OH, OW, PH, PW, C));
    VQA_CHECK_LAUNCH("pool2x2_fwd");
    return 0;
}

extern "C" int vqa_unpool2x2_bwd(const void* da, const uint8_t* mask, void* dy, int B, int OH, int OW, int C, void* stream) {
    VQA_REQUIRE(da && mask && dy && B > 0 && OH >= 2 && OW >= 2 && C > 0 && C % 8 == 0, "unpool2x2_bwd: bad arguments");
    const int PH = OH / 2, PW = OW / 2;
    const int64_t total = (int64_t)B * OH * OW * (C / 8);
    VQA_CUDA(vqa_launch_pdl(unpool2x2_bwd_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)da, mask, (bf16*)dy,
                            total, OH, OW, PH, PW, C));
    VQA_CHECK_LAUNCH("unpool2x2_bwd");
    return 0;
}

extern "C" int vqa_col2im(const void* dcol, void* dx, int B, int IH, int IW, int Cin, int KS, int stride, int Kp, void* stream) {
    int OH, OW;
    if (int e = conv_dims_ok(B, IH, IW, Cin, KS, stride, &OH, &OW)) return e;
    VQA_REQUIRE(dcol && dx && Cin % 8 == 0 && Kp >= KS * KS * Cin && Kp % 8 == 0, "col2im: Cin %% 8 == 0 and Kp %% 8 == 0 required");
    const int64_t total = (int64_t)B * IH * IW * (Cin / 8);
    VQA_CUDA(vqa_launch_pdl(col2im_vec8_kernel, dim3(grid_for(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const bf16*)dcol, (bf16*)dx, total,
                            IH, IW, Cin, KS, stride, OH, OW, Kp));
    VQA_CHECK_LAUNCH("col2im");
    return 0;
}
