/* libvqa_b200 -- C ABI of the B200-native DL_VQA training / inference step.
 *
 * The reference (OmerShubi/DL_VQA) has no native boundary: its hot path is torch.nn modules called
 * from models/model.py:53-67 and the loss/score tail of train.py:190-207.  Each entry point below
 * replaces the torch/cuDNN/cuBLAS call(s) named in its comment (reference file:line).  The Python
 * host code in dl_vqa_b200/ binds these with ctypes; INTEGRATION.md shows the stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch's allocator); nothing here
 *     allocates, frees or synchronises; `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = ok, VQA_ERR_INVALID_ARGUMENT (<0) = rejected arguments, >0 = cudaError_t;
 *     vqa_last_error_string() describes the last failure on the calling thread;
 *   - dtype codes: VQA_F32 / VQA_BF16 describe ACTIVATION tensors; parameters are always fp32 in
 *     the PyTorch layouts of the reference's state_dict (OIHW conv, [out,in] linear, gate-stacked
 *     i,f,g,o LSTM) unless the entry says "packed";
 *   - image activations are NHWC ([B, H, W, C], channel contiguous); the network input is the
 *     reference's NCHW fp32 tensor;
 *   - dropout is a stateless counter-based mask keyed by (seed, site, element index): forward and
 *     backward entries take the same (p, seed) and regenerate the same mask.  A `seed` argument with
 *     VQA_SEED_ON_DEVICE set carries, in its low 63 bits, the DEVICE ADDRESS of a uint64 seed that is
 *     read when the kernel runs: a captured CUDA graph of the step then draws a fresh mask on every
 *     replay (vqa_step_tick advances the seed; the reference draws a new mask per call, models/model.py:84,156,185,194);
 *   - threading: every entry may be called from any host thread on any stream; entries keep no per-call state
 *     between calls.  Process-wide state is limited to (i) values read once and never written again (the
 *     environment switches VQA_PDL and VQA_LSTM_CLUSTER, per-kernel "shared-memory limit already raised" flags),
 *     (ii) the launch counter (atomic), (iii) ONE tuning knob, vqa_tc_conv_set_cta_group (atomic; every value
 *     produces identical results).  vqa_last_error_string() is per thread.
 */
#ifndef VQA_B200_H
#define VQA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQA_ABI_VERSION 1
#define VQA_F32 0
#define VQA_BF16 1
#define VQA_F16 2    /* the network INPUT (the reference stores its pre-processed images as float16, preprocessing/preprocess_images.py:40),
                        the output of vqa_tc_gemm where asked for, and v' of vqa_attention_*_x */
#define VQA_ERR_INVALID_ARGUMENT (-1)
#define VQA_ERR_UNSUPPORTED (-2)
#define VQA_SEED_ON_DEVICE (1ull << 63)

/* attention fusion operator, config.yaml train.attention.do_option (models/model.py:188-193) */
#define VQA_ATT_ADD 0
#define VQA_ATT_MUL 1
#define VQA_ATT_CAT 2   /* '|': x = relu(cat[v', q']), x_conv weight [G][2A], dwx_part [B][G][2A] (generic kernels only) */

const char* vqa_last_error_string(void);
int vqa_abi_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t vqa_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Image encoder -- models/model.py:72-84 (ImageNet2): Conv2d(k, stride, pad 0) -> ReLU -> MaxPool2d(2,2)
 * One fused kernel per layer; the un-pooled conv output never reaches HBM.
 *   x      : [B,IH,IW,Cin] NHWC (x_nchw=0) or [B,Cin,IH,IW] NCHW (x_nchw=1, the network input)
 *   w,bias : fp32 OIHW [Cout,Cin,KS,KS], [Cout]
 *   out    : [B,PH,PW,Cout] NHWC act_dtype, PH = ((IH-KS)/stride+1)/2 (floor, as MaxPool2d)
 *   mask   : [B,PH,PW,Cout] uint8: 0..3 = (dy*2+dx) of the window maximum, 4 = ReLU-dead
 * --------------------------------------------------------------------------------------------- */
int vqa_conv_relu_pool_fwd(const void* x, int x_dtype, int x_nchw, const float* w, const float* bias,
                           void* out, uint8_t* mask, int act_dtype,
                           int B, int IH, int IW, int Cin, int Cout, int KS, int stride, void* stream);
/* gradient w.r.t. the layer input (autograd of the three modules above):
 *   dpool [B,PH,PW,Cout] act_dtype, mask from forward -> dx [B,IH,IW,Cin] NHWC act_dtype */
int vqa_conv_bwd_data(const void* dpool, const uint8_t* mask, const float* w, void* dx, int act_dtype,
                      int B, int IH, int IW, int Cin, int Cout, int KS, int stride, void* stream);
/* gradient w.r.t. weight (OIHW fp32, overwritten) and bias (fp32, overwritten) */
int vqa_conv_bwd_weight(const void* x, int x_dtype, int x_nchw, const void* dpool, const uint8_t* mask,
                        float* dw, float* db, int act_dtype,
                        int B, int IH, int IW, int Cin, int Cout, int KS, int stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * image.drop + channel L2 normalisation -- models/model.py:84 and :56
 *   x [R,C] (R = B*H*W rows, NHWC) -> vn = drop_img(x) / (||drop_img(x)||_2 + 1e-12)   [R,C]
 *   vnd = drop_att(vn) (the input of attention.v_conv, models/model.py:185); may be NULL when p_att == 0
 *   nrm [R] fp32 = the norms, saved for backward
 * backward: dvn (may be NULL) and dvnd (may be NULL) are the gradients w.r.t. vn and vnd -> dx
 * --------------------------------------------------------------------------------------------- */
int vqa_dropnorm_fwd(const void* x, void* vn, void* vnd, float* nrm, int act_dtype, int64_t R, int C,
                     float p_img, float p_att, uint64_t seed, void* stream);
int vqa_dropnorm_bwd(const void* dvn, const void* dvnd, const void* vn, const float* nrm, void* dx,
                     int act_dtype, int64_t R, int C, float p_img, float p_att, uint64_t seed, void* stream);
/* bf16 arm: the same backward fused with the 2x2 max-pool backward of the last conv layer and its bias gradient:
 * writes dy [B,2PH,2PW,C] (un-pooled gradient, zero where mask != window element) and db [C] (overwritten) instead of
 * the compact gradient; mask [B,PH,PW,C] from the conv forward.  C % 8 == 0, C <= 256. */
int vqa_dropnorm_bwd_unpool(const void* dvn, const void* dvnd, const void* vn, const float* nrm, const uint8_t* mask,
                            void* dy, float* db, int B, int PH, int PW, int C, float p_img, float p_att,
                            uint64_t seed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Question encoder -- models/model.py:151-166
 * Sequence buffers are STEP-indexed: [dirs][T][B][...]; step s of direction 0 consumes token t=s,
 * step s of direction 1 consumes t=len-1-s (what pack_padded_sequence + a reverse LSTM does);
 * steps s >= len[b] are inactive (state frozen).
 *
 * embed: xs[dir][s][b][0:E] = tanh(drop(embedding[q[b,t]]))  (models/model.py:155-157); pad columns
 *        E..ldx-1 are zero-filled; inactive rows are zero.
 * --------------------------------------------------------------------------------------------- */
int vqa_embed_tanh_fwd(const int64_t* q, const int64_t* q_len, const float* emb, void* xs, int act_dtype,
                       int B, int T, int E, int ldx, int dirs, float p, uint64_t seed, void* stream);
/* demb (fp32 [V,E], ACCUMULATED into; caller zeroes) += scatter of dxs * (1-xs^2) * mask; token 0
 * (padding_idx, models/model.py:138-140) receives nothing */
int vqa_embed_tanh_bwd(const int64_t* q, const int64_t* q_len, const void* xs, const void* dxs, float* demb,
                       int act_dtype, int B, int T, int E, int ldx, int dirs, float p, uint64_t seed, void* stream);
/* Host-side diagnostic (no GPU work): the vector dropout scheme of the attention / dropout+L2-norm kernels reads each
 * 16-bit random field as a bf16 bit pattern and drops the element iff it compares below a threshold; *pattern is that
 * threshold's bit pattern for probability p (0 = dropout off), *count16 = round(p * 65536), the number of patterns that
 * should compare below it (nn.Dropout(p), models/model.py:84,185,194). */
int vqa_dropout_threshold_pattern(float p, uint32_t* pattern, uint32_t* count16);
/* Length order of the batch -- the sort pack_padded_sequence(enforce_sorted=False) does on the host in the reference
 * (models/model.py:160): order[j] (int32) = sample at position j of the batch sorted by DESCENDING length, ties in sample
 * order; len_sorted[j] (int64) = q_len[order[j]] clamped to [0, T].  No host synchronisation.  B <= 8192. */
int vqa_length_order(const int64_t* q_len, int* order, int64_t* len_sorted, int B, int T, void* stream);
/* The same two kernels with the step-indexed rows in that order: row j of xs / dxs belongs to sample order[j]
 * (order == NULL: identity).  q and q_len stay in sample order; the dropout mask is a function of the sample index, so
 * ordering the batch does not change it. */
int vqa_embed_tanh_fwd_ordered(const int64_t* q, const int64_t* q_len, const int* order, const float* emb, void* xs,
                               int act_dtype, int B, int T, int E, int ldx, int dirs, float p, uint64_t seed, void* stream);
int vqa_embed_tanh_bwd_ordered(const int64_t* q, const int64_t* q_len, const int* order, const void* xs, const void* dxs,
                               float* demb, int act_dtype, int B, int T, int E, int ldx, int dirs, float p, uint64_t seed,
                               void* stream);

/* one LSTM time step for all directions (nn.LSTM, models/model.py:164): gates = gx[:,s] + h_{s-1} W_hh^T,
 * i,f,g,o nonlinearities, c/h update fused in the GEMM epilogue.
 *   gx [dirs][T][B][4H] act_dtype: in = x W_ih^T + b_ih + b_hh, out = activated gates (kept for backward)
 *   cs [dirs][T][B][H] fp32, hs [dirs][T][B][H] act_dtype; qf [B][dirs*H] act_dtype gets c at s == T-1
 *   w_hh fp32 [4H][H] per direction; direction d at w_hh + d*w_hh_dir_stride (elements) */
int vqa_lstm_step_fwd(void* gx, float* cs, void* hs, void* qf, const float* w_hh, int64_t w_hh_dir_stride,
                      const int64_t* q_len, int act_dtype, int s, int T, int B, int H, int dirs, void* stream);
/* backward pointwise part of step s: consumes dh (fp32 [dirs][B][H], gradient w.r.t. h_s) and CLEARS the entries it
 * read (the next step's dh = dg W_hh may then be accumulated by a split-K GEMM), updates the
 * running dc (fp32 [dirs][B][H]) in place, writes pre-activation gate gradients dg[dirs][s][B][4H].
 * dc_init (act_dtype [B][dirs*H], gradient w.r.t. the final cell state) is non-NULL on the first
 * processed step (s == T-1) and replaces the running dc there */
int vqa_lstm_step_bwd_pointwise(const void* gates, const float* cs, float* dh, float* dc,
                                const void* dc_init, void* dg, const int64_t* q_len, int act_dtype,
                                int s, int T, int B, int H, int dirs, void* stream);

/* ---------------------------------------------------------------------------------------------
 * General dense contraction (nn.Linear / 1x1 Conv2d / their autograd): for z in [0,nbatch)
 *   C[z][m, n] (op)= act( sum_k A[z](m,k) * B[z](n,k) + bias[z][n] ) * dropout
 * A(m,k) = A[m*a_sr + k*a_sk], B(n,k) = B[n*b_sr + k*b_sk] (element strides), C row-major with ldc.
 *   flags: VQA_GEMM_RELU, VQA_GEMM_ACCUMULATE (C += ...), VQA_GEMM_SPLITK (fp32 C only: atomically
 *   accumulates K-slices into C, which the caller has zeroed; bias/relu/dropout not allowed)
 * --------------------------------------------------------------------------------------------- */
#define VQA_GEMM_RELU 1
#define VQA_GEMM_ACCUMULATE 2
#define VQA_GEMM_SPLITK 4
#define VQA_GEMM_OPERANDS_MN 8   /* vqa_tc_gemm only: A stored [K,M], B stored [K,N] (reduction index = row) */
#define VQA_GEMM_B_MN 16         /* vqa_tc_gemm only: A stored [M,K], B stored [K,N]: dX = dY W with W as stored */
int vqa_gemm(const void* A, int a_dtype, int64_t a_sr, int64_t a_sk, int64_t a_sb,
             const void* B, int b_dtype, int64_t b_sr, int64_t b_sk, int64_t b_sb,
             void* C, int c_dtype, int64_t ldc, int64_t c_sb,
             const float* bias, const float* bias2, int64_t bias_sb,
             int M, int N, int K, int nbatch, int flags,
             float p_drop, uint64_t seed, uint32_t site, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused attention -- models/model.py:187-195 (tile, + / * / | fusion, ReLU, dropout, x_conv 1x1) and
 * models/model.py:208-221 (spatial softmax per glimpse, weighted pooling).  One memory-bound kernel.
 *   vp [B,P,A] act_dtype = v_conv output; qp [B,A] fp32 = q_lin output; vn [B,P,C] act_dtype (normalised,
 *   un-dropped features); wx [G,A] (op VQA_ATT_CAT: [G,2A], v' half first), bx [G] fp32 = x_conv;
 *   the backward's dwx_part is [B,G,A] ([B,G,2A] for CAT); prob [B,G,P] fp32 (softmax, saved);
 *   out: row b at out + b*ldo, [G*C] act_dtype, glimpse-major (the first G*C columns of `combined`)
 * --------------------------------------------------------------------------------------------- */
int vqa_attention_fwd(const void* vp, const float* qp, const void* vn, const float* wx, const float* bx,
                      float* prob, void* out, int64_t ldo, int act_dtype, int op,
                      int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream);
/* dout: row b at dout + b*ldd, [G*C] act_dtype.  Writes dvp [B,P,A], dvn [B,P,C] (act_dtype),
 * dqp [B,A] fp32, and per-sample partials dwx_part [B,G,A], dbx_part [B,G] fp32 (column-sum them) */
int vqa_attention_bwd(const void* dout, int64_t ldd, const void* vp, const float* qp, const void* vn,
                      const float* wx, const float* prob, void* dvp, void* dvn, float* dqp,
                      float* dwx_part, float* dbx_part, int act_dtype, int op,
                      int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream);

/* The same two entries with the dtype of v' explicit.  vp_dtype = act_dtype, or VQA_F16 where
 * vqa_attention_streaming_ok(...) returns 1 (tensor-core arm at the config.yaml widths: A = 1024, C = 256, '+' / '*',
 * G <= 2, grid small enough for shared memory): the '+' fusion of models/model.py:188-191 rounds v' + q' to the 16-bit
 * format of v', and q' is about ten times larger than the spatial variation of v' that the x_conv weight gradient sees,
 * so fp16 (11-bit mantissa; the v_conv GEMM writes it directly, saturating) instead of bf16 takes that gradient's error
 * from 1-2 % to ~0.3 % of max-norm at no cost.  dvp stays act_dtype. */
int vqa_attention_streaming_ok(int act_dtype, int op, int P, int A, int C, int G);
int vqa_attention_fwd_x(const void* vp, int vp_dtype, const float* qp, const void* vn, const float* wx, const float* bx,
                        float* prob, void* out, int64_t ldo, int act_dtype, int op,
                        int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream);
int vqa_attention_bwd_x(const void* dout, int64_t ldd, const void* vp, int vp_dtype, const float* qp, const void* vn,
                        const float* wx, const float* prob, void* dvp, void* dvn, float* dqp,
                        float* dwx_part, float* dbx_part, int act_dtype, int op,
                        int B, int P, int A, int C, int G, float p_drop, uint64_t seed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Soft-target loss + VQA score -- train.py:190-206 and utils/train_utils.py:12-25, one pass:
 *   loss = (1/B) sum_b sum_j [a_idx[b,j] != 0] * (a_val[b,j]/10) * (-log_softmax(logits[b]))[a_idx[b,j]-1]
 *   score = sum_b min(0.3 * count_b(argmax logits[b]), 1)
 *   dlogits (may be NULL) = d loss / d logits
 *   loss_rows/score_rows: [B] fp32 scratch; loss_out/score_out: 1 fp32 each
 * --------------------------------------------------------------------------------------------- */
int vqa_softloss_fwd_bwd(const float* logits, const int64_t* a_idx, const int64_t* a_val, float* dlogits,
                         float* loss_rows, float* score_rows, float* loss_out, float* score_out,
                         int B, int N, int A, void* stream);

/* ---------------------------------------------------------------------------------------------
 * small fused helpers
 * --------------------------------------------------------------------------------------------- */
/* out[r, c] = in[r, c] * dropout(site, r*cols + c); in/out row pitches ld_in/ld_out (elements) */
int vqa_dropout_apply(const void* in, int64_t ld_in, void* out, int64_t ld_out, int dtype, int64_t rows,
                      int cols, float p, uint64_t seed, uint32_t site, void* stream);
/* keep[i] = 1 iff element i of dropout site `site` (the VQA_SITE_* stream of one nn.Dropout call site) survives under
 * (p, seed).  vector8 = 0: the per-element scheme used by the embedding / classifier / q' sites; 1: the 8-element vector
 * scheme used by the image, v and attention-x sites.  Exposes the stateless mask so that a caller (the parity tests) can
 * restate a train-mode forward/backward exactly; the step itself never materialises a mask. */
#define VQA_SITE_IMAGE 0    /* models/model.py:84   image.drop,        element = NHWC index of the last conv output   */
#define VQA_SITE_ATT_V 1    /* models/model.py:185  attention.drop(v), element = index into vn [B*P, C]               */
#define VQA_SITE_EMBED 2    /* models/model.py:156  text.drop,         element = (b*T + t)*E + e                      */
#define VQA_SITE_ATT_Q 3    /* models/model.py:186  attention.drop(q), element = index into q [B, dirs*H]             */
#define VQA_SITE_ATT_X 4    /* models/model.py:194  attention.drop(x), element = (b*P + s)*A + a                      */
#define VQA_SITE_CLS_IN 5   /* models/model.py:201  classifier.drop1,  element = index into combined [B, G*C + QF]    */
#define VQA_SITE_CLS_HID 6  /* models/model.py:204  classifier.drop2,  element = index into hidden [B, hidden]        */
int vqa_dropout_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint32_t site, int vector8, void* stream);
/* out[c] += sum_r in[r*ld + c]   (fp32 out, caller zeroes); mask != NULL: only rows where mask[r*ld+c] < 4 */
int vqa_colsum(const void* in, int dtype, int64_t ld, const uint8_t* mask, float* out, int64_t rows, int cols,
               void* stream);
/* dst = (dtype) src, n elements */
int vqa_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, void* stream);
/* dst[r, 0:dst_cols] = (dtype) src[r, 0:cols] with zero fill of columns cols..dst_cols-1; row pitches lds/ldd */
int vqa_cast_2d(const void* src, int src_dtype, int64_t lds, void* dst, int dst_dtype, int64_t ldd,
                int64_t rows, int cols, int dst_cols, void* stream);
/* ReLU + dropout backward for classifier.lin1: dz = dy * [y > 0] * scale, y = dropped ReLU output */
int vqa_relu_drop_bwd(const void* dy, const void* y, void* dz, int dtype, int64_t n, float p, void* stream);
/* dst[r, 0:cols] = a[r,:] + b[r,:] * dropout(site, r*cols+c)  (b may be NULL) -- gradient merges */
int vqa_add_dropped(const void* a, int64_t lda, const void* b, int64_t ldb, void* dst, int64_t ldd, int dtype,
                    int64_t rows, int cols, float p, uint64_t seed, uint32_t site, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Adam -- train.py:55,76-80 (torch.optim.Adam defaults, lr set per step by update_learning_rate).
 * Multi-tensor: n tensors described by device pointer tables (each [n] of device pointers / sizes
 * resident in DEVICE memory).  bf16_copy[i] may be NULL; otherwise receives the updated parameter.
 * --------------------------------------------------------------------------------------------- */
int vqa_adam_multi(float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, void* const* bf16_copy, const int64_t* sizes, int n,
                   int64_t max_size, float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Device-resident step state -- the per-iteration host work of the training loop (train.py:69-81) moved onto the
 * GPU so that the whole step (forward, loss, backward, all-reduce, Adam) can be ONE captured CUDA graph that is
 * replayed with no per-step host arithmetic:
 *   seed        dropout seed of the step (read by every kernel whose seed argument is VQA_SEED_ON_DEVICE | &state->seed)
 *   iteration   train.py:50,81 total_iterations
 *   adam_step   torch.optim.Adam's per-parameter `step`
 *   lr          train.py:31-35 update_learning_rate: lr0 * 0.5 ** (iteration / half_life)
 *   lr_over_bc1, inv_sqrt_bc2   Adam's bias corrections folded as the update kernel consumes them
 * vqa_step_tick (one thread) runs FIRST in the step: it publishes lr / bias corrections for the CURRENT
 * iteration, then advances iteration, adam_step and the seed (splitmix64) -- the order of train.py:76-81.
 * vqa_adam_multi_dev is vqa_adam_multi with lr and step taken from the state instead of from the host.
 * --------------------------------------------------------------------------------------------- */
typedef struct VqaStepState {
    uint64_t seed;
    int64_t iteration;
    int64_t adam_step;
    float lr;
    float lr_over_bc1;
    float inv_sqrt_bc2;
    float reserved[7];
} VqaStepState;                                   /* 64 bytes */
int vqa_step_tick(VqaStepState* state, double lr0, double half_life, double beta1, double beta2, void* stream);
int vqa_adam_multi_dev(float* const* params, const float* const* grads, float* const* exp_avg,
                       float* const* exp_avg_sq, void* const* bf16_copy, const int64_t* sizes, int n,
                       int64_t max_size, const VqaStepState* state, float beta1, float beta2, float eps,
                       float grad_scale, void* stream);
/* dst[i] = src[i] * *scalar (fp32; dst may alias src): the incoming d(loss) of the loss node, without a host read */
int vqa_scale_by_device_scalar(const float* src, float* dst, const float* scalar, int64_t n, void* stream);
/* bytes of device memory set to zero on `stream` (a memset node under graph capture, not a kernel) */
int vqa_zero(void* ptr, int64_t bytes, void* stream);
/* device-to-device copy of `bytes` bytes on `stream` (a memcpy node under graph capture) */
int vqa_copy(void* dst, const void* src, int64_t bytes, void* stream);

/* =============================================================================================
 * Tensor-core arm (bf16 operands, fp32 accumulation): tcgen05.mma with TMEM accumulators, operands
 * staged by TMA into 128-byte-swizzled shared memory.  Same math as the entries above.
 * ============================================================================================= */

/* C[z][m,n] (c_dtype VQA_F32 / VQA_BF16 / VQA_F16 (saturating), row pitch ldc) = act(sum_k A[z][m,k] * B[z][n,k] + bias[z][n] + bias2[z][n]) * dropout
 * A [M,K] and B [N,K] are bf16, K contiguous, row pitches lda/ldb (multiples of 8 elements), 16-byte
 * aligned.  flags as vqa_gemm (VQA_GEMM_RELU, VQA_GEMM_SPLITK; ACCUMULATE unsupported).  With
 * VQA_GEMM_OPERANDS_MN the operands are stored reduction-major: A [K,M], B [K,N] (pitches lda/ldb); this is the
 * weight-gradient form dW[N,K'] = dY[rows,N]^T X[rows,K'] consumed without transposes. */
int vqa_tc_gemm(const void* A, int64_t lda, int64_t a_sb, const void* B, int64_t ldb, int64_t b_sb,
                void* C, int c_dtype, int64_t ldc, int64_t c_sb,
                const float* bias, const float* bias2, int64_t bias_sb,
                int M, int N, int K, int nbatch, int flags,
                float p_drop, uint64_t seed, uint32_t site, void* stream);
/* The reduction-major form of vqa_tc_gemm (VQA_GEMM_OPERANDS_MN, optionally VQA_GEMM_SPLITK into a zeroed C; nbatch = 1,
 * fp32 C [M][ldc]) restricted to the 64-row reduction blocks named by a DEVICE-side list: kblocks[0] = n,
 * kblocks[1..n] = block indices (block j = rows 64j..64j+63 of A [K,M] and B [K,N]), each at most once.  Blocks that are
 * not listed are not read and contribute nothing: the caller guarantees one operand is all zero there.  Replaces what
 * cuDNN's packed-sequence RNN backward gets from pack_padded_sequence (models/model.py:160-164): padded (step, row)
 * positions never enter the LSTM weight gradients.  The list is read when the kernel RUNS (graph-capture safe). */
int vqa_tc_gemm_kblocks(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc,
                        int M, int N, int K, int flags, const int32_t* kblocks, void* stream);
/* Block lists for the above from the row lengths of the step-indexed LSTM buffers (len_rows [B] int64 in ROW order: the
 * len_sorted of vqa_length_order, or q_len when the rows are in sample order; B % 64 == 0): list0 = live blocks of steps
 * 0..T-1 (capacity 1 + T*B/64: the dW_ih reduction), list1 = live blocks of steps 1..T-1, indexed from step 1 (capacity
 * 1 + (T-1)*B/64: the dW_hh reduction, whose operands start at step 1 / h_0).  A block is live when any of its rows r has
 * len_rows[r] > step.  Either list may be NULL. */
int vqa_lstm_active_kblocks(const int64_t* len_rows, int32_t* list0, int32_t* list1, int B, int T, void* stream);
/* dst[z][c, r] (bf16, row pitch ldd) = src[z][r, c] (fp32 or bf16, row pitch lds): operand re-layout for the
 * contractions whose reduction index is not contiguous in memory (weight / data gradients) */
int vqa_transpose_bf16(const void* src, int src_dtype, int64_t lds, int64_t s_sb, void* dst, int64_t ldd,
                       int64_t d_sb, int rows, int cols, int nbatch, void* stream);

/* 3x3 stride-1 conv + bias + ReLU + 2x2 max-pool as an implicit GEMM on tcgen05 (replaces the cuDNN conv +
 * ATen relu + max_pool2d of models/model.py:80-82 for layers with Cin % 64 == 0, Cout in {64,128,256}).
 *   x [B,IH,IW,Cin] bf16 NHWC; wp [Cout][9*Cin] bf16 packed by vqa_pack_conv3x3_weight; bias fp32
 *   out / mask [B,PH,PW,Cout] as vqa_conv_relu_pool_fwd */
int vqa_tc_conv3x3_relu_pool_fwd(const void* x, const void* wp, const float* bias, void* out, uint8_t* mask,
                                 int B, int IH, int IW, int Cin, int Cout, void* stream);
/* data gradient: dy [B,2PH,2PW,Cout] bf16 = un-pooled gradient (vqa_unpool_bf16), wd [Cin][9*Cout] bf16 packed,
 * dx [B,IH,IW,Cin] bf16 */
int vqa_tc_conv3x3_bwd_data(const void* dy, const void* wd, void* dx,
                            int B, int IH, int IW, int Cin, int Cout, void* stream);
/* the same data gradient fused with the 2x2 max-pool backward and the bias gradient of the layer BELOW (whose pooled
 * output [B,IH,IW,Cin] is this layer's input): mask_below [B,IH,IW,Cin] uint8 from that layer's forward; writes its
 * un-pooled gradient dy_below [B,2IH,2IW,Cin] bf16 and bias gradient db_below [Cin] fp32 (both overwritten).
 * Replaces vqa_tc_conv3x3_bwd_data followed by vqa_unpool_bf16. */
int vqa_tc_conv3x3_bwd_data_unpool(const void* dy, const void* wd, const uint8_t* mask_below, void* dy_below,
                                   float* db_below, int B, int IH, int IW, int Cin, int Cout, void* stream);
/* Tuning knob of the entries above: 1 = single-CTA tcgen05.mma (cta_group::1), 2 = CTA pairs (cluster of 2,
 * cta_group::2: each CTA stages half of the weight rows), 0 (default) = chosen per layer shape from measurements.
 * Process-wide; results are identical either way. */
int vqa_tc_conv_set_cta_group(int cta_group);
/* w fp32 OIHW [Cout,Cin,3,3] -> wp[co][tap][ci] and/or wd[ci][tap][co] (bf16; either may be NULL) */
int vqa_pack_conv3x3_weight(const float* w, void* wp, void* wd, int Cout, int Cin, void* stream);
/* dy[b,2ph+dy,2pw+dx,c] = mask[b,ph,pw,c] == dy*2+dx ? dpool[b,ph,pw,c] : 0  (bf16, C % 8 == 0); when db != NULL
 * the same pass also produces the conv bias gradient db[c] = sum of dpool over positions with mask < 4 (overwritten) */
int vqa_unpool_bf16(const void* dpool, const uint8_t* mask, void* dy, float* db, int B, int PH, int PW, int C, void* stream);

/* weight gradient of the 3x3 conv on tcgen05, straight from the NHWC tensors (MN-major UMMA operands):
 * x [B,IH,IW,Cin] bf16 (layer input), dy [B,2PH,2PW,Cout] bf16 (vqa_unpool_bf16).  dw fp32 OIHW, overwritten.
 * Cin in {64,128}, Cout % 128 == 0. */
int vqa_tc_conv3x3_bwd_weight(const void* x, const void* dy, float* dw,
                              int B, int IH, int IW, int Cin, int Cout, void* stream);
/* first layer (Cin = 3, Cout = 64, 3x3, stride 1) on tcgen05, window-major: the im2col rows of every 2x2 pooling
 * window are built in shared memory by the kernel itself from the NCHW fp32 network input (K = 27 is too small for a
 * TMA pipeline), the four window elements accumulate side by side in TMEM and the max-pool is a per-thread max.
 * out / mask as vqa_conv_relu_pool_fwd.
 * backward: weight AND bias gradient straight from the pooled gradient dpool [B,PH,PW,64] bf16 and the forward's mask
 * (the un-pooled gradient is expanded in shared memory only); dw fp32 [64,3,3,3] and db fp32 [64] are overwritten.
 * Replaces the autograd of models/model.py:80-82 for layer 0 (cuDNN wgrad + max_pool2d backward + bias reduction). */
int vqa_tc_conv0_relu_pool_fwd(const float* x, const float* w, const float* bias, void* out, uint8_t* mask,
                               int B, int IH, int IW, int Cin, int Cout, void* stream);
int vqa_tc_conv0_bwd_weight_bias(const float* x, const void* dpool, const uint8_t* mask, float* dw, float* db,
                                 int B, int IH, int IW, int Cin, int Cout, void* stream);
/* the same two entries with the input dtype explicit: x_dtype VQA_F32 or VQA_F16.  Float16 is what the reference's
 * preprocessing stores (preprocessing/preprocess_images.py:40) and its Dataset widens on the host
 * (preprocessing/data_preprocessing.py:174); reading it here directly halves the host->device copy and the input reads
 * of both kernels and needs no cast kernel.  Widening fp16 -> fp32 is exact, so results are bit-identical. */
int vqa_tc_conv0_relu_pool_fwd_x(const void* x, int x_dtype, const float* w, const float* bias, void* out, uint8_t* mask,
                                 int B, int IH, int IW, int Cin, int Cout, void* stream);
int vqa_tc_conv0_bwd_weight_bias_x(const void* x, int x_dtype, const void* dpool, const uint8_t* mask, float* dw, float* db,
                                   int B, int IH, int IW, int Cin, int Cout, void* stream);
/* ---------------------------------------------------------------------------------------------
 * General convolution layers on the tensor-core arm (models/model.py:72-84 for ANY cfg image.stride /
 * image.kernel_size / image.num_channels, e.g. config/config_eval.yaml:52-62: stride 2): the layers outside the range
 * of the direct kernels above run as  vqa_im2col -> vqa_tc_gemm (bias + ReLU epilogue) -> vqa_pool2x2_fwd  and, backward,
 * vqa_unpool2x2_bwd -> vqa_tc_gemm (VQA_GEMM_OPERANDS_MN: weight gradient) / vqa_tc_gemm (VQA_GEMM_B_MN) -> vqa_col2im.
 * Patch-matrix column order: k = (kh*KS + kw)*Cin + ci, zero padded to Kp (multiple of 8).  OH = (IH-KS)/stride + 1.
 * --------------------------------------------------------------------------------------------- */
/* col [B*OH*OW][Kp] bf16 from x: NCHW (nchw = 1; fp32 / fp16 / bf16) or NHWC (nchw = 0) */
int vqa_im2col(const void* x, int x_dtype, int nchw, void* col, int B, int IH, int IW, int Cin, int KS, int stride,
               int Kp, void* stream);
/* w [Cout][Cin][KS][KS] fp32 (nn.Conv2d layout) -> wp [Cout][Kp] bf16 in patch order; and the fp32 gradient back */
int vqa_conv_weight_pack_im2col(const float* w, void* wp, int Cout, int Cin, int KS, int Kp, void* stream);
int vqa_conv_weight_grad_unpack_im2col(const float* dwp, float* dw, int Cout, int Cin, int KS, int Kp, void* stream);
/* nn.MaxPool2d(2, 2) (floor) of y = relu(conv + bias) [B][OH][OW][C] bf16 -> out [B][OH/2][OW/2][C] and the arg-max mask
 * of the direct kernels (0..3 = dy*2+dx of the first maximum, 4 = ReLU-dead); C % 8 == 0 */
int vqa_pool2x2_fwd(const void* y, void* out, uint8_t* mask, int B, int OH, int OW, int C, void* stream);
/* its backward: da [B][OH/2][OW/2][C] -> dy [B][OH][OW][C] (rows / columns dropped by the floor get zeros) */
int vqa_unpool2x2_bwd(const void* da, const uint8_t* mask, void* dy, int B, int OH, int OW, int C, void* stream);
/* dx [B][IH][IW][Cin] bf16 = overlap-add of dcol [B*OH*OW][Kp]; Cin % 8 == 0 */
int vqa_col2im(const void* dcol, void* dx, int B, int IH, int IW, int Cin, int KS, int stride, int Kp, void* stream);

/* Persistent LSTM recurrence, all steps and directions in one cooperative launch (replaces the cuDNN RNN of
 * models/model.py:164).  W_hh stays resident in shared memory (64 gate rows per CTA), h is exchanged through L2.
 *   gx [dirs][T][B][4H] bf16 (in: x W_ih^T + b_ih + b_hh, out: activated gates), cs [dirs][T][B][H] fp32,
 *   hs [dirs][T+1][B][H] bf16 with slot 0 zero-filled by the caller (slot s+1 = h after step s),
 *   qf [B][dirs*H] bf16 = final cell state, wp [dirs][4H][H] bf16 from vqa_pack_lstm_whh (one call per direction),
 *   sync: `dirs` uint32 scratch counters.  H % 64 == 0, H <= 1024; batches above 256 sequences run as consecutive
 *   cooperative launches of 256 (two 128-row TMEM accumulator tiles per CTA). */
int vqa_tc_lstm_fwd(void* gx, float* cs, void* hs, void* qf, const void* wp, const int64_t* q_len,
                    unsigned int* sync, int T, int B, int H, int dirs, void* stream);
int vqa_pack_lstm_whh(const float* w_hh, void* wp, int H, void* stream);
/* backward recurrence of the same LSTM in ONE cooperative launch (autograd of models/model.py:164; replaces T calls of
 * vqa_lstm_step_bwd_pointwise interleaved with T-1 split-K vqa_tc_gemm calls): gates / dg [dirs][T][B][4H] bf16
 * (activated gates saved by the forward -> gate gradients), cs [dirs][T][B][H] fp32, dh [dirs][B][H] fp32 ZEROED by
 * the caller, dc [dirs][B][H] fp32 scratch, dc_init [B][dirs*H] bf16 (gradient w.r.t. the final cell state),
 * whh [dirs][4H][H] bf16 (recurrent weights as stored), sync: 256 uint32 of scratch.  H % 128 == 0 and
 * ceil(B/128) * (H/128) * dirs <= #SMs. */
int vqa_tc_lstm_bwd(const void* gates, const float* cs, float* dh, float* dc, const void* dc_init, void* dg,
                    const void* whh, const int64_t* q_len, unsigned int* sync, int T, int B, int H, int dirs,
                    void* stream);
/* Length-ordered forms (what the packed sequence buys the reference's cuDNN RNN, models/model.py:160-164): the rows of
 * gx / cs / hs / gates / dg are the samples in descending length order (row j = sample order[j], vqa_length_order),
 * q_len is the matching len_sorted; qf is written and dc_init is read in SAMPLE order.  Both kernels drop a 128-row
 * tile from every step at which none of its sequences is active, so with ordered rows the short half of the batch
 * leaves the recurrent GEMMs, the operand ingest and the step synchronisation early.  They are correct for any row
 * order (order == NULL: rows are samples, which is what the two entries above pass). */
int vqa_tc_lstm_fwd_ordered(void* gx, float* cs, void* hs, void* qf, const void* wp, const int64_t* q_len, const int* order,
                            unsigned int* sync, int T, int B, int H, int dirs, void* stream);
int vqa_tc_lstm_bwd_ordered(const void* gates, const float* cs, float* dh, float* dc, const void* dc_init, void* dg,
                            const void* whh, const int64_t* q_len, const int* order, unsigned int* sync, int T, int B, int H,
                            int dirs, void* stream);
/* diagnostic: 4 = vqa_tc_lstm_fwd runs as clusters of four CTAs with TMA multicast of h (opt-in through the
 * environment variable VQA_LSTM_CLUSTER=4), 1 = single CTAs (default, or the driver rejected the cooperative cluster
 * launch), 0 = not launched yet */
int vqa_tc_lstm_cluster_size(void);
/* channel-major re-layout helpers (NHWC -> [B,C,H,Wp], zero padded pitch) */
int vqa_nhwc_to_nchw_pad_bf16(const void* x, void* xT, int B, int H, int W, int C, int Wp, void* stream);
int vqa_unpool_nchw_bf16(const void* dpool, const uint8_t* mask, void* dyT, int B, int PH, int PW, int C,
                         int OWpp, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQA_B200_H */
