"""Prototypes of the tcgen05 / TMA (tensor-core) entry points of libvqa_b200.so."""
import ctypes as C

_vp, _i, _i64, _u64, _u32, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_float

PROTOTYPES = {
    "vqa_tc_gemm": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i, _i64, _i64, _vp, _vp, _i64,
                    _i, _i, _i, _i, _i, _f, _u64, _u32, _vp],
    "vqa_tc_gemm_kblocks": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _vp, _vp],
    "vqa_lstm_active_kblocks": [_vp, _vp, _vp, _i, _i, _vp],
    "vqa_transpose_bf16": [_vp, _i, _i64, _i64, _vp, _i64, _i64, _i, _i, _i, _vp],
    "vqa_tc_conv3x3_relu_pool_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_tc_conv3x3_bwd_data": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_tc_conv3x3_bwd_data_unpool": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_tc_conv_set_cta_group": [_i],
    "vqa_pack_conv3x3_weight": [_vp, _vp, _vp, _i, _i, _vp],
    "vqa_unpool_bf16": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_tc_conv3x3_bwd_weight": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_tc_conv0_relu_pool_fwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_tc_conv0_bwd_weight_bias": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_tc_conv0_relu_pool_fwd_x": [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_tc_conv0_bwd_weight_bias_x": [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_tc_lstm_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_tc_lstm_fwd_ordered": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_tc_lstm_bwd_ordered": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_im2col": [_vp, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vqa_conv_weight_pack_im2col": [_vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_conv_weight_grad_unpack_im2col": [_vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_pool2x2_fwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_unpool2x2_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_col2im": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vqa_pack_lstm_whh": [_vp, _vp, _i, _vp],
    "vqa_tc_lstm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "vqa_tc_lstm_cluster_size": [],
    "vqa_nhwc_to_nchw_pad_bf16": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "vqa_unpool_nchw_bf16": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
}
