"""ctypes binding of libvqa_b200.so (the C ABI declared in include/vqa_b200.h).

There is no CPU fallback: if the shared library is missing, or a call is attempted on a non-CUDA
tensor, this module raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"`
or `make -C dl_vqa_b200/csrc`.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvqa_b200.so")

F32, BF16, F16 = 0, 1, 2          # F16: the network input, and v' (attention.v_conv output) on the streaming attention kernels
ATT_ADD, ATT_MUL, ATT_CAT = 0, 1, 2
GEMM_RELU, GEMM_ACCUMULATE, GEMM_SPLITK, GEMM_OPERANDS_MN, GEMM_B_MN = 1, 2, 4, 8, 16
SITE_IMAGE, SITE_ATT_V, SITE_EMBED, SITE_ATT_Q, SITE_ATT_X, SITE_CLS_IN, SITE_CLS_HID = range(7)

DEFAULT_CONV_CTA_GROUP = 0      # tcgen05 cta_group of the 3x3 conv kernels: 0 = per-shape choice, 1 / 2 forced (vqa_tc_conv_set_cta_group)

_vp, _i, _i64, _u64, _u32, _f, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_float, C.c_double

# name -> argtypes, mirrors include/vqa_b200.h one to one
PROTOTYPES = {
    "vqa_conv_relu_pool_fwd": [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vqa_conv_bwd_data": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vqa_conv_bwd_weight": [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "vqa_dropnorm_fwd": [_vp, _vp, _vp, _vp, _i, _i64, _i, _f, _f, _u64, _vp],
    "vqa_dropnorm_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i64, _i, _f, _f, _u64, _vp],
    "vqa_dropnorm_bwd_unpool": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _f, _u64, _vp],
    "vqa_embed_tanh_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "vqa_embed_tanh_bwd": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "vqa_dropout_threshold_pattern": [_f, _vp, _vp],
    "vqa_length_order": [_vp, _vp, _vp, _i, _i, _vp],
    "vqa_embed_tanh_fwd_ordered": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "vqa_embed_tanh_bwd_ordered": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "vqa_lstm_step_fwd": [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "vqa_lstm_step_bwd_pointwise": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "vqa_gemm": [_vp, _i, _i64, _i64, _i64, _vp, _i, _i64, _i64, _i64, _vp, _i, _i64, _i64,
                 _vp, _vp, _i64, _i, _i, _i, _i, _i, _f, _u64, _u32, _vp],
    "vqa_attention_fwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "vqa_attention_bwd": [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i,
                          _i, _i, _i, _i, _i, _f, _u64, _vp],
    "vqa_attention_fwd_x": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _f, _u64, _vp],
    "vqa_attention_bwd_x": [_vp, _i64, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i,
                            _i, _i, _i, _i, _i, _f, _u64, _vp],
    "vqa_attention_streaming_ok": [_i, _i, _i, _i, _i, _i],
    "vqa_softloss_fwd_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp],
    "vqa_dropout_apply": [_vp, _i64, _vp, _i64, _i, _i64, _i, _f, _u64, _u32, _vp],
    "vqa_colsum": [_vp, _i, _i64, _vp, _vp, _i64, _i, _vp],
    "vqa_cast": [_vp, _i, _vp, _i, _i64, _vp],
    "vqa_cast_2d": [_vp, _i, _i64, _vp, _i, _i64, _i64, _i, _i, _vp],
    "vqa_relu_drop_bwd": [_vp, _vp, _vp, _i, _i64, _f, _vp],
    "vqa_add_dropped": [_vp, _i64, _vp, _i64, _vp, _i64, _i, _i64, _i, _f, _u64, _u32, _vp],
    "vqa_adam_multi": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _f, _f, _f, _f, _i, _f, _vp],
    "vqa_adam_multi_dev": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i64, _vp, _f, _f, _f, _f, _vp],
    "vqa_step_tick": [_vp, _d, _d, _d, _d, _vp],
    "vqa_scale_by_device_scalar": [_vp, _vp, _vp, _i64, _vp],
    "vqa_zero": [_vp, _i64, _vp],
    "vqa_copy": [_vp, _vp, _i64, _vp],
    "vqa_dropout_mask": [_vp, _i64, _f, _u64, _u32, _i, _vp],
}
SEED_ON_DEVICE = 1 << 63

_lib = None


class VqaLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libvqa_b200.so and set prototypes.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VqaLibraryError(
            f"{LIB_PATH} not found: build it with `make -C dl_vqa_b200/csrc` (nvcc, sm_100a). "
            "dl_vqa_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.vqa_last_error_string.restype = C.c_char_p
    lib.vqa_last_error_string.argtypes = []
    lib.vqa_abi_version.restype = C.c_int
    lib.vqa_abi_version.argtypes = []
    lib.vqa_launch_count.restype = C.c_uint64
    lib.vqa_launch_count.argtypes = []
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = C.c_int
        fn.argtypes = argtypes
    for name, argtypes in _optional_prototypes().items():
        if hasattr(lib, name):
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = argtypes
    _lib = lib
    return lib


def _optional_prototypes():
    from . import lib_tc
    return lib_tc.PROTOTYPES


_timing = None      # {"names": set, "events": [(tag, start, stop)]} while kernel timing is enabled


def enable_kernel_timing(names) -> None:
    """Record CUDA events (on the launching stream) around every call whose entry name is in `names`."""
    global _timing
    _timing = {"names": set(names), "events": []}


def collect_kernel_timing():
    """Synchronise, return {tag: (calls, total_ms)} and switch timing off."""
    global _timing
    t, _timing = _timing, None
    out = {}
    if t is None:
        return out
    torch.cuda.synchronize()
    for tag, a, b in t["events"]:
        n, ms = out.get(tag, (0, 0.0))
        out[tag] = (n + 1, ms + a.elapsed_time(b))
    return out


def call(name: str, *args, tag: Optional[str] = None) -> None:
    lib = load()
    timed = _timing is not None and ((tag or name) in _timing["names"] or name in _timing["names"])
    if timed:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
    rc = getattr(lib, name)(*args)
    if timed:
        b.record()
        _timing["events"].append((tag or name, a, b))
    if rc != 0:
        msg = lib.vqa_last_error_string().decode("utf-8", "replace")
        raise VqaLibraryError(f"{name} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().vqa_launch_count())


def dtype_code(t: torch.dtype) -> int:
    if t == torch.float32:
        return F32
    if t == torch.bfloat16:
        return BF16
    raise VqaLibraryError(f"unsupported activation dtype {t}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a CUDA tensor (None -> NULL).  Refuses host tensors: no CPU path exists."""
    if t is None:
        return None
    if not t.is_cuda:
        raise VqaLibraryError("dl_vqa_b200 kernels take CUDA tensors only (no CPU fallback)")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream
