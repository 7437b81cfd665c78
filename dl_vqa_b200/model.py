"""Drop-in `VqaNet` for OmerShubi/DL_VQA (reference models/model.py:7-67), B200-native.

Same constructor `VqaNet(cfg, embedding_tokens)`, same `forward(v, q, q_len)`, same sub-module attribute
names (`text`, `image`, `attention`, `classifier`) and the same 24 `state_dict` keys / shapes / layouts,
so `train.py`, `main.py` and existing `model.pth` checkpoints work unchanged.  The arithmetic runs in
hand-written CUDA (libvqa_b200.so, include/vqa_b200.h) through ONE autograd node: forward saves the
activations it needs, backward computes every parameter gradient in a fixed order (classifier ->
attention -> question encoder -> image encoder) so that a data-parallel wrapper can start reducing the
large text-side gradients while the convolution backward is still running (dl_vqa_b200/dp.py).

The torch.nn layer classes below are used ONLY as parameter containers (identical default init and
state_dict names to the reference); their forward() is never called.
"""
from __future__ import annotations

import os

from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import lib
from .lib import call, ptr


class _QuestionParams(nn.Module):
    """Parameters of reference questionNet (models/model.py:134-149)."""

    def __init__(self, embedding_tokens, embedding_features, lstm_features, num_lstm_layers, drop, bidirectional):
        super().__init__()
        if num_lstm_layers != 1:
            raise NotImplementedError("num_lstm_layers != 1 (the reference itself notes it needs code changes)")
        self.embedding = nn.Embedding(embedding_tokens, embedding_features, padding_idx=0)
        self.drop = nn.Dropout(drop)
        self.tanh = nn.Tanh()
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.lstm = nn.LSTM(input_size=embedding_features, hidden_size=lstm_features, num_layers=num_lstm_layers,
                                dropout=drop, bidirectional=bidirectional)

    def forward(self, q, q_len):
        """reference questionNet.forward (models/model.py:151-166) on the CUDA kernels of the owning VqaNet"""
        owner = self.__dict__.get("_owner")
        net = owner() if owner is not None else None
        if net is None:
            raise RuntimeError("question encoder parameters without an owning VqaNet")
        return net.encode_question(q, q_len)


class _ImageParams(nn.Module):
    """Parameters of reference ImageNet2 (models/model.py:72-84)."""

    def __init__(self, image_cfg):
        super().__init__()
        ch = image_cfg["num_channels"]
        for i in range(len(ch) - 1):
            self.add_module(f"conv{i}", nn.Conv2d(ch[i], ch[i + 1], kernel_size=image_cfg["kernel_size"],
                                                  stride=image_cfg["stride"]))
            self.add_module(f"relu{i}", nn.ReLU())
            self.add_module(f"maxpool{i}", nn.MaxPool2d(2, 2))
        self.add_module("drop", nn.Dropout(image_cfg["dropout"]))

    def forward(self, *a, **k):
        raise RuntimeError("parameter container only")


class _AttentionParams(nn.Module):
    """Parameters of reference Attention (models/model.py:169-181)."""

    def __init__(self, v_features, q_features, mid_features, glimpses, do_option, drop):
        super().__init__()
        self.do_option = do_option
        self.v_conv = nn.Conv2d(v_features, mid_features, kernel_size=1, bias=False)
        self.q_lin = nn.Linear(q_features, mid_features)
        self.x_conv = nn.Conv2d(2 * mid_features if do_option == "|" else mid_features, glimpses, kernel_size=1)
        self.drop = nn.Dropout(drop)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, *a, **k):
        raise RuntimeError("parameter container only")


class _ClassifierParams(nn.Module):
    """Parameters of reference Classifier (models/model.py:198-205)."""

    def __init__(self, in_features, mid_features, out_features, drop):
        super().__init__()
        self.add_module("drop1", nn.Dropout(drop))
        self.add_module("lin1", nn.Linear(in_features, mid_features))
        self.add_module("relu", nn.ReLU())
        self.add_module("drop2", nn.Dropout(drop))
        self.add_module("lin2", nn.Linear(mid_features, out_features))

    def forward(self, *a, **k):
        raise RuntimeError("parameter container only")


def _elem_stride(t0: torch.Tensor, t1: torch.Tensor) -> int:
    d = t1.data_ptr() - t0.data_ptr()
    assert d % t0.element_size() == 0
    return d // t0.element_size()


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _Math:
    """The three dense contractions of the step, on either arm.

    fp32 arm  : vqa_gemm (SIMT fp32, exact), operands addressed in place through strides.
    bf16 arm  : vqa_tc_gemm (tcgen05 / TMEM / TMA).  Weights get ONE bf16 shadow per step (same [N,K] layout), used
                K-major by the forward and MN-major (GEMM_B_MN) by the data gradient; weight gradients consume both
                operands MN-major (GEMM_OPERANDS_MN).  No transposed copies anywhere.
    """

    def __init__(self, tc: bool, dev, st, wcache=None, shadows=None):
        self.tc, self.dev, self.st = tc, dev, st
        self._w = wcache if wcache is not None else {}     # bf16 weight shadows; the backward pass reuses the forward's
        self._shadows = shadows                            # persistent shadows kept current by FusedAdam (optional)

    # ---- bf16 shadows of fp32 parameters (cached for the duration of one forward/backward)
    def wbf(self, W: torch.Tensor):
        """[N,K] fp32 -> bf16 [N,Kp], Kp = K rounded up to 8 (zero padded)."""
        key = ("n", W.data_ptr())
        if key not in self._w and self._shadows is not None:
            sh = self._shadows.get(W.data_ptr())           # (bf16 tensor, parameter, version at which it was written)
            if sh is not None and sh[1]._version == sh[2] and sh[0].device == W.device:
                self._w[key] = (sh[0], W[0].numel())
        if key not in self._w:
            N, K = W.shape[0], W[0].numel()
            Kp = _rup(K, 8)
            out = torch.empty(N, Kp, dtype=torch.bfloat16, device=self.dev)
            call("vqa_cast_2d", ptr(W), lib.F32, K, ptr(out), lib.BF16, Kp, N, K, Kp, self.st, tag="w_cast")
            self._w[key] = (out, Kp)
        return self._w[key]

    def _as_bf16(self, x_ptr, x_dt, ld, rows, cols):
        if x_dt == lib.BF16:
            return x_ptr, ld, None
        cp = _rup(cols, 8)
        buf = torch.empty(rows, cp, dtype=torch.bfloat16, device=self.dev)
        call("vqa_cast_2d", x_ptr, x_dt, ld, ptr(buf), lib.BF16, cp, rows, cols, cp, self.st, tag="act_cast")
        return buf.data_ptr(), cp, buf

    # ---- out[M,N] = act(x[M,K] W[N,K]^T + bias (+ bias2)) * dropout
    def lin_fwd(self, x_ptr, x_dt, ldx, W, out_ptr, out_dt, ldo, M, N, K, bias=None, bias2=None, relu=False,
                p=0.0, seed=0, site=0, tag=None):
        flags = lib.GEMM_RELU if relu else 0
        if self.tc:
            Wb, Kp = self.wbf(W)
            xp, ldx2, keep = self._as_bf16(x_ptr, x_dt, ldx, M, K)
            call("vqa_tc_gemm", xp, ldx2, 0, ptr(Wb), Kp, 0, out_ptr, out_dt, ldo, 0, ptr(bias), ptr(bias2), 0,
                 M, N, K, 1, flags, p, seed, site, self.st, tag=tag)
        else:
            call("vqa_gemm", x_ptr, x_dt, ldx, 1, 0, ptr(W), lib.F32, K, 1, 0, out_ptr, out_dt, ldo, 0,
                 ptr(bias), ptr(bias2), 0, M, N, K, 1, flags, p, seed, site, self.st, tag=tag)

    # ---- dx[M,K] = dy[M,N] W[N,K]
    def lin_bwd_data(self, dy_ptr, dy_dt, ldy, W, dx_ptr, dx_dt, ldd, M, N, K, tag=None):
        if self.tc:
            # W [N,K] read as stored (the forward's bf16 shadow): reduction index = its row => MN-major B operand
            Wb, Kp = self.wbf(W)
            yp, ldy2, keep = self._as_bf16(dy_ptr, dy_dt, ldy, M, N)
            call("vqa_tc_gemm", yp, ldy2, 0, ptr(Wb), Kp, 0, dx_ptr, dx_dt, ldd, 0, None, None, 0,
                 M, K, N, 1, lib.GEMM_B_MN, 0.0, 0, 0, self.st, tag=tag)
        else:
            call("vqa_gemm", dy_ptr, dy_dt, ldy, 1, 0, ptr(W), lib.F32, 1, K, 0, dx_ptr, dx_dt, ldd, 0,
                 None, None, 0, M, K, N, 1, 0, 0.0, 0, 0, self.st, tag=tag)

    # ---- dW[N,K] (fp32) = dy[M,N]^T x[M,K]      (reduction over the M rows)
    def lin_bwd_weight(self, dy_ptr, dy_dt, ldy, x_ptr, x_dt, ldx, dW, M, N, K, tag=None, zeroed=False, kblocks=None):
        """`zeroed`: dW already holds zeros (the gradient arena is cleared once per backward pass).  `kblocks` (tensor-core
        arm): device list of the 64-row reduction blocks that are not all zero in dy (vqa_lstm_active_kblocks)."""
        big = M >= 4096

        def clear():
            if not zeroed:
                call("vqa_zero", ptr(dW), dW.numel() * dW.element_size(), self.st)
        if self.tc:
            if M == 0:
                clear()
                return
            # reduction index = row of both operands: MN-major tcgen05 operands, no transposes
            yp, ldy2, keep1 = self._as_bf16(dy_ptr, dy_dt, ldy, M, N)
            xp, ldx2, keep2 = self._as_bf16(x_ptr, x_dt, ldx, M, K)
            if big:
                clear()
            if kblocks is not None:
                call("vqa_tc_gemm_kblocks", yp, ldy2, xp, ldx2, ptr(dW), K, N, K, M,
                     lib.GEMM_OPERANDS_MN | (lib.GEMM_SPLITK if big else 0), ptr(kblocks), self.st, tag=tag)
                return
            call("vqa_tc_gemm", yp, ldy2, 0, xp, ldx2, 0, ptr(dW), lib.F32, K, 0, None, None, 0,
                 N, K, M, 1, lib.GEMM_OPERANDS_MN | (lib.GEMM_SPLITK if big else 0), 0.0, 0, 0, self.st, tag=tag)
        else:
            if big:
                clear()
            call("vqa_gemm", dy_ptr, dy_dt, 1, ldy, 0, x_ptr, x_dt, 1, ldx, 0, ptr(dW), lib.F32, K, 0,
                 None, None, 0, N, K, M, 1, lib.GEMM_SPLITK if big else 0, 0.0, 0, 0, self.st, tag=tag)


class VqaNet(nn.Module):
    """Show, Ask, Attend and Answer -- B200-native drop-in for reference models/model.py:VqaNet."""

    def __init__(self, cfg, embedding_tokens, compute_dtype: Optional[str] = None):
        super().__init__()
        text_cfg, image_cfg = cfg["text"], cfg["image"]
        attention_cfg, classifier_cfg = cfg["attention"], cfg["classifier"]
        self.H = int(text_cfg["question_features"])
        self.E = int(text_cfg["embedding_features"])
        self.dirs = 2 if text_cfg["bidirectional"] else 1
        self.G = int(attention_cfg["glimpses"])
        self.A = int(attention_cfg["hidden_dim"])
        self.channels = [int(c) for c in image_cfg["num_channels"]]
        self.KS = int(image_cfg["kernel_size"])
        self.stride = int(image_cfg["stride"])
        self.do_option = attention_cfg["do_option"]
        if self.do_option not in ("+", "*", "|"):
            raise ValueError(f"attention.do_option {self.do_option!r}")
        self.hidden = int(classifier_cfg["hidden_dim"])
        self.max_answers = int(cfg["max_answers"])
        self.p_text = float(text_cfg["dropout"])
        self.p_img = float(image_cfg["dropout"])
        self.p_att = float(attention_cfg["dropout"])
        self.p_cls = float(classifier_cfg["dropout"])
        lstm_out = self.H * self.dirs

        # registration order = reference order (models/model.py:26-51): identical init under a seed
        self.text = _QuestionParams(embedding_tokens, self.E, self.H, text_cfg["num_lstm_layers"],
                                    text_cfg["dropout"], text_cfg["bidirectional"])
        self.image = _ImageParams(image_cfg)
        self.attention = _AttentionParams(self.channels[-1], lstm_out, self.A, self.G, self.do_option, self.p_att)
        self.classifier = _ClassifierParams(self.G * self.channels[-1] + lstm_out, self.hidden, self.max_answers,
                                            self.p_cls)
        import weakref
        self.text.__dict__["_owner"] = weakref.ref(self)       # not a sub-module registration: no cycle in the module tree
        self.compute_dtype = torch.float32
        if compute_dtype is not None:
            self.set_compute_dtype(compute_dtype)
        self._seed_counter = 0
        self.grad_ready_hook = None      # callable(list[(name, grad)]) fired as each stage's grads complete
        self._arena = None               # see use_gradient_arena()
        self._shadows = None             # see use_weight_shadows()
        self._whh_shadow = None
        self._step_state = None          # see use_device_step_state()

    # ------------------------------------------------------------------ configuration
    def use_device_step_state(self, state: Optional[torch.Tensor]) -> "VqaNet":
        """Take the dropout seed of every training forward from a device-resident `VqaStepState` (include/vqa_b200.h;
        a 64-byte CUDA tensor whose first 8 bytes are the seed) instead of drawing it on the host: the step can then be
        captured in a CUDA graph and still draw a fresh mask on every replay (dl_vqa_b200/graph.py).  None switches back."""
        if state is not None and (not state.is_cuda or state.numel() * state.element_size() < 64):
            raise ValueError("step state must be a CUDA tensor of at least 64 bytes")
        self._step_state = state
        return self

    def set_compute_dtype(self, dt) -> "VqaNet":
        """'float32' (exact arm, SIMT fp32) or 'bfloat16' (tensor-core arm, fp32 accumulate)."""
        if isinstance(dt, str):
            dt = {"float32": torch.float32, "fp32": torch.float32, "bfloat16": torch.bfloat16,
                  "bf16": torch.bfloat16}[dt]
        if dt not in (torch.float32, torch.bfloat16):
            raise ValueError(f"compute dtype {dt}")
        self.compute_dtype = dt
        return self

    def _params(self) -> List[nn.Parameter]:
        return list(self.parameters())

    # ------------------------------------------------------------------ gradient arena
    STAGES = ("classifier", "attention", "text", "image")      # order in which _run_backward finishes them

    def use_gradient_arena(self, enable: bool = True) -> "VqaNet":
        """Write every parameter gradient into a persistent flat fp32 buffer per stage (classifier / attention /
        text / image) instead of freshly allocated tensors.  `p.grad` then aliases the arena: gradient pointers are
        stable from step to step (FusedAdam never re-uploads its pointer table) and a data-parallel wrapper can
        all-reduce each stage's bucket IN PLACE (`dp.GradientAllReduce`) with no concatenate / scatter copies.
        Contract: call `optimizer.zero_grad(set_to_none=True)` (or consume the gradients) before the next backward;
        if a live `.grad` still aliases the arena, that backward falls back to fresh tensors so that autograd's
        accumulation stays correct."""
        self._arena = {} if enable else None
        return self

    def _arena_views(self, dev):
        """{name: view} and {stage: flat bucket}, (re)built lazily for the current device / parameter shapes."""
        a = self._arena
        if a and a.get("dev") == dev:
            return a
        named = list(self.named_parameters())
        per_stage = {st: [(n, p) for n, p in named if n.startswith(st + ".")] for st in self.STAGES}
        sizes = {st: _rup(sum(_rup(p.numel(), 4) for _, p in mine), 64) for st, mine in per_stage.items()}
        # ONE allocation (a single memset clears it at the start of every backward pass); each stage's bucket is a
        # 256-byte aligned slice of it, all-reduced on its own as soon as the stage is finished
        whole = torch.zeros(sum(sizes.values()), dtype=torch.float32, device=dev)
        buckets, views, base = {}, {}, 0
        for stage in self.STAGES:
            flat = whole[base:base + sizes[stage]]
            base += sizes[stage]
            off = 0
            for n, p in per_stage[stage]:
                views[n] = flat[off:off + p.numel()].view(p.shape)
                off += _rup(p.numel(), 4)                      # keep every view 16-byte aligned
            buckets[stage] = flat
        self._arena = {"dev": dev, "views": views, "buckets": buckets, "whole": whole}
        return self._arena

    # ------------------------------------------------------------------ persistent bf16 weight shadows
    def use_weight_shadows(self, optimizer) -> "VqaNet":
        """Tensor-core arm: let `optimizer` (a FusedAdam) write the bf16 shadow of every GEMM weight in the same kernel
        that updates the fp32 master, instead of re-casting the weights every step.  A shadow is used only while the
        parameter's autograd version equals the version recorded when it was written, so `load_state_dict`, another
        optimizer or any in-place torch op on the parameter silently falls back to the per-step cast.  Contract:
        do not modify the parameters through `.data` (that bypasses the version counter)."""
        if not hasattr(optimizer, "register_bf16_shadow"):
            raise TypeError("use_weight_shadows needs a dl_vqa_b200.FusedAdam")
        self._shadows = {}
        lstm = self.text.lstm
        weights = [self.classifier.lin1.weight, self.classifier.lin2.weight, self.attention.q_lin.weight, self.attention.v_conv.weight]
        whh = [getattr(lstm, f"weight_hh_l0{s}") for s in ["", "_reverse"][:self.dirs]]
        # the recurrent weights of both directions share one [dirs, 4H, H] buffer (the backward GEMM is batched over them)
        self._whh_shadow = torch.empty(self.dirs, 4 * self.H, self.H, dtype=torch.bfloat16, device=whh[0].device) if whh[0].is_cuda else None
        pairs = [(W, None) for W in weights] + [(W, self._whh_shadow[d] if self._whh_shadow is not None else None) for d, W in enumerate(whh)]
        for W, sh in pairs:
            if W[0].numel() % 8 != 0 or not W.is_cuda:
                continue                                   # shadows are unpadded [N,K] copies: TMA needs 16-byte row pitches
            if sh is None:
                sh = torch.empty(W.shape[0], W[0].numel(), dtype=torch.bfloat16, device=W.device)
            entry = [sh, W, -1]                            # version -1: not valid until the optimizer has written it
            self._shadows[W.data_ptr()] = entry
            optimizer.register_bf16_shadow(W, sh, entry)
        return self

    def gradient_buckets(self):
        """{stage: flat fp32 tensor} of the arena (None when the arena is off or not built yet)."""
        return self._arena.get("buckets") if self._arena else None

    def _next_seed(self) -> int:
        # host-side only (CPU generator): respects torch.manual_seed, never synchronises the device
        return int(torch.empty((), dtype=torch.int64).random_().item())

    # ------------------------------------------------------------------ forward
    def forward(self, v, q, q_len):
        if not v.is_cuda:
            raise lib.VqaLibraryError("VqaNet.forward: inputs must be CUDA tensors (no CPU fallback); call model.cuda()")
        if v.requires_grad:
            raise NotImplementedError("gradient w.r.t. the input image is not produced (the reference never needs it)")
        dev = v.device
        if isinstance(q_len, (list, tuple)):
            q_len = torch.as_tensor([int(x) for x in q_len], dtype=torch.int64)
        q_len = q_len.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        q = q.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        v = self._image_input(v)
        params = self._params()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if not self.training:
            seed = 0
        elif self._step_state is not None:       # seed lives in device memory, advanced by vqa_step_tick (graph replay)
            seed = lib.SEED_ON_DEVICE | self._step_state.data_ptr()
        else:
            seed = self._next_seed()
        if need_grad:
            return _VqaFunction.apply(self, seed, v, q, q_len, *params)
        logits, _ = self._run_forward(v, q, q_len, seed, save=False)
        return logits

    def _image_input(self, v: torch.Tensor) -> torch.Tensor:
        """The network input as the first-layer kernels take it: contiguous NCHW, float32 (the reference's dtype) or -- where
        the tcgen05 first-layer kernels run -- float16 read natively (the dtype the reference stores its pre-processed
        images in, preprocessing/preprocess_images.py:40; widening fp16 -> fp32 is exact, so results are bit-identical)."""
        if v.dtype == torch.float16 and self._conv0_reads_fp16():
            return v.contiguous()
        return v.to(torch.float32).contiguous()

    def _conv0_reads_fp16(self) -> bool:
        ch = self.channels
        direct = self.compute_dtype == torch.bfloat16 and ch[0] == 3 and ch[1] == 64 and self.KS == 3 and self.stride == 1
        return direct or self._im2col_conv_ok(0)          # both first-layer forms of the tensor-core arm read fp16 natively

    # dropout probabilities in effect
    def _p(self, p: float) -> float:
        return p if self.training else 0.0

    def _tc_conv_ok(self, i: int) -> bool:
        """tcgen05 implicit-GEMM conv is used for the bf16 arm when the layer shape allows it."""
        Cin, Cout = self.channels[i], self.channels[i + 1]
        return (self.compute_dtype == torch.bfloat16 and self.KS == 3 and self.stride == 1 and Cin % 64 == 0
                and Cout in (64, 128, 256))

    def _im2col_conv_ok(self, i: int) -> bool:
        """Layer i runs as im2col + tcgen05 GEMM + pool (bf16 arm, shapes the direct kernels do not cover)."""
        Cin, Cout = self.channels[i], self.channels[i + 1]
        return (self.compute_dtype == torch.bfloat16 and Cout % 8 == 0 and (i == 0 or Cin % 8 == 0)
                and os.environ.get("VQA_CONV_IM2COL", "1") != "0")

    def _run_forward(self, v, q, q_len, seed: int, save: bool):
        adt = self.compute_dtype
        dt = lib.dtype_code(adt)
        tc = adt == torch.bfloat16
        dev = v.device
        st = lib.stream()
        mm = _Math(tc, dev, st, shadows=self._shadows if tc else None)
        B = v.shape[0]
        f32 = torch.float32
        ctx = {} if save else None
        p_text, p_img, p_att, p_cls = self._p(self.p_text), self._p(self.p_img), self._p(self.p_att), self._p(self.p_cls)

        def empty(*shape, dtype=adt):
            return torch.empty(shape, dtype=dtype, device=dev)

        # ---------------- image encoder: fused conv+ReLU+pool per layer (models/model.py:72-84)
        assert v.shape[1] == self.channels[0], "image channel count"
        x, x_dt, nchw = v, (lib.F16 if v.dtype == torch.float16 else lib.F32), 1
        IH, IW = int(v.shape[2]), int(v.shape[3])
        conv_saved = []
        conv_wd = {}                 # layer -> weights packed for the data gradient (training forward only)
        conv_cols = {}               # layer -> (patch matrix, packed weights, Kp) of the im2col + GEMM layers
        nl = len(self.channels) - 1
        for i in range(nl):
            conv = getattr(self.image, f"conv{i}")
            Cin, Cout = self.channels[i], self.channels[i + 1]
            OH, OW = (IH - self.KS) // self.stride + 1, (IW - self.KS) // self.stride + 1
            PH, PW = OH // 2, OW // 2
            if PH <= 0 or PW <= 0:
                raise ValueError(f"image too small: layer {i} conv output {OH}x{OW}")
            out = empty(B, PH, PW, Cout)
            mask = empty(B, PH, PW, Cout, dtype=torch.uint8)
            if tc and nchw == 1 and Cin == 3 and Cout == 64 and self.KS == 3 and self.stride == 1:
                call("vqa_tc_conv0_relu_pool_fwd_x", ptr(x), x_dt, ptr(conv.weight), ptr(conv.bias), ptr(out), ptr(mask),
                     B, IH, IW, Cin, Cout, st, tag=f"conv{i}_fwd")
            elif self._tc_conv_ok(i) and nchw == 0:
                wp = empty(Cout, 9 * Cin)
                wd = empty(Cin, 9 * Cout) if (save and i > 0) else None      # data-gradient packing in the same pass
                call("vqa_pack_conv3x3_weight", ptr(conv.weight), ptr(wp), ptr(wd), Cout, Cin, st, tag="w_cast")
                if wd is not None:
                    conv_wd[i] = wd
                call("vqa_tc_conv3x3_relu_pool_fwd", ptr(x), ptr(wp), ptr(conv.bias), ptr(out), ptr(mask),
                     B, IH, IW, Cin, Cout, st, tag=f"conv{i}_fwd")
            elif self._im2col_conv_ok(i):
                # any stride / kernel size / channel list on the tensor cores: patch matrix -> tcgen05 GEMM with bias + ReLU
                # in its epilogue -> 2x2 max-pool + arg-max mask (im2col.cu); e.g. the stride-2 encoder of config_eval.yaml
                K = self.KS * self.KS * Cin
                Kp = _rup(K, 8)
                M = B * OH * OW
                wp = empty(Cout, Kp)
                call("vqa_conv_weight_pack_im2col", ptr(conv.weight), ptr(wp), Cout, Cin, self.KS, Kp, st, tag="w_cast")
                col = empty(M, Kp)
                call("vqa_im2col", ptr(x), x_dt, nchw, ptr(col), B, IH, IW, Cin, self.KS, self.stride, Kp, st, tag=f"conv{i}_fwd")
                y = empty(M, Cout)
                call("vqa_tc_gemm", ptr(col), Kp, 0, ptr(wp), Kp, 0, ptr(y), lib.BF16, Cout, 0, ptr(conv.bias), None, 0,
                     M, Cout, Kp, 1, lib.GEMM_RELU, 0.0, 0, 0, st, tag=f"conv{i}_fwd")
                call("vqa_pool2x2_fwd", ptr(y), ptr(out), ptr(mask), B, OH, OW, Cout, st, tag=f"conv{i}_fwd")
                if save:
                    conv_cols[i] = (col, wp, Kp)
            else:
                call("vqa_conv_relu_pool_fwd", ptr(x), x_dt, nchw, ptr(conv.weight), ptr(conv.bias), ptr(out), ptr(mask),
                     dt, B, IH, IW, Cin, Cout, self.KS, self.stride, st, tag=f"conv{i}_fwd")
            conv_saved.append((x, x_dt, nchw, mask, IH, IW, Cin, Cout))
            x, x_dt, nchw, IH, IW = out, dt, 0, PH, PW
        P, Cimg = IH * IW, self.channels[-1]

        # ---------------- image.drop + channel L2 norm (models/model.py:84, :56)
        vn = empty(B * P, Cimg)
        vnd = empty(B * P, Cimg) if p_att > 0 else None
        nrm = empty(B * P, dtype=f32)
        call("vqa_dropnorm_fwd", ptr(x), ptr(vn), ptr(vnd), ptr(nrm), dt, B * P, Cimg, p_img, p_att, seed, st)
        v_in = vnd if vnd is not None else vn

        # ---------------- question encoder (models/model.py:151-166)
        tx = self._text_forward(q, q_len, seed, mm, st, p_text)
        T, H, dirs, qf = tx["T"], self.H, self.dirs, tx["qf"]

        # ---------------- attention (models/model.py:183-195, :208-221)
        att = self.attention
        QF = dirs * H
        A = self.A
        qd = qf
        if p_att > 0:
            qd = empty(B, QF)
            call("vqa_dropout_apply", ptr(qf), QF, ptr(qd), QF, dt, B, QF, p_att, seed, lib.SITE_ATT_Q, st)
        qp = empty(B, A, dtype=f32)
        mm.lin_fwd(ptr(qd), dt, QF, att.q_lin.weight, ptr(qp), lib.F32, A, B, A, QF, bias=att.q_lin.bias, tag="q_lin")
        G = self.G
        op = {"+": lib.ATT_ADD, "*": lib.ATT_MUL, "|": lib.ATT_CAT}[self.do_option]
        # v' is handed to the streaming attention kernels as FLOAT16 (written so by the v_conv GEMM): the '+' fusion rounds
        # v' + q' to the 16-bit format, and 11 mantissa bits instead of 8 take the x_conv weight gradient from 1-2 % to
        # ~0.3 % max-norm error (attention.cu); |v'| <= ||W_row|| is nowhere near the fp16 range and the GEMM saturates
        vp_f16 = tc and bool(lib.load().vqa_attention_streaming_ok(dt, op, P, A, Cimg, G)) and os.environ.get("VQA_ATT_VP_F16", "1") != "0"
        vp_dt = lib.F16 if vp_f16 else dt
        vp = empty(B * P, A, dtype=torch.float16 if vp_f16 else adt)
        mm.lin_fwd(ptr(v_in), dt, Cimg, att.v_conv.weight, ptr(vp), vp_dt, A, B * P, A, Cimg, tag="v_conv")
        KC = G * Cimg + QF
        comb = empty(B, KC)
        prob = empty(B, G, P, dtype=f32)
        call("vqa_attention_fwd_x", ptr(vp), vp_dt, ptr(qp), ptr(vn), ptr(att.x_conv.weight), ptr(att.x_conv.bias),
             ptr(prob), ptr(comb), KC, dt, op, B, P, A, Cimg, G, p_att, seed, st, tag="vqa_attention_fwd")
        # combined = cat([pooled, q])  (models/model.py:64)
        esz = comb.element_size()
        call("vqa_dropout_apply", ptr(qf), QF, comb.data_ptr() + G * Cimg * esz, KC, dt, B, QF, 0.0, 0, 0, st)

        # ---------------- classifier (models/model.py:198-205)
        cl = self.classifier
        combd = comb
        if p_cls > 0:
            combd = empty(B, KC)
            call("vqa_dropout_apply", ptr(comb), KC, ptr(combd), KC, dt, B, KC, p_cls, seed, lib.SITE_CLS_IN, st)
        h1d = empty(B, self.hidden)
        mm.lin_fwd(ptr(combd), dt, KC, cl.lin1.weight, ptr(h1d), dt, self.hidden, B, self.hidden, KC,
                   bias=cl.lin1.bias, relu=True, p=p_cls, seed=seed, site=lib.SITE_CLS_HID, tag="lin1")
        logits = empty(B, self.max_answers, dtype=f32)
        mm.lin_fwd(ptr(h1d), dt, self.hidden, cl.lin2.weight, ptr(logits), lib.F32, self.max_answers,
                   B, self.max_answers, self.hidden, bias=cl.lin2.bias, tag="lin2")

        if save:
            ctx.update(wcache=mm._w, B=B, P=P, T=T, seed=seed, conv_saved=conv_saved, conv_wd=conv_wd, conv_cols=conv_cols, vn=vn, v_in=v_in, nrm=nrm, a_last=x,
                       text=tx, qd=qd, qp=qp, vp=vp, prob=prob, combd=combd, h1d=h1d,
                       q=q, q_len=q_len,
                       p=(p_text, p_img, p_att, p_cls), dt=dt, adt=adt)
        return logits, ctx

    # ------------------------------------------------------------------ question encoder (reference questionNet)
    def _text_forward(self, q, q_len, seed: int, mm: "_Math", st, p_text: float) -> dict:
        """models/model.py:151-166: embedding -> dropout -> tanh -> (bi)LSTM, final CELL state [B, dirs*H].
        Returns the output `qf` and everything the backward needs."""
        adt = self.compute_dtype
        dt = lib.dtype_code(adt)
        tc = adt == torch.bfloat16
        dev = q.device
        f32 = torch.float32
        B = int(q.shape[0])

        def empty(*shape, dtype=adt):
            return torch.empty(shape, dtype=dtype, device=dev)

        T, E, H, dirs = int(q.shape[1]), self.E, self.H, self.dirs
        lstm = self.text.lstm
        sfx = ["", "_reverse"][:dirs]
        w_ih = [getattr(lstm, f"weight_ih_l0{s}") for s in sfx]
        w_hh = [getattr(lstm, f"weight_hh_l0{s}") for s in sfx]
        b_ih = [getattr(lstm, f"bias_ih_l0{s}") for s in sfx]
        b_hh = [getattr(lstm, f"bias_hh_l0{s}") for s in sfx]
        ldx = _rup(E, 8) if tc else E
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        persistent = tc and H % 64 == 0 and H <= 1024 and dirs * (H // 16) <= sms
        # Length order (what pack_padded_sequence(enforce_sorted=False) computes on the host, models/model.py:160): on the
        # persistent kernels the rows of every step-indexed buffer are the samples in descending length order, so whole
        # 128-row tiles leave the recurrence once their longest question has ended.  qf comes back in sample order.
        ordered = (persistent and T > 1 and B <= 8192 and self._persistent_lstm_bwd_ok(B, T, sms)
                   and os.environ.get("VQA_LSTM_ORDER", "1") != "0")
        order = len_rows = None
        if ordered:
            order = torch.empty(B, dtype=torch.int32, device=dev)
            len_rows = torch.empty(B, dtype=torch.int64, device=dev)
            call("vqa_length_order", ptr(q_len), ptr(order), ptr(len_rows), B, T, st, tag="embed_fwd")
        xs = empty(dirs, T, B, ldx)
        call("vqa_embed_tanh_fwd_ordered", ptr(q), ptr(q_len), ptr(order), ptr(self.text.embedding.weight), ptr(xs), dt,
             B, T, E, ldx, dirs, p_text, seed, st, tag="embed_fwd")
        gx = empty(dirs, T, B, 4 * H)
        for d in range(dirs):   # hoisted input projection: x W_ih^T + b_ih + b_hh for all steps at once
            mm.lin_fwd(ptr(xs[d]), dt, ldx, w_ih[d], ptr(gx[d]), dt, 4 * H, T * B, 4 * H, E,
                       bias=b_ih[d], bias2=b_hh[d], tag="lstm_inproj")
        cs = empty(dirs, T, B, H, dtype=f32)
        qf = empty(B, dirs * H)
        whh_stride = _elem_stride(w_hh[0], w_hh[1]) if dirs == 2 else 0
        if persistent:
            # one cooperative launch for all steps and directions; W_hh resident in shared memory
            wp = empty(dirs, 4 * H, H)
            for d in range(dirs):
                call("vqa_pack_lstm_whh", ptr(w_hh[d]), ptr(wp[d]), H, st, tag="w_cast")
            hs_ext = torch.empty(dirs, T + 1, B, H, dtype=adt, device=dev)      # the kernel writes slots 1..T of every row
            for d in range(dirs):                                               # slot 0 = h_{-1} = 0
                call("vqa_zero", ptr(hs_ext[d, 0]), B * H * hs_ext.element_size(), st)
            sync = torch.empty(dirs, dtype=torch.int32, device=dev)            # cleared by the entry itself
            call("vqa_tc_lstm_fwd_ordered", ptr(gx), ptr(cs), ptr(hs_ext), ptr(qf), ptr(wp), ptr(len_rows if ordered else q_len),
                 ptr(order), ptr(sync), T, B, H, dirs, st, tag="lstm_recurrence_fwd")
            h_prev = [hs_ext[d, 1] for d in range(dirs)]     # h_0 .. h_{T-1} of direction d start here
            hs = hs_ext
        else:
            hs = empty(dirs, T, B, H)
            for s in range(T):
                call("vqa_lstm_step_fwd", ptr(gx), ptr(cs), ptr(hs), ptr(qf), ptr(w_hh[0]), whh_stride, ptr(q_len),
                     dt, s, T, B, H, dirs, st, tag="lstm_step_fwd")
            h_prev = [hs[d, 0] for d in range(dirs)]
        return dict(T=T, B=B, xs=xs, gx=gx, cs=cs, hs=hs, h_prev=h_prev, qf=qf, ldx=ldx, whh_stride=whh_stride,
                    q=q, q_len=q_len, seed=seed, p_text=p_text, order=order, len_rows=len_rows)

    def _persistent_lstm_bwd_ok(self, B: int, T: int, sms: int) -> bool:
        """Shapes the one-launch backward recurrence (vqa_tc_lstm_bwd) covers; decided in the forward as well, because the
        length-ordered row layout is only understood by the two persistent kernels."""
        H, dirs = self.H, self.dirs
        return (self.compute_dtype == torch.bfloat16 and T > 1 and H % 128 == 0
                and ((B + 127) // 128) * (H // 128) * dirs <= sms and os.environ.get("VQA_LSTM_BWD_PERSISTENT", "1") != "0")

    def _text_backward(self, tx: dict, dqf, mm: "_Math", st, galloc, colsum, zeroed: bool, after_recurrence=None,
                       duplicate_bias: bool = True) -> dict:
        """BPTT through _text_forward (only c_n feeds the model, models/model.py:164-166).  `dqf` [B, dirs*H] is the gradient
        w.r.t. the final cell states.  `after_recurrence()` is called once the serial part is enqueued (the data-parallel
        wrapper starts its first all-reduces there).  Returns {state_dict key: gradient} of the text parameters."""
        adt = self.compute_dtype
        dt = lib.dtype_code(adt)
        tc = adt == torch.bfloat16
        dev = dqf.device
        f32 = torch.float32
        H, E, dirs = self.H, self.E, self.dirs
        T, B, seed, p_text = tx["T"], tx["B"], tx["seed"], tx["p_text"]
        grads = {}

        def empty(*shape, dtype=adt):
            return torch.empty(shape, dtype=dtype, device=dev)

        def zeros(*shape, dtype=f32):
            t = torch.empty(shape, dtype=dtype, device=dev)
            call("vqa_zero", ptr(t), t.numel() * t.element_size(), st)
            return t

        lstm = self.text.lstm
        sfx = ["", "_reverse"][:dirs]
        w_ih = [getattr(lstm, f"weight_ih_l0{s}") for s in sfx]
        w_hh = [getattr(lstm, f"weight_hh_l0{s}") for s in sfx]
        gx, cs, hs, xs, ldx = tx["gx"], tx["cs"], tx["hs"], tx["xs"], tx["ldx"]
        q, q_len = tx["q"], tx["q_len"]
        dh = zeros(dirs, B, H)
        dc = empty(dirs, B, H, dtype=f32)
        dg = empty(dirs, T, B, 4 * H)
        gsz = dg.element_size()
        if tc and T > 1:
            # bf16 shadow of W_hh as stored [dirs, 4H, H]: MN-major B operand of dh = dg W_hh
            shs = [mm.wbf(w_hh[d])[0] for d in range(dirs)]          # Adam-maintained shadows when valid, else one cast each
            base = getattr(self, "_whh_shadow", None)
            if base is not None and all(shs[d].data_ptr() == base[d].data_ptr() for d in range(dirs)):
                whhb = base
            else:
                whhb = empty(dirs, 4 * H, H)
                for d in range(dirs):
                    call("vqa_copy", ptr(whhb[d]), ptr(shs[d]), 4 * H * H * 2, st)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        order, len_rows = tx.get("order"), tx.get("len_rows")
        # rows in length order exist only where the forward saw that this backward would be the persistent one
        persistent_bwd = order is not None or self._persistent_lstm_bwd_ok(B, T, sms)
        if persistent_bwd:
            # all T steps and both directions in one cooperative launch (pointwise + split-K tcgen05 GEMM per step,
            # two point-to-point synchronisations per step) instead of 2T - 1 dependent launches
            sync_b = torch.empty(256, dtype=torch.int32, device=dev)           # cleared by the entry itself
            call("vqa_tc_lstm_bwd_ordered", ptr(gx), ptr(cs), ptr(dh), ptr(dc), ptr(dqf), ptr(dg), ptr(whhb),
                 ptr(len_rows if order is not None else q_len), ptr(order), ptr(sync_b),
                 T, B, H, dirs, st, tag="lstm_bwd_persistent")
        for s in (range(T - 1, -1, -1) if not persistent_bwd else ()):
            call("vqa_lstm_step_bwd_pointwise", ptr(gx), ptr(cs), ptr(dh), ptr(dc),
                 ptr(dqf) if s == T - 1 else None, ptr(dg), ptr(q_len), dt, s, T, B, H, dirs, st, tag="lstm_bwd_pointwise")
            if s > 0:   # dh_{s-1} = dgates_s W_hh
                if tc:
                    # split-K with vector reductions into dh, which the pointwise kernel above has just cleared:
                    # M = B is small, so an unsplit GEMM leaves most SMs idle and each CTA ingests all of K
                    call("vqa_tc_gemm", dg.data_ptr() + s * B * 4 * H * gsz, 4 * H, T * B * 4 * H, ptr(whhb), H,
                         4 * H * H, ptr(dh), lib.F32, H, B * H, None, None, 0, B, H, 4 * H, dirs,
                         lib.GEMM_SPLITK | lib.GEMM_B_MN, 0.0, 0, 0, st, tag="lstm_step_bwd")
                else:
                    call("vqa_gemm", dg.data_ptr() + s * B * 4 * H * gsz, dt, 4 * H, 1, T * B * 4 * H,
                         ptr(w_hh[0]), lib.F32, 1, H, tx["whh_stride"], ptr(dh), lib.F32, H, B * H,
                         None, None, 0, B, H, 4 * H, dirs, 0, 0.0, 0, 0, st, tag="lstm_step_bwd")
        if after_recurrence is not None:
            after_recurrence()
        # dg is exactly zero at (step, row) positions past the end of a question (what pack_padded_sequence drops,
        # models/model.py:160): the weight-gradient reductions skip every 64-row block that holds only such rows -- with the
        # rows in length order about 40 % of them at B = 256, half at B = 1024
        kb_ih = kb_hh = None
        if tc and T > 1 and B % 64 == 0 and B <= 8192 and os.environ.get("VQA_LSTM_KSKIP", "1") != "0":
            kb_ih = torch.empty(1 + T * (B // 64), dtype=torch.int32, device=dev)
            kb_hh = torch.empty(1 + (T - 1) * (B // 64), dtype=torch.int32, device=dev)
            call("vqa_lstm_active_kblocks", ptr(len_rows if order is not None else q_len), ptr(kb_ih), ptr(kb_hh), B, T, st,
                 tag="lstm_whh_wgrad")
        for d in range(dirs):
            dWhh = galloc(f"text.lstm.weight_hh_l0{sfx[d]}", 4 * H, H)
            mm.lin_bwd_weight(dg[d].data_ptr() + B * 4 * H * gsz, dt, 4 * H, ptr(tx["h_prev"][d]), dt, H, dWhh,
                              (T - 1) * B, 4 * H, H, tag="lstm_whh_wgrad", zeroed=zeroed, kblocks=kb_hh)
            dWih = galloc(f"text.lstm.weight_ih_l0{sfx[d]}", 4 * H, E)
            mm.lin_bwd_weight(ptr(dg[d]), dt, 4 * H, ptr(xs[d]), dt, ldx, dWih, T * B, 4 * H, E, tag="lstm_wih_wgrad", zeroed=zeroed,
                              kblocks=kb_ih)
            db = colsum(dg[d], dt, 4 * H, T * B, 4 * H, f"text.lstm.bias_ih_l0{sfx[d]}")
            grads[f"text.lstm.weight_hh_l0{sfx[d]}"] = dWhh
            grads[f"text.lstm.weight_ih_l0{sfx[d]}"] = dWih
            grads[f"text.lstm.bias_ih_l0{sfx[d]}"] = db
            if duplicate_bias:           # b_ih and b_hh always enter as a sum: identical gradients, separate storage
                db2 = galloc(f"text.lstm.bias_hh_l0{sfx[d]}", 4 * H)
                call("vqa_copy", ptr(db2), ptr(db), 4 * H * 4, st)
            else:
                db2 = db
            grads[f"text.lstm.bias_hh_l0{sfx[d]}"] = db2
        dxs = empty(dirs, T, B, ldx)
        for d in range(dirs):
            mm.lin_bwd_data(ptr(dg[d]), dt, 4 * H, w_ih[d], ptr(dxs[d]), dt, ldx, T * B, 4 * H, E, tag="lstm_inproj_dgrad")
        demb = galloc("text.embedding.weight", *self.text.embedding.weight.shape, zero=True)
        call("vqa_embed_tanh_bwd_ordered", ptr(q), ptr(q_len), ptr(order), ptr(xs), ptr(dxs), ptr(demb), dt, B, T, E, ldx, dirs,
             p_text, seed, st, tag="embed_bwd")
        grads["text.embedding.weight"] = demb
        return grads

    def encode_question(self, q, q_len):
        """The question encoder alone -- reference `questionNet.forward(q, q_len)` (models/model.py:151-166), which is also
        what `model.text(q, q_len)` computes: q int64 [B,T] zero padded, q_len [B] (tensor or list) -> [B, dirs*H] in the
        compute dtype, differentiable w.r.t. the text parameters.  BASELINE.json configs[3] measures this stage."""
        if not q.is_cuda:
            raise lib.VqaLibraryError("encode_question: inputs must be CUDA tensors (no CPU fallback)")
        dev = q.device
        if isinstance(q_len, (list, tuple)):
            q_len = torch.as_tensor([int(x) for x in q_len], dtype=torch.int64)
        q_len = q_len.to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        q = q.to(dtype=torch.int64).contiguous()
        params = list(self.text.parameters())
        if not self.training:
            seed = 0
        elif self._step_state is not None:
            seed = lib.SEED_ON_DEVICE | self._step_state.data_ptr()
        else:
            seed = self._next_seed()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _TextFunction.apply(self, seed, q, q_len, *params)
        st = lib.stream()
        mm = _Math(self.compute_dtype == torch.bfloat16, dev, st, shadows=self._shadows if self.compute_dtype == torch.bfloat16 else None)
        return self._text_forward(q, q_len, seed, mm, st, self._p(self.p_text))["qf"]

    # ------------------------------------------------------------------ backward
    def _run_backward(self, ctx, dlogits: torch.Tensor, before_image=None):
        """Returns {state_dict key: gradient}.  Order: classifier, attention, text, image.  `before_image()` is called when
        the classifier / attention / text gradients (98 % of the bytes) are complete and only the convolution backward
        remains (GraphedTrainStep splits its CUDA graph there and starts the all-reduce of those buckets)."""
        B, P, T, seed = ctx["B"], ctx["P"], ctx["T"], ctx["seed"]
        p_text, p_img, p_att, p_cls = ctx["p"]
        dt, adt = ctx["dt"], ctx["adt"]
        tc = adt == torch.bfloat16
        dev = dlogits.device
        st = lib.stream()
        mm = _Math(tc, dev, st, ctx.get("wcache"), shadows=self._shadows if tc else None)
        f32 = torch.float32
        H, E, dirs, G, A = self.H, self.E, self.dirs, self.G, self.A
        Cimg = self.channels[-1]
        QF = dirs * H
        KC = G * Cimg + QF
        N = self.max_answers
        hid = self.hidden
        grads = {}

        def empty(*shape, dtype=adt):
            return torch.empty(shape, dtype=dtype, device=dev)

        def zeros(*shape, dtype=f32):
            t = torch.empty(shape, dtype=dtype, device=dev)
            call("vqa_zero", ptr(t), t.numel() * t.element_size(), st)
            return t

        deferred = []

        def fire(names, defer=False):
            """Hand finished gradients to the data-parallel hook.  `defer`: hold them back until flush_deferred() -- the
            cooperative LSTM kernels need (nearly) every SM to themselves, and an all-reduce that is already running on the
            communication stream would make their launch wait for it."""
            if self.grad_ready_hook is None:
                return
            if defer:
                deferred.append(list(names))
            else:
                self.grad_ready_hook([(n, grads[n]) for n in names])

        def flush_deferred():
            while deferred:
                fire(deferred.pop(0))

        arena = None
        if self._arena is not None:
            arena = self._arena_views(dev)["views"]
            owned = {b.untyped_storage().data_ptr() for b in self._arena["buckets"].values()}
            for p_ in self.parameters():                 # a live .grad still aliasing the arena: accumulate safely
                if p_.grad is not None and p_.grad.untyped_storage().data_ptr() in owned:
                    arena = None
                    break
            if arena is not None:        # every accumulate-into gradient (split-K sums, column sums, embedding scatter) starts at 0
                whole = self._arena["whole"]
                call("vqa_zero", ptr(whole), whole.numel() * 4, st)
        zeroed = arena is not None

        def galloc(name, *shape, zero=False):
            """fp32 gradient tensor of parameter `name`: a view of the stage arena, or a fresh tensor"""
            if arena is not None:
                t = arena[name]
                assert tuple(t.shape) == tuple(shape) or t.numel() == int(torch.Size(shape).numel()), name
                return t.view(*shape)                       # cleared above, together with the whole arena
            return zeros(*shape) if zero else empty(*shape, dtype=f32)

        def colsum(src, src_dt, ld, rows, cols, name):
            out = galloc(name, cols, zero=True)
            call("vqa_colsum", ptr(src), src_dt, ld, None, ptr(out), rows, cols, st)
            return out

        dlogits = dlogits.to(f32).contiguous()
        cl, att = self.classifier, self.attention
        h1d, combd = ctx["h1d"], ctx["combd"]

        # ---- classifier.lin2
        dh1d = empty(B, hid)
        mm.lin_bwd_data(ptr(dlogits), lib.F32, N, cl.lin2.weight, ptr(dh1d), dt, hid, B, N, hid, tag="lin2_dgrad")
        dW2 = galloc("classifier.lin2.weight", N, hid)
        mm.lin_bwd_weight(ptr(dlogits), lib.F32, N, ptr(h1d), dt, hid, dW2, B, N, hid, tag="lin2_wgrad", zeroed=zeroed)
        grads["classifier.lin2.weight"] = dW2
        grads["classifier.lin2.bias"] = colsum(dlogits, lib.F32, N, B, N, "classifier.lin2.bias")
        # ---- classifier.lin1 (ReLU + drop2 folded: h1d > 0 <=> unit alive and kept)
        dz1 = empty(B, hid)
        call("vqa_relu_drop_bwd", ptr(dh1d), ptr(h1d), ptr(dz1), dt, B * hid, p_cls, st)
        dcomb = empty(B, KC)
        mm.lin_bwd_data(ptr(dz1), dt, hid, cl.lin1.weight, ptr(dcomb), dt, KC, B, hid, KC, tag="lin1_dgrad")
        dW1 = galloc("classifier.lin1.weight", hid, KC)
        mm.lin_bwd_weight(ptr(dz1), dt, hid, ptr(combd), dt, KC, dW1, B, hid, KC, tag="lin1_wgrad", zeroed=zeroed)
        grads["classifier.lin1.weight"] = dW1
        grads["classifier.lin1.bias"] = colsum(dz1, dt, hid, B, hid, "classifier.lin1.bias")
        fire(["classifier.lin2.weight", "classifier.lin2.bias", "classifier.lin1.weight", "classifier.lin1.bias"], defer=True)
        if p_cls > 0:   # through classifier.drop1 (in place)
            call("vqa_dropout_apply", ptr(dcomb), KC, ptr(dcomb), KC, dt, B, KC, p_cls, seed, lib.SITE_CLS_IN, st)

        # ---- fused attention backward
        vp, qp, vn, v_in, prob, qd = ctx["vp"], ctx["qp"], ctx["vn"], ctx["v_in"], ctx["prob"], ctx["qd"]
        dvp = empty(B * P, A)
        dvn_pool = empty(B * P, Cimg)
        dqp = empty(B, A, dtype=f32)
        AW = 2 * A if self.do_option == "|" else A            # x_conv input channels (models/model.py:175-178)
        dwx_part = empty(B, G * AW, dtype=f32)
        dbx_part = empty(B, G, dtype=f32)
        op = {"+": lib.ATT_ADD, "*": lib.ATT_MUL, "|": lib.ATT_CAT}[self.do_option]
        call("vqa_attention_bwd_x", ptr(dcomb), KC, ptr(vp), lib.F16 if vp.dtype == torch.float16 else dt, ptr(qp), ptr(vn),
             ptr(att.x_conv.weight), ptr(prob), ptr(dvp), ptr(dvn_pool), ptr(dqp), ptr(dwx_part), ptr(dbx_part), dt, op,
             B, P, A, Cimg, G, p_att, seed, st, tag="vqa_attention_bwd")
        grads["attention.x_conv.weight"] = colsum(dwx_part, lib.F32, G * AW, B, G * AW, "attention.x_conv.weight").view(G, AW, 1, 1)
        grads["attention.x_conv.bias"] = colsum(dbx_part, lib.F32, G, B, G, "attention.x_conv.bias")
        # ---- attention.v_conv (1x1 conv == GEMM over B*P rows)
        dvnd = empty(B * P, Cimg)
        mm.lin_bwd_data(ptr(dvp), dt, A, att.v_conv.weight, ptr(dvnd), dt, Cimg, B * P, A, Cimg, tag="v_conv_dgrad")
        dWv = galloc("attention.v_conv.weight", A, Cimg)
        mm.lin_bwd_weight(ptr(dvp), dt, A, ptr(v_in), dt, Cimg, dWv, B * P, A, Cimg, tag="v_conv_wgrad", zeroed=zeroed)
        grads["attention.v_conv.weight"] = dWv.view(A, Cimg, 1, 1)
        # ---- attention.q_lin
        dqd = empty(B, QF)
        mm.lin_bwd_data(ptr(dqp), lib.F32, A, att.q_lin.weight, ptr(dqd), dt, QF, B, A, QF, tag="q_lin_dgrad")
        dWq = galloc("attention.q_lin.weight", A, QF)
        mm.lin_bwd_weight(ptr(dqp), lib.F32, A, ptr(qd), dt, QF, dWq, B, A, QF, tag="q_lin_wgrad", zeroed=zeroed)
        grads["attention.q_lin.weight"] = dWq
        grads["attention.q_lin.bias"] = colsum(dqp, lib.F32, A, B, A, "attention.q_lin.bias")
        fire(["attention.v_conv.weight", "attention.q_lin.weight", "attention.q_lin.bias",
              "attention.x_conv.weight", "attention.x_conv.bias"], defer=True)
        # gradient w.r.t. the question feature: concat branch + (dropped) q_lin branch
        dqf = empty(B, QF)
        esz = dcomb.element_size()
        call("vqa_add_dropped", dcomb.data_ptr() + G * Cimg * esz, KC, ptr(dqd), QF, ptr(dqf), QF, dt, B, QF,
             p_att, seed, lib.SITE_ATT_Q, st)

        # ---- question encoder: BPTT (only c_n feeds the model, models/model.py:164-166)
        tgrads = self._text_backward(ctx["text"], dqf, mm, st, galloc, colsum, zeroed, after_recurrence=flush_deferred,
                                     duplicate_bias=(arena is not None or self.grad_ready_hook is not None))
        grads.update(tgrads)
        sfx = ["", "_reverse"][:dirs]
        names = ["text.embedding.weight"]
        for s_ in sfx:
            names += [f"text.lstm.weight_ih_l0{s_}", f"text.lstm.weight_hh_l0{s_}", f"text.lstm.bias_ih_l0{s_}",
                      f"text.lstm.bias_hh_l0{s_}"]
        fire(names)
        if before_image is not None:
            before_image()

        # ---- image encoder
        nl = len(self.channels) - 1

        def tc_path(i):
            """layer i runs the tcgen05 un-pool + wgrad (+ dgrad) kernels with the bias gradient fused into the un-pool"""
            _, _, nchw_i, _, _, _, Cin_i, Cout_i = ctx["conv_saved"][i]
            return (self._tc_conv_ok(i) and nchw_i == 0 and Cin_i in (64, 128) and Cout_i % 128 == 0 and Cout_i <= 256)

        dy_ready = None          # (un-pooled gradient, bias gradient) of the layer about to be processed, when a producer fused it
        if tc and tc_path(nl - 1) and Cimg % 8 == 0 and Cimg <= 256:
            # L2-norm / dropout backward fused with the last layer's max-pool backward and bias gradient
            _, _, _, mask_l, IH_l, IW_l, _, _ = ctx["conv_saved"][nl - 1]
            PH_l, PW_l = ((IH_l - self.KS) // self.stride + 1) // 2, ((IW_l - self.KS) // self.stride + 1) // 2
            dy_l = empty(B, 2 * PH_l, 2 * PW_l, Cimg)
            db_l = galloc(f"image.conv{nl - 1}.bias", Cimg)
            call("vqa_dropnorm_bwd_unpool", ptr(dvn_pool), ptr(dvnd), ptr(vn), ptr(ctx["nrm"]), ptr(mask_l), ptr(dy_l), ptr(db_l),
                 B, PH_l, PW_l, Cimg, p_img, p_att, seed, st, tag="unpool")
            dy_ready = (dy_l, db_l)
            da = None
        else:
            da = empty(B * P, Cimg)
            call("vqa_dropnorm_bwd", ptr(dvn_pool), ptr(dvnd), ptr(vn), ptr(ctx["nrm"]), ptr(da), dt, B * P, Cimg,
                 p_img, p_att, seed, st)
        names = []
        for i in range(nl - 1, -1, -1):
            conv = getattr(self.image, f"conv{i}")
            x, x_dt, nchw, mask, IH, IW, Cin, Cout = ctx["conv_saved"][i]
            PH, PW = ((IH - self.KS) // self.stride + 1) // 2, ((IW - self.KS) // self.stride + 1) // 2
            dW = galloc(f"image.conv{i}.weight", *conv.weight.shape)
            use_tc = self._tc_conv_ok(i) and nchw == 0
            dy = None
            fused_db = False
            if dy_ready is not None:         # the producer of this layer's gradient already un-pooled it and summed the bias gradient
                dy, db = dy_ready
                dy_ready, fused_db = None, True
            else:
                db = galloc(f"image.conv{i}.bias", Cout)
            tc0 = tc and nchw == 1 and Cin == 3 and Cout == 64 and self.KS == 3 and self.stride == 1
            if tc0:     # fused un-pool + weight gradient + bias gradient straight from (dpool, mask): no dY tensor
                call("vqa_tc_conv0_bwd_weight_bias_x", ptr(x), x_dt, ptr(da), ptr(mask), ptr(dW), ptr(db), B, IH, IW, Cin, Cout, st,
                     tag=f"conv{i}_wgrad")
            elif use_tc and dy is None:   # un-pooled gradient, shared by the weight and the data gradient
                dy = empty(B, 2 * PH, 2 * PW, Cout)
                fused_db = use_tc and Cin in (64, 128) and Cout % 128 == 0 and Cout <= 256
                call("vqa_unpool_bf16", ptr(da), ptr(mask), ptr(dy), ptr(db) if fused_db else None, B, PH, PW, Cout, st,
                     tag="unpool")
            cols = ctx.get("conv_cols", {}).get(i)
            if tc0:
                pass
            elif cols is not None:
                # im2col layer: un-pool -> weight gradient (reduction-major tcgen05 GEMM over the saved patch matrix) and data
                # gradient (GEMM against the packed weights as stored, then col2im)
                col, wp, Kp = cols
                OH, OW = (IH - self.KS) // self.stride + 1, (IW - self.KS) // self.stride + 1
                M = B * OH * OW
                dy = empty(M, Cout)
                call("vqa_unpool2x2_bwd", ptr(da), ptr(mask), ptr(dy), B, OH, OW, Cout, st, tag="unpool")
                call("vqa_zero", ptr(db), db.numel() * 4, st)
                call("vqa_colsum", ptr(da), dt, Cout, ptr(mask), ptr(db), B * PH * PW, Cout, st)
                dwp = zeros(Cout, Kp)
                call("vqa_tc_gemm", ptr(dy), Cout, 0, ptr(col), Kp, 0, ptr(dwp), lib.F32, Kp, 0, None, None, 0,
                     Cout, Kp, M, 1, lib.GEMM_OPERANDS_MN | (lib.GEMM_SPLITK if M >= 4096 else 0), 0.0, 0, 0, st, tag=f"conv{i}_wgrad")
                call("vqa_conv_weight_grad_unpack_im2col", ptr(dwp), ptr(dW), Cout, Cin, self.KS, Kp, st, tag=f"conv{i}_wgrad")
            elif use_tc and Cin in (64, 128) and Cout % 128 == 0:
                call("vqa_tc_conv3x3_bwd_weight", ptr(x), ptr(dy), ptr(dW), B, IH, IW, Cin, Cout, st,
                     tag=f"conv{i}_wgrad")
                if not fused_db:
                    call("vqa_zero", ptr(db), db.numel() * 4, st)
                    call("vqa_colsum", ptr(da), dt, Cout, ptr(mask), ptr(db), B * PH * PW, Cout, st)
            else:
                call("vqa_conv_bwd_weight", ptr(x), x_dt, nchw, ptr(da), ptr(mask), ptr(dW), ptr(db), dt,
                     B, IH, IW, Cin, Cout, self.KS, self.stride, st, tag=f"conv{i}_wgrad")
            grads[f"image.conv{i}.weight"] = dW
            grads[f"image.conv{i}.bias"] = db
            names += [f"image.conv{i}.weight", f"image.conv{i}.bias"]
            if i > 0:
                dx = None
                if cols is not None:
                    col, wp, Kp = cols
                    dcol = empty(M, Kp)
                    call("vqa_tc_gemm", ptr(dy), Cout, 0, ptr(wp), Kp, 0, ptr(dcol), lib.BF16, Kp, 0, None, None, 0,
                         M, Kp, Cout, 1, lib.GEMM_B_MN, 0.0, 0, 0, st, tag=f"conv{i}_dgrad")
                    dx = empty(B, IH, IW, Cin)
                    call("vqa_col2im", ptr(dcol), ptr(dx), B, IH, IW, Cin, self.KS, self.stride, Kp, st, tag=f"conv{i}_dgrad")
                    del dy, dcol
                elif use_tc:
                    wd = ctx.get("conv_wd", {}).get(i)
                    if wd is None:
                        wd = empty(Cin, 9 * Cout)
                        call("vqa_pack_conv3x3_weight", ptr(conv.weight), None, ptr(wd), Cout, Cin, st, tag="w_cast")
                    if tc_path(i - 1):
                        # the data gradient lands directly in the layer below's UN-POOLED gradient (its max-pool backward
                        # and bias gradient run in the dgrad epilogue): no dx tensor, no un-pool kernel
                        dy_b = empty(B, 2 * IH, 2 * IW, Cin)
                        db_b = galloc(f"image.conv{i - 1}.bias", Cin)
                        call("vqa_tc_conv3x3_bwd_data_unpool", ptr(dy), ptr(wd), ptr(ctx["conv_saved"][i - 1][3]), ptr(dy_b),
                             ptr(db_b), B, IH, IW, Cin, Cout, st, tag=f"conv{i}_dgrad")
                        dy_ready = (dy_b, db_b)
                    else:
                        dx = empty(B, IH, IW, Cin)
                        call("vqa_tc_conv3x3_bwd_data", ptr(dy), ptr(wd), ptr(dx), B, IH, IW, Cin, Cout, st,
                             tag=f"conv{i}_dgrad")
                    del dy
                else:
                    dx = empty(B, IH, IW, Cin)
                    call("vqa_conv_bwd_data", ptr(da), ptr(mask), ptr(conv.weight), ptr(dx), dt,
                         B, IH, IW, Cin, Cout, self.KS, self.stride, st, tag=f"conv{i}_dgrad")
                da = dx
        fire(names)
        return grads


class _VqaFunction(torch.autograd.Function):
    """One autograd node for the whole network (forward saves, backward = VqaNet._run_backward)."""

    @staticmethod
    def forward(ctx, model: VqaNet, seed: int, v, q, q_len, *params):
        logits, saved = model._run_forward(v, q, q_len, seed, save=True)
        ctx.model = model
        ctx.saved = saved
        ctx.names = [n for n, _ in model.named_parameters()]
        ctx.params = params
        ctx.param_versions = [p._version for p in params]
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        model = ctx.model
        if ctx.saved is None:
            raise RuntimeError("VqaNet: backward through the same forward a second time (saved activations were freed)")
        # backward re-reads the live fp32 weights and their optimizer-maintained bf16 shadows: an update between forward and
        # backward would silently mix two parameter versions (torch raises for its own ops; so must this node)
        for n, p, ver in zip(ctx.names, ctx.params, ctx.param_versions):
            if p._version != ver:
                raise RuntimeError(f"VqaNet: parameter {n} was modified in place between forward and backward "
                                   f"(version {ver} -> {p._version}); run backward before optimizer.step() / load_state_dict()")
        grads = model._run_backward(ctx.saved, dlogits)
        ctx.saved = None
        ctx.params = None
        out = []
        for n, need in zip(ctx.names, ctx.needs_input_grad[5:]):
            out.append(grads[n] if need else None)
        return (None, None, None, None, None, *out)


class _TextFunction(torch.autograd.Function):
    """Autograd node of VqaNet.encode_question (the question encoder alone)."""

    @staticmethod
    def forward(ctx, model: VqaNet, seed: int, q, q_len, *params):
        st = lib.stream()
        tc = model.compute_dtype == torch.bfloat16
        mm = _Math(tc, q.device, st, shadows=model._shadows if tc else None)
        tx = model._text_forward(q, q_len, seed, mm, st, model._p(model.p_text))
        ctx.model, ctx.tx, ctx.wcache = model, tx, mm._w
        ctx.names = ["text." + n for n, _ in model.text.named_parameters()]
        ctx.params, ctx.param_versions = params, [p._version for p in params]
        return tx["qf"]

    @staticmethod
    def backward(ctx, dqf):
        model = ctx.model
        if ctx.tx is None:
            raise RuntimeError("encode_question: backward through the same forward a second time")
        for n, p, ver in zip(ctx.names, ctx.params, ctx.param_versions):
            if p._version != ver:
                raise RuntimeError(f"VqaNet: parameter {n} was modified in place between forward and backward")
        st = lib.stream()
        dev = dqf.device
        tc = model.compute_dtype == torch.bfloat16
        mm = _Math(tc, dev, st, ctx.wcache, shadows=model._shadows if tc else None)

        def galloc(name, *shape, zero=False):
            t = torch.empty(shape, dtype=torch.float32, device=dev)
            if zero:
                call("vqa_zero", ptr(t), t.numel() * 4, st)
            return t

        def colsum(src, src_dt, ld, rows, cols, name):
            out = galloc(name, cols, zero=True)
            call("vqa_colsum", ptr(src), src_dt, ld, None, ptr(out), rows, cols, st)
            return out

        dqf = dqf.to(model.compute_dtype).contiguous()
        grads = model._text_backward(ctx.tx, dqf, mm, st, galloc, colsum, False, duplicate_bias=False)
        ctx.tx = None
        ctx.params = None
        return (None, None, None, None, *[grads[n] if need else None for n, need in zip(ctx.names, ctx.needs_input_grad[4:])])
