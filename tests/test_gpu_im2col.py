"""The im2col + tcgen05-GEMM convolution layers of the tensor-core arm (csrc/im2col.cu; reference models/model.py:72-84 for any
cfg image.stride / kernel_size / num_channels, e.g. the stride-2 encoder of config/config_eval.yaml:52-62): every glue
kernel against the torch restatement of the same op (F.unfold / F.max_pool2d / F.fold), and one whole layer
(forward, weight / bias / data gradient) against torch autograd of Conv2d -> ReLU -> MaxPool2d on the same bf16-rounded
operands.  Index / byte work is held bit-exactly, the GEMM-backed results to bf16 accuracy (2e-2)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rup(x, m):
    return (x + m - 1) // m * m


def _err(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("B,IH,IW,Cin,KS,stride,nchw,dtype", [
    (2, 37, 41, 3, 3, 2, 1, torch.float32), (2, 37, 41, 3, 3, 2, 1, torch.float16), (3, 20, 17, 16, 3, 2, 0, torch.bfloat16),
    (2, 13, 13, 64, 3, 1, 0, torch.bfloat16), (1, 11, 9, 8, 5, 2, 0, torch.bfloat16), (2, 9, 9, 12, 3, 1, 0, torch.bfloat16)])
def test_im2col_matches_unfold(B, IH, IW, Cin, KS, stride, nchw, dtype):
    from dl_vqa_b200 import lib
    torch.manual_seed(B + IH + Cin)
    x = torch.randn(B, Cin, IH, IW, device="cuda").to(dtype)
    xin = x.contiguous() if nchw else x.permute(0, 2, 3, 1).contiguous()
    OH, OW = (IH - KS) // stride + 1, (IW - KS) // stride + 1
    K = KS * KS * Cin
    Kp = _rup(K, 8)
    col = torch.full((B * OH * OW, Kp), float("nan"), device="cuda", dtype=torch.bfloat16)
    code = {torch.float32: lib.F32, torch.float16: lib.F16, torch.bfloat16: lib.BF16}[dtype]
    lib.call("vqa_im2col", lib.ptr(xin), code, nchw, lib.ptr(col), B, IH, IW, Cin, KS, stride, Kp, lib.stream())
    torch.cuda.synchronize()
    # F.unfold orders columns (ci, kh, kw); the patch matrix orders them (kh, kw, ci)
    u = F.unfold(x.float(), KS, stride=stride).view(B, Cin, KS * KS, OH * OW).permute(0, 3, 2, 1).reshape(B * OH * OW, K)
    assert torch.equal(col[:, :K].float(), u.bfloat16().float())
    assert bool((col[:, K:] == 0).all())


@pytest.mark.parametrize("Cout,Cin,KS", [(64, 3, 3), (16, 8, 5), (128, 64, 3)])
def test_weight_pack_and_gradient_unpack_are_inverse_permutations(Cout, Cin, KS):
    from dl_vqa_b200 import lib
    torch.manual_seed(Cout)
    w = torch.randn(Cout, Cin, KS, KS, device="cuda")
    K = KS * KS * Cin
    Kp = _rup(K, 8)
    wp = torch.empty(Cout, Kp, device="cuda", dtype=torch.bfloat16)
    lib.call("vqa_conv_weight_pack_im2col", lib.ptr(w), lib.ptr(wp), Cout, Cin, KS, Kp, lib.stream())
    want = w.permute(0, 2, 3, 1).reshape(Cout, K).bfloat16()
    assert torch.equal(wp[:, :K], want) and bool((wp[:, K:] == 0).all())
    dwp = torch.zeros(Cout, Kp, device="cuda")
    dwp[:, :K] = w.permute(0, 2, 3, 1).reshape(Cout, K)
    dw = torch.empty_like(w)
    lib.call("vqa_conv_weight_grad_unpack_im2col", lib.ptr(dwp), lib.ptr(dw), Cout, Cin, KS, Kp, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(dw, w)


@pytest.mark.parametrize("B,OH,OW,C", [(2, 27, 27, 128), (3, 6, 7, 16), (1, 111, 111, 64), (2, 2, 2, 8)])
def test_pool_and_unpool_match_torch(B, OH, OW, C):
    from dl_vqa_b200 import lib
    torch.manual_seed(OH + C)
    y = torch.relu(torch.randn(B, OH, OW, C, device="cuda")).bfloat16()
    y[0, 0:2, 0:2, 0] = 1.5                              # a four-way tie: the first element must win
    y[0, 0:2, 0:2, 1] = 0.0                              # a dead window
    PH, PW = OH // 2, OW // 2
    out = torch.empty(B, PH, PW, C, device="cuda", dtype=torch.bfloat16)
    mask = torch.empty(B, PH, PW, C, device="cuda", dtype=torch.uint8)
    st = lib.stream()
    lib.call("vqa_pool2x2_fwd", lib.ptr(y), lib.ptr(out), lib.ptr(mask), B, OH, OW, C, st)
    yt = y.float().permute(0, 3, 1, 2)
    want, idx = F.max_pool2d(yt, 2, 2, return_indices=True)
    assert torch.equal(out.float().permute(0, 3, 1, 2), want)
    r, c = idx // OW, idx % OW
    e = torch.where(want > 0, (r % 2) * 2 + (c % 2), torch.full_like(idx, 4))
    assert torch.equal(mask.permute(0, 3, 1, 2).long(), e)
    assert int(mask[0, 0, 0, 0]) == 0 and int(mask[0, 0, 0, 1]) == 4
    # backward: autograd of relu -> max_pool2d on the same tensor
    da = torch.randn(B, PH, PW, C, device="cuda").bfloat16()
    dy = torch.full((B, OH, OW, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.call("vqa_unpool2x2_bwd", lib.ptr(da), lib.ptr(mask), lib.ptr(dy), B, OH, OW, C, st)
    torch.cuda.synchronize()
    leaf = yt.clone().requires_grad_(True)
    F.max_pool2d(leaf, 2, 2).backward(da.float().permute(0, 3, 1, 2))
    want_dy = leaf.grad * (yt > 0)                      # ReLU-dead windows pass nothing down
    assert torch.equal(dy.float().permute(0, 3, 1, 2), want_dy)


@pytest.mark.parametrize("B,IH,IW,Cin,KS,stride", [(2, 27, 27, 64, 3, 2), (3, 13, 10, 16, 3, 1), (1, 12, 12, 8, 5, 2), (2, 55, 55, 64, 3, 2)])
def test_col2im_matches_fold(B, IH, IW, Cin, KS, stride):
    from dl_vqa_b200 import lib
    torch.manual_seed(IH + Cin)
    OH, OW = (IH - KS) // stride + 1, (IW - KS) // stride + 1
    K = KS * KS * Cin
    Kp = _rup(K, 8)
    dcol = torch.randn(B * OH * OW, Kp, device="cuda").bfloat16()
    dx = torch.full((B, IH, IW, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.call("vqa_col2im", lib.ptr(dcol), lib.ptr(dx), B, IH, IW, Cin, KS, stride, Kp, lib.stream())
    torch.cuda.synchronize()
    cols = dcol[:, :K].float().view(B, OH * OW, KS * KS, Cin).permute(0, 3, 2, 1).reshape(B, Cin * KS * KS, OH * OW)
    # F.fold covers the (OH-1)*stride + KS rows the windows reach; rows beyond (floor of the conv output size) get zeros
    ch, cw = (OH - 1) * stride + KS, (OW - 1) * stride + KS
    want = torch.zeros(B, Cin, IH, IW, device="cuda")
    want[:, :, :ch, :cw] = F.fold(cols, (ch, cw), KS, stride=stride)
    assert _err(dx.float().permute(0, 3, 1, 2), want) < 1e-2          # fp32 sums of up to KS*KS bf16 values, rounded once


@pytest.mark.parametrize("B,S,Cin,Cout,stride", [(3, 55, 64, 128, 2), (2, 31, 32, 64, 1), (4, 64, 3, 64, 2), (2, 13, 128, 256, 2)])
def test_im2col_layer_forward_and_backward_match_torch_autograd(B, S, Cin, Cout, stride):
    """One encoder layer through VqaNet's own im2col branch would need a whole model; here the same call sequence
    (model.py: _run_forward / _run_backward) is issued by hand and compared with Conv2d -> ReLU -> MaxPool2d in torch."""
    from dl_vqa_b200 import lib
    torch.manual_seed(S + Cin)
    KS = 3
    st = lib.stream()
    x = torch.randn(B, Cin, S, S, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, KS, KS, device="cuda") / (Cin * 9) ** 0.5)
    b = torch.randn(Cout, device="cuda") * 0.1
    OH = (S - KS) // stride + 1
    PH = OH // 2
    K, M = KS * KS * Cin, B * OH * OH
    Kp = _rup(K, 8)
    nchw = 1 if Cin == 3 else 0
    xin = x.float().contiguous() if nchw else x.permute(0, 2, 3, 1).contiguous()
    wp = torch.empty(Cout, Kp, device="cuda", dtype=torch.bfloat16)
    col = torch.empty(M, Kp, device="cuda", dtype=torch.bfloat16)
    y = torch.empty(M, Cout, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(B, PH, PH, Cout, device="cuda", dtype=torch.bfloat16)
    mask = torch.empty(B, PH, PH, Cout, device="cuda", dtype=torch.uint8)
    lib.call("vqa_conv_weight_pack_im2col", lib.ptr(w), lib.ptr(wp), Cout, Cin, KS, Kp, st)
    lib.call("vqa_im2col", lib.ptr(xin), lib.F32 if nchw else lib.BF16, nchw, lib.ptr(col), B, S, S, Cin, KS, stride, Kp, st)
    lib.call("vqa_tc_gemm", lib.ptr(col), Kp, 0, lib.ptr(wp), Kp, 0, lib.ptr(y), lib.BF16, Cout, 0, lib.ptr(b), None, 0,
             M, Cout, Kp, 1, lib.GEMM_RELU, 0.0, 0, 0, st)
    lib.call("vqa_pool2x2_fwd", lib.ptr(y), lib.ptr(out), lib.ptr(mask), B, OH, OH, Cout, st)
    # backward
    da = (torch.randn(B, PH, PH, Cout, device="cuda") * 0.1).bfloat16()
    dy = torch.empty(M, Cout, device="cuda", dtype=torch.bfloat16)
    lib.call("vqa_unpool2x2_bwd", lib.ptr(da), lib.ptr(mask), lib.ptr(dy), B, OH, OH, Cout, st)
    dwp = torch.zeros(Cout, Kp, device="cuda")
    lib.call("vqa_tc_gemm", lib.ptr(dy), Cout, 0, lib.ptr(col), Kp, 0, lib.ptr(dwp), lib.F32, Kp, 0, None, None, 0,
             Cout, Kp, M, 1, lib.GEMM_OPERANDS_MN | (lib.GEMM_SPLITK if M >= 4096 else 0), 0.0, 0, 0, st)
    dw = torch.empty_like(w)
    lib.call("vqa_conv_weight_grad_unpack_im2col", lib.ptr(dwp), lib.ptr(dw), Cout, Cin, KS, Kp, st)
    dx = None
    if not nchw:
        dcol = torch.empty(M, Kp, device="cuda", dtype=torch.bfloat16)
        lib.call("vqa_tc_gemm", lib.ptr(dy), Cout, 0, lib.ptr(wp), Kp, 0, lib.ptr(dcol), lib.BF16, Kp, 0, None, None, 0,
                 M, Kp, Cout, 1, lib.GEMM_B_MN, 0.0, 0, 0, st)
        dx = torch.empty(B, S, S, Cin, device="cuda", dtype=torch.bfloat16)
        lib.call("vqa_col2im", lib.ptr(dcol), lib.ptr(dx), B, S, S, Cin, KS, stride, Kp, st)
    torch.cuda.synchronize()

    # torch reference on the same rounded operands, gated by the kernel's own pooling decisions
    xt = x.float().requires_grad_(True)
    wt = w.bfloat16().float().requires_grad_(True)
    bt = b.clone().requires_grad_(True)
    conv = F.conv2d(xt, wt, bt, stride=stride)
    ref = F.max_pool2d(torch.relu(conv), 2, 2)
    assert _err(out.float().permute(0, 3, 1, 2), ref) < 2e-2
    m = mask.permute(0, 3, 1, 2).long()
    win = conv[:, :, :2 * PH, :2 * PH].reshape(B, Cout, PH, 2, PH, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, Cout, PH, PH, 4)
    gated = win.gather(4, m.clamp(max=3).unsqueeze(-1)).squeeze(-1) * (m < 4)
    gated.backward(da.float().permute(0, 3, 1, 2))
    assert _err(dw, wt.grad) < 2e-2
    if dx is not None:
        assert _err(dx.float().permute(0, 3, 1, 2), xt.grad) < 2e-2
