"""tcgen05 / TMA kernels against plain torch fp32 references of the same op (bf16 inputs, fp32 accumulate)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _tc_gemm(A, B, out_dtype=torch.float32, bias=None, relu=False, splitk=False, nbatch=1):
    from dl_vqa_b200 import lib
    if nbatch == 1:
        M, K = A.shape; N = B.shape[0]
        a_sb = b_sb = c_sb = 0
    else:
        _, M, K = A.shape; N = B.shape[1]
        a_sb, b_sb, c_sb = A.stride(0), B.stride(0), M * N
    C = (torch.zeros if splitk else torch.empty)((nbatch, M, N) if nbatch > 1 else (M, N), dtype=out_dtype, device="cuda")
    flags = (lib.GEMM_RELU if relu else 0) | (lib.GEMM_SPLITK if splitk else 0)
    lib.call("vqa_tc_gemm", lib.ptr(A), A.stride(-2), a_sb, lib.ptr(B), B.stride(-2), b_sb, lib.ptr(C),
             lib.dtype_code(out_dtype), N, c_sb, lib.ptr(bias), None, 0, M, N, K, nbatch, flags, 0.0, 0, 0, lib.stream())
    torch.cuda.synchronize()
    return C


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 256), (300, 200, 320), (77, 3000, 1024),
                                   (1000, 64, 72), (4096, 1024, 304), (130, 40, 8), (20000, 1024, 256), (19050, 768, 136)])
def test_tc_gemm_matches_torch(M, N, K):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda").bfloat16()
    B = torch.randn(N, K, device="cuda").bfloat16()
    want = A.float() @ B.float().t()
    got = _tc_gemm(A, B)
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 1e-5, err
    bias = torch.randn(N, device="cuda")
    got = _tc_gemm(A, B, out_dtype=torch.bfloat16, bias=bias, relu=True)
    want2 = torch.relu(want + bias)
    err = float((got.float() - want2).abs().max() / want2.abs().max())
    assert err < 1e-2, err


def test_tc_gemm_splitk_and_batched():
    torch.manual_seed(0)
    A = torch.randn(256, 8192, device="cuda").bfloat16()
    B = torch.randn(192, 8192, device="cuda").bfloat16()
    want = A.float() @ B.float().t()
    got = _tc_gemm(A, B, splitk=True)
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5
    A = torch.randn(2, 200, 128, device="cuda").bfloat16()
    B = torch.randn(2, 136, 128, device="cuda").bfloat16()
    want = torch.einsum("zmk,znk->zmn", A.float(), B.float())
    got = _tc_gemm(A, B, nbatch=2)
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5


def test_transpose_bf16():
    from dl_vqa_b200 import lib
    src = torch.randn(3, 70, 45, device="cuda")
    dst = torch.empty(3, 45, 72, dtype=torch.bfloat16, device="cuda").fill_(7)
    lib.call("vqa_transpose_bf16", lib.ptr(src), lib.F32, 45, 70 * 45, lib.ptr(dst), 72, 45 * 72, 70, 45, 3, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(dst[:, :, :70], src.transpose(1, 2).bfloat16())
    assert float(dst[:, :, 70:].float().min()) == 7.0


def _conv_case(B, IH, IW, Cin, Cout, seed):
    torch.manual_seed(seed)
    x = torch.randn(B, IH, IW, Cin, device="cuda").bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5))
    bias = torch.randn(Cout, device="cuda") * 0.1
    return x, w, bias


@pytest.fixture
def conv_cta_group(request):
    """Runs a conv test with single-CTA MMAs (1) or CTA pairs (2) and restores the default afterwards."""
    from dl_vqa_b200 import lib
    lib.call("vqa_tc_conv_set_cta_group", request.param)
    yield request.param
    lib.call("vqa_tc_conv_set_cta_group", lib.DEFAULT_CONV_CTA_GROUP)


@pytest.mark.parametrize("conv_cta_group", [1, 2], indirect=True)
@pytest.mark.parametrize("B,IH,IW,Cin,Cout", [(2, 20, 36, 64, 128), (3, 111, 111, 64, 128), (2, 54, 54, 128, 256),
                                              (1, 11, 9, 64, 64), (5, 30, 31, 128, 64), (1, 20, 12, 256, 256),
                                              (1, 7, 40, 192, 128)])
def test_tc_conv_fwd_matches_torch(B, IH, IW, Cin, Cout, conv_cta_group):
    import torch.nn.functional as F
    from dl_vqa_b200 import lib
    x, w, bias = _conv_case(B, IH, IW, Cin, Cout, IH + Cin)
    PH, PW = (IH - 2) // 2, (IW - 2) // 2
    wp = torch.empty(Cout, 9 * Cin, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_pack_conv3x3_weight", lib.ptr(w), lib.ptr(wp), None, Cout, Cin, lib.stream())
    out = torch.empty(B, PH, PW, Cout, dtype=torch.bfloat16, device="cuda")
    mask = torch.empty(B, PH, PW, Cout, dtype=torch.uint8, device="cuda")
    lib.call("vqa_tc_conv3x3_relu_pool_fwd", lib.ptr(x), lib.ptr(wp), lib.ptr(bias), lib.ptr(out), lib.ptr(mask),
             B, IH, IW, Cin, Cout, lib.stream())
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    pre = F.conv2d(x.float().permute(0, 3, 1, 2), w.bfloat16().float(), bias)
    want, idx = F.max_pool2d(torch.relu(pre), 2, 2, return_indices=True)
    want = want.permute(0, 2, 3, 1)
    err = float((out.float() - want).abs().max() / want.abs().max())
    assert err < 1e-2, err
    # arg-max agreement where the maximum is positive and unambiguous
    OW = IW - 2
    idx = idx.permute(0, 2, 3, 1)
    oh, ow = idx // OW, idx % OW
    e = (oh % 2) * 2 + (ow % 2)
    alive = want > 1e-3
    agree = float(((mask.long() == e) | ~alive).float().mean())
    assert agree > 0.999, agree
    assert bool(((mask == 4) == (out == 0)).all())


@pytest.mark.parametrize("conv_cta_group", [1, 2], indirect=True)
@pytest.mark.parametrize("B,IH,IW,Cin,Cout", [(2, 20, 36, 64, 128), (2, 111, 111, 64, 128), (2, 54, 54, 128, 256),
                                              (1, 13, 10, 64, 64), (3, 19, 9, 256, 192), (1, 40, 8, 64, 256)])
def test_tc_conv_dgrad_matches_torch(B, IH, IW, Cin, Cout, conv_cta_group):
    import torch.nn.functional as F
    from dl_vqa_b200 import lib
    torch.manual_seed(IH)
    PH, PW = (IH - 2) // 2, (IW - 2) // 2
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cout ** 0.5))
    dpool = torch.randn(B, PH, PW, Cout, device="cuda").bfloat16()
    mask = torch.randint(0, 5, (B, PH, PW, Cout), device="cuda", dtype=torch.uint8)
    dy = torch.empty(B, 2 * PH, 2 * PW, Cout, dtype=torch.bfloat16, device="cuda")
    fused_db = 256 % (Cout // 8) == 0
    db = torch.full((Cout,), 5.0, device="cuda")
    lib.call("vqa_unpool_bf16", lib.ptr(dpool), lib.ptr(mask), lib.ptr(dy), lib.ptr(db) if fused_db else None,
             B, PH, PW, Cout, lib.stream())
    # reference un-pool
    ref = torch.zeros(B, 2 * PH, 2 * PW, Cout, device="cuda")
    for e in range(4):
        ref[:, e // 2::2, e % 2::2, :] = torch.where(mask == e, dpool.float(), torch.zeros_like(dpool.float()))
    torch.cuda.synchronize()
    assert torch.equal(dy.float(), ref)
    if fused_db:                                     # fused conv-bias gradient
        want_db = ref.sum(dim=(0, 1, 2))
        assert float((db - want_db).abs().max()) < 1e-3 * float(want_db.abs().max() + 1)
    wd = torch.empty(Cin, 9 * Cout, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_pack_conv3x3_weight", lib.ptr(w), None, lib.ptr(wd), Cout, Cin, lib.stream())
    dx = torch.empty(B, IH, IW, Cin, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_tc_conv3x3_bwd_data", lib.ptr(dy), lib.ptr(wd), lib.ptr(dx), B, IH, IW, Cin, Cout, lib.stream())
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    full = torch.zeros(B, IH - 2, IW - 2, Cout, device="cuda")
    full[:, :2 * PH, :2 * PW] = ref
    want = F.conv_transpose2d(full.permute(0, 3, 1, 2), w.bfloat16().float()).permute(0, 2, 3, 1)
    err = float((dx.float() - want).abs().max() / want.abs().max())
    assert err < 1e-2, err


@pytest.mark.parametrize("conv_cta_group", [1, 2], indirect=True)
@pytest.mark.parametrize("B,IH,IW,Cin,Cout", [(2, 54, 54, 128, 256), (3, 20, 36, 64, 128), (1, 13, 10, 256, 64),
                                              (2, 17, 40, 128, 192)])
def test_tc_conv_dgrad_fused_unpool_equals_dgrad_then_unpool(B, IH, IW, Cin, Cout, conv_cta_group):
    """vqa_tc_conv3x3_bwd_data_unpool == vqa_tc_conv3x3_bwd_data followed by vqa_unpool_bf16 of the layer below
    (bit-identical un-pooled gradient; bias gradient up to fp32 summation order)."""
    from dl_vqa_b200 import lib
    torch.manual_seed(IH * 7 + Cin)
    OH, OW = ((IH - 2) // 2) * 2, ((IW - 2) // 2) * 2
    w = (torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cout ** 0.5))
    dy = torch.randn(B, OH, OW, Cout, device="cuda").bfloat16()
    mask_below = torch.randint(0, 5, (B, IH, IW, Cin), device="cuda", dtype=torch.uint8)
    wd = torch.empty(Cin, 9 * Cout, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_pack_conv3x3_weight", lib.ptr(w), None, lib.ptr(wd), Cout, Cin, lib.stream())
    dx = torch.empty(B, IH, IW, Cin, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_tc_conv3x3_bwd_data", lib.ptr(dy), lib.ptr(wd), lib.ptr(dx), B, IH, IW, Cin, Cout, lib.stream())
    want = torch.zeros(B, 2 * IH, 2 * IW, Cin, device="cuda")
    for e in range(4):
        want[:, e // 2::2, e % 2::2, :] = torch.where(mask_below == e, dx.float(), torch.zeros_like(dx.float()))
    got = torch.full((B, 2 * IH, 2 * IW, Cin), 3.0, dtype=torch.bfloat16, device="cuda")
    db = torch.full((Cin,), 5.0, device="cuda")
    lib.call("vqa_tc_conv3x3_bwd_data_unpool", lib.ptr(dy), lib.ptr(wd), lib.ptr(mask_below), lib.ptr(got), lib.ptr(db),
             B, IH, IW, Cin, Cout, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(got.float(), want)
    want_db = want.double().sum(dim=(0, 1, 2))
    assert float((db.double() - want_db).abs().max()) < 1e-4 * float(want_db.abs().max() + 1)


@pytest.mark.parametrize("B,PH,PW,C", [(3, 26, 26, 256), (2, 5, 9, 128), (1, 3, 3, 64)])
@pytest.mark.parametrize("p_img,p_att", [(0.0, 0.0), (0.3, 0.2)])
def test_dropnorm_bwd_fused_unpool_equals_the_two_kernels(B, PH, PW, C, p_img, p_att):
    """vqa_dropnorm_bwd_unpool == vqa_dropnorm_bwd followed by vqa_unpool_bf16 (+ fused bias gradient)."""
    from dl_vqa_b200 import lib
    torch.manual_seed(PH + C)
    R = B * PH * PW
    v = torch.randn(R, C, device="cuda").bfloat16()
    vn = torch.empty_like(v)
    vnd = torch.empty_like(v)
    nrm = torch.empty(R, device="cuda")
    seed = 1234
    lib.call("vqa_dropnorm_fwd", lib.ptr(v), lib.ptr(vn), lib.ptr(vnd), lib.ptr(nrm), lib.BF16, R, C, p_img, p_att, seed,
             lib.stream())
    dvn = torch.randn(R, C, device="cuda").bfloat16()
    dvnd = torch.randn(R, C, device="cuda").bfloat16()
    mask = torch.randint(0, 5, (B, PH, PW, C), device="cuda", dtype=torch.uint8)
    da = torch.empty(R, C, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_dropnorm_bwd", lib.ptr(dvn), lib.ptr(dvnd), lib.ptr(vn), lib.ptr(nrm), lib.ptr(da), lib.BF16, R, C,
             p_img, p_att, seed, lib.stream())
    want = torch.zeros(B, 2 * PH, 2 * PW, C, device="cuda")
    daf = da.float().view(B, PH, PW, C)
    for e in range(4):
        want[:, e // 2::2, e % 2::2, :] = torch.where(mask == e, daf, torch.zeros_like(daf))
    got = torch.full((B, 2 * PH, 2 * PW, C), 3.0, dtype=torch.bfloat16, device="cuda")
    db = torch.full((C,), 5.0, device="cuda")
    lib.call("vqa_dropnorm_bwd_unpool", lib.ptr(dvn), lib.ptr(dvnd), lib.ptr(vn), lib.ptr(nrm), lib.ptr(mask), lib.ptr(got),
             lib.ptr(db), B, PH, PW, C, p_img, p_att, seed, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(got.float(), want)
    want_db = want.double().sum(dim=(0, 1, 2))
    assert float((db.double() - want_db).abs().max()) < 1e-4 * float(want_db.abs().max() + 1)


@pytest.mark.parametrize("B,IH,IW,Cin,Cout", [(2, 20, 36, 64, 128), (3, 111, 111, 64, 128), (2, 54, 54, 128, 256),
                                              (1, 13, 10, 64, 128)])
def test_tc_conv_wgrad_matches_torch(B, IH, IW, Cin, Cout):
    from dl_vqa_b200 import lib
    torch.manual_seed(IW)
    PH, PW = (IH - 2) // 2, (IW - 2) // 2
    x = torch.randn(B, IH, IW, Cin, device="cuda").bfloat16()
    dpool = torch.randn(B, PH, PW, Cout, device="cuda").bfloat16()
    mask = torch.randint(0, 5, (B, PH, PW, Cout), device="cuda", dtype=torch.uint8)
    dy = torch.empty(B, 2 * PH, 2 * PW, Cout, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_unpool_bf16", lib.ptr(dpool), lib.ptr(mask), lib.ptr(dy), None, B, PH, PW, Cout, lib.stream())
    dw = torch.empty(Cout, Cin, 3, 3, device="cuda")
    lib.call("vqa_tc_conv3x3_bwd_weight", lib.ptr(x), lib.ptr(dy), lib.ptr(dw), B, IH, IW, Cin, Cout, lib.stream())
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    full = torch.zeros(B, IH - 2, IW - 2, Cout, device="cuda")
    full[:, :2 * PH, :2 * PW] = dy.float()
    want = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, 3, 3), full.permute(0, 3, 1, 2))
    err = float((dw - want).abs().max() / want.abs().max())
    assert err < 2e-3, err


def test_layout_helpers():
    from dl_vqa_b200 import lib
    B, IH, IW, Cin, Cout = 2, 20, 36, 64, 128
    PH, PW = (IH - 2) // 2, (IW - 2) // 2
    x = torch.randn(B, IH, IW, Cin, device="cuda").bfloat16()
    dpool = torch.randn(B, PH, PW, Cout, device="cuda").bfloat16()
    mask = torch.randint(0, 5, (B, PH, PW, Cout), device="cuda", dtype=torch.uint8)
    IWp, OWpp = (IW + 7) // 8 * 8, (2 * PW + 7) // 8 * 8
    xT = torch.full((B, Cin, IH, IWp), 9.0, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_nhwc_to_nchw_pad_bf16", lib.ptr(x), lib.ptr(xT), B, IH, IW, Cin, IWp, lib.stream())
    dyT = torch.full((B, Cout, 2 * PH, OWpp), 9.0, dtype=torch.bfloat16, device="cuda")
    lib.call("vqa_unpool_nchw_bf16", lib.ptr(dpool), lib.ptr(mask), lib.ptr(dyT), B, PH, PW, Cout, OWpp, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(xT[..., :IW], x.permute(0, 3, 1, 2)) and float(xT[..., IW:].float().abs().sum()) == 0
    ref = torch.zeros(B, 2 * PH, 2 * PW, Cout, device="cuda")
    for e in range(4):
        ref[:, e // 2::2, e % 2::2, :] = torch.where(mask == e, dpool.float(), torch.zeros_like(dpool.float()))
    assert torch.equal(dyT[..., :2 * PW].float(), ref.permute(0, 3, 1, 2)) and float(dyT[..., 2 * PW:].float().abs().sum()) == 0


@pytest.mark.parametrize("B,IH,IW", [(2, 38, 38), (3, 224, 224), (1, 19, 45), (150, 36, 70)])
def test_tc_conv0_fwd_and_wgrad_match_torch(B, IH, IW):
    import torch.nn.functional as F
    from dl_vqa_b200 import lib
    torch.manual_seed(IH)
    Cin, Cout = 3, 64
    x = torch.randn(B, Cin, IH, IW, device="cuda").half().float()
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / 5
    bias = torch.randn(Cout, device="cuda") * 0.1
    PH, PW = (IH - 2) // 2, (IW - 2) // 2
    out = torch.empty(B, PH, PW, Cout, dtype=torch.bfloat16, device="cuda")
    mask = torch.empty(B, PH, PW, Cout, dtype=torch.uint8, device="cuda")
    lib.call("vqa_tc_conv0_relu_pool_fwd", lib.ptr(x), lib.ptr(w), lib.ptr(bias), lib.ptr(out), lib.ptr(mask),
             B, IH, IW, Cin, Cout, lib.stream())
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    pre = F.conv2d(x.bfloat16().float(), w.bfloat16().float(), bias)
    want, idx = F.max_pool2d(torch.relu(pre), 2, 2, return_indices=True)
    want = want.permute(0, 2, 3, 1)
    err = float((out.float() - want).abs().max() / want.abs().max())
    assert err < 1e-2, err
    OW = IW - 2
    idx = idx.permute(0, 2, 3, 1)
    e = ((idx // OW) % 2) * 2 + ((idx % OW) % 2)
    agree = float(((mask.long() == e) | ~(want > 1e-3)).float().mean())
    assert agree > 0.999, agree
    # weight + bias gradient, fused with the un-pooling (no dY tensor)
    dpool = torch.randn(B, PH, PW, Cout, device="cuda").bfloat16()
    dw = torch.full((Cout, Cin, 3, 3), 7.0, device="cuda")
    db = torch.full((Cout,), 7.0, device="cuda")
    lib.call("vqa_tc_conv0_bwd_weight_bias", lib.ptr(x), lib.ptr(dpool), lib.ptr(mask), lib.ptr(dw), lib.ptr(db),
             B, IH, IW, Cin, Cout, lib.stream())
    torch.cuda.synchronize()
    dy = torch.zeros(B, 2 * PH, 2 * PW, Cout, device="cuda")          # reference un-pool
    for e in range(4):
        dy[:, e // 2::2, e % 2::2, :] = torch.where(mask == e, dpool.float(), torch.zeros_like(dpool.float()))
    full = torch.zeros(B, IH - 2, IW - 2, Cout, device="cuda")
    full[:, :2 * PH, :2 * PW] = dy
    wantw = torch.nn.grad.conv2d_weight(x.bfloat16().float(), (Cout, Cin, 3, 3), full.permute(0, 3, 1, 2))
    errw = float((dw - wantw).abs().max() / wantw.abs().max())
    assert errw < 2e-3, errw
    wantb = dy.sum(dim=(0, 1, 2))
    errb = float((db - wantb).abs().max() / wantb.abs().max())
    assert errb < 2e-3, errb


def test_tc_conv0_argmax_arithmetic_ties_and_dead_windows():
    """The first-layer epilogue finds the arg-max ARITHMETICALLY (saturated differences on the FMA pipe, conv0_tc.cu).
    Pin its corner cases on inputs where every 2x2 window is an exact four-way tie: the first element must win (id 0,
    torch's max_pool2d order), ReLU-dead windows must carry id 4 and an exactly zero output, and a window that is tied
    in its first two elements only must still report the first."""
    from dl_vqa_b200 import lib
    B, IH, IW, Cin, Cout = 2, 38, 70, 3, 64
    PH, PW = (IH - 2) // 2, (IW - 2) // 2
    x = torch.full((B, Cin, IH, IW), 0.5, device="cuda")
    w = torch.rand(Cout, Cin, 3, 3, device="cuda") + 0.1
    w[Cout // 2:] *= -1.0                                        # second half of the channels: negative everywhere -> dead
    bias = torch.zeros(Cout, device="cuda")
    bias[:4] = torch.tensor([0.25, -0.125, 1.0, 3.0], device="cuda")
    out = torch.empty(B, PH, PW, Cout, dtype=torch.bfloat16, device="cuda")
    mask = torch.full((B, PH, PW, Cout), 9, dtype=torch.uint8, device="cuda")
    lib.call("vqa_tc_conv0_relu_pool_fwd", lib.ptr(x), lib.ptr(w), lib.ptr(bias), lib.ptr(out), lib.ptr(mask),
             B, IH, IW, Cin, Cout, lib.stream())
    torch.cuda.synchronize()
    alive = slice(0, Cout // 2)
    dead = slice(Cout // 2, Cout)
    assert int((mask[..., alive] != 0).sum()) == 0               # four-way ties: first element
    assert float(out[..., alive].float().min()) > 0
    assert int((mask[..., dead] != 4).sum()) == 0                # all four elements negative: ReLU-dead
    assert float(out[..., dead].float().abs().max()) == 0.0
    want = 0.5 * w[:Cout // 2].bfloat16().float().sum(dim=(1, 2, 3)) + bias[:Cout // 2]
    got = out[0, 0, 0, alive].float()
    assert float((got - want).abs().max() / want.abs().max()) < 1e-2
    # partial tie: lower the second row of every window's source pixels so that only elements 0 and 1 tie
    x2 = x.clone()
    x2[:, :, 1::2, :] = 0.25                                     # odd input rows smaller: elements with dy = 1 see two small rows
    lib.call("vqa_tc_conv0_relu_pool_fwd", lib.ptr(x2), lib.ptr(w), lib.ptr(bias), lib.ptr(out), lib.ptr(mask),
             B, IH, IW, Cin, Cout, lib.stream())
    torch.cuda.synchronize()
    import torch.nn.functional as F
    pre = F.conv2d(x2.bfloat16().float(), w.bfloat16().float(), bias)
    _, idx = F.max_pool2d(torch.relu(pre), 2, 2, return_indices=True)
    OW = IW - 2
    idx = idx.permute(0, 2, 3, 1)
    e = ((idx // OW) % 2) * 2 + ((idx % OW) % 2)
    assert torch.equal(mask[..., alive].long(), e[..., alive])


@pytest.mark.parametrize("R,N,K", [(256, 3000, 1024), (5888, 4096, 304), (1000, 64, 72), (4, 136, 40), (20000, 1024, 256)])
def test_tc_gemm_mn_major_weight_gradient_form(R, N, K):
    """dW[N,K] = dY[R,N]^T X[R,K] with both operands consumed row-major (reduction index = row)."""
    from dl_vqa_b200 import lib
    torch.manual_seed(R)
    dY = torch.randn(R, N, device="cuda").bfloat16()
    X = torch.randn(R, K, device="cuda").bfloat16()
    big = R >= 4096
    dW = (torch.zeros if big else torch.empty)(N, K, device="cuda")
    flags = lib.GEMM_OPERANDS_MN | (lib.GEMM_SPLITK if big else 0)
    lib.call("vqa_tc_gemm", lib.ptr(dY), N, 0, lib.ptr(X), K, 0, lib.ptr(dW), lib.F32, K, 0, None, None, 0,
             N, K, R, 1, flags, 0.0, 0, 0, lib.stream())
    torch.cuda.synchronize()
    want = dY.float().t() @ X.float()
    err = float((dW - want).abs().max() / want.abs().max())
    assert err < 1e-4, err


@pytest.mark.parametrize("M,N,K,splitk", [(256, 1024, 4096, True), (256, 2560, 1024, False), (173056, 256, 1024, False),
                                          (5888, 300, 4096, False), (77, 72, 200, False), (300, 3000, 136, False)])
def test_tc_gemm_b_mn_major_data_gradient_form(M, N, K, splitk):
    """dX[M,N] = dY[M,K] W[K,N] with W consumed as stored (GEMM_B_MN): the data-gradient form of nn.Linear."""
    from dl_vqa_b200 import lib
    torch.manual_seed(M + N)
    dY = torch.randn(M, K, device="cuda").bfloat16()
    Np = (N + 7) // 8 * 8
    W = torch.zeros(K, Np, device="cuda", dtype=torch.bfloat16)
    W[:, :N] = (torch.randn(K, N, device="cuda") / K ** 0.5).bfloat16()
    out = (torch.zeros if splitk else torch.empty)(M, Np, device="cuda")
    flags = lib.GEMM_B_MN | (lib.GEMM_SPLITK if splitk else 0)
    lib.call("vqa_tc_gemm", lib.ptr(dY), K, 0, lib.ptr(W), Np, 0, lib.ptr(out), lib.F32, Np, 0, None, None, 0,
             M, N, K, 1, flags, 0.0, 0, 0, lib.stream())
    torch.cuda.synchronize()
    want = dY.float() @ W[:, :N].float()
    err = float((out[:, :N] - want).abs().max() / want.abs().max())
    assert err < 1e-4, err


@pytest.mark.parametrize("B,T,H,dirs", [(256, 23, 1024, 2), (5, 4, 64, 1), (130, 7, 256, 2), (600, 9, 256, 2)])
def test_persistent_lstm_matches_stepwise_kernel(B, T, H, dirs):
    """The cooperative tcgen05 recurrence against the per-step SIMT kernel (itself parity-checked against the
    oracle in test_gpu_parity) on identical bf16 inputs."""
    from dl_vqa_b200 import lib
    torch.manual_seed(B + T)
    dev = "cuda"
    gx0 = (torch.randn(dirs, T, B, 4 * H, device=dev) * 0.8).bfloat16()
    w_hh = torch.randn(dirs, 4 * H, H, device=dev) / H ** 0.5
    w_hh_b = w_hh.bfloat16().float().contiguous()            # same rounded weights on both sides
    q_len = torch.randint(1, T + 1, (B,), device=dev)
    q_len[0] = T
    st = lib.stream()
    # reference: step kernel
    gx_a = gx0.clone(); cs_a = torch.empty(dirs, T, B, H, device=dev); hs_a = torch.empty(dirs, T, B, H, device=dev, dtype=torch.bfloat16)
    qf_a = torch.empty(B, dirs * H, device=dev, dtype=torch.bfloat16)
    for s in range(T):
        lib.call("vqa_lstm_step_fwd", lib.ptr(gx_a), lib.ptr(cs_a), lib.ptr(hs_a), lib.ptr(qf_a), lib.ptr(w_hh_b), 4 * H * H,
                 lib.ptr(q_len), lib.BF16, s, T, B, H, dirs, st)
    # persistent kernel
    gx_b = gx0.clone(); cs_b = torch.empty(dirs, T, B, H, device=dev)
    hs_b = torch.zeros(dirs, T + 1, B, H, device=dev, dtype=torch.bfloat16)
    qf_b = torch.empty(B, dirs * H, device=dev, dtype=torch.bfloat16)
    wp = torch.empty(dirs, 4 * H, H, device=dev, dtype=torch.bfloat16)
    for d in range(dirs):
        lib.call("vqa_pack_lstm_whh", lib.ptr(w_hh[d]), lib.ptr(wp[d]), H, st)
    sync = torch.zeros(dirs, dtype=torch.int32, device=dev)
    lib.call("vqa_tc_lstm_fwd", lib.ptr(gx_b), lib.ptr(cs_b), lib.ptr(hs_b), lib.ptr(qf_b), lib.ptr(wp), lib.ptr(q_len),
             lib.ptr(sync), T, B, H, dirs, st)
    torch.cuda.synchronize()
    def err(a, b):
        return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-9))
    assert err(qf_b, qf_a) < 2e-2
    assert err(cs_b, cs_a) < 2e-2
    assert err(hs_b[:, 1:], hs_a) < 2e-2
    # activated gates only where the step was active
    act = (torch.arange(T, device=dev)[None, :, None] < q_len[None, None, :]).unsqueeze(-1)
    assert err(torch.where(act, gx_b.float(), torch.zeros((), device=dev)), torch.where(act, gx_a.float(), torch.zeros((), device=dev))) < 2e-2


def test_persistent_lstm_cluster_multicast_variant_in_fresh_process():
    """The opt-in cluster / TMA-multicast form of the persistent LSTM (VQA_LSTM_CLUSTER=4 is read at the first launch,
    hence a fresh interpreter): same parity test, and the library must report that clusters were really used."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, VQA_LSTM_CLUSTER="4")
    if env.get("VQA_LSTM_CLUSTER_CHILD"):
        pytest.skip("already inside the child run")
    env["VQA_LSTM_CLUSTER_CHILD"] = "1"
    code = ("import sys, pytest; rc = pytest.main(['-q', '-x', 'tests/test_gpu_tc.py', '-k', 'persistent_lstm_matches']);"
            "from dl_vqa_b200 import lib; cs = lib.load().vqa_tc_lstm_cluster_size(); print('cluster_size', cs);"
            "sys.exit(int(rc) if int(rc) else (0 if cs == 4 else 7))")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]


@pytest.mark.parametrize("B,T,H,dirs", [(256, 23, 1024, 2), (5, 4, 128, 1), (130, 7, 256, 2), (300, 6, 256, 2)])
def test_persistent_lstm_backward_matches_the_per_step_kernels(B, T, H, dirs):
    """vqa_tc_lstm_bwd (one cooperative launch) == T x vqa_lstm_step_bwd_pointwise + (T-1) x split-K vqa_tc_gemm."""
    from dl_vqa_b200 import lib
    torch.manual_seed(B + T)
    dev = "cuda"
    gates = torch.rand(dirs, T, B, 4 * H, device=dev).bfloat16()                 # activated gates in (0, 1)
    gates[:, :, :, 2 * H:3 * H] = (torch.rand(dirs, T, B, H, device=dev) * 2 - 1).bfloat16()   # g gate in (-1, 1)
    cs = torch.randn(dirs, T, B, H, device=dev) * 0.5
    dqf = (torch.randn(B, dirs * H, device=dev) * 0.1).bfloat16()
    whh = (torch.randn(dirs, 4 * H, H, device=dev) / H ** 0.5).bfloat16()
    q_len = torch.randint(1, T + 1, (B,), device=dev, dtype=torch.int64)
    q_len[0] = T
    st = lib.stream()

    def per_step():
        dh = torch.zeros(dirs, B, H, device=dev)
        dc = torch.empty(dirs, B, H, device=dev)
        dg = torch.empty(dirs, T, B, 4 * H, dtype=torch.bfloat16, device=dev)
        for s in range(T - 1, -1, -1):
            lib.call("vqa_lstm_step_bwd_pointwise", lib.ptr(gates), lib.ptr(cs), lib.ptr(dh), lib.ptr(dc),
                     lib.ptr(dqf) if s == T - 1 else None, lib.ptr(dg), lib.ptr(q_len), lib.BF16, s, T, B, H, dirs, st)
            if s > 0:
                lib.call("vqa_tc_gemm", dg.data_ptr() + s * B * 4 * H * 2, 4 * H, T * B * 4 * H, lib.ptr(whh), H, 4 * H * H,
                         lib.ptr(dh), lib.F32, H, B * H, None, None, 0, B, H, 4 * H, dirs,
                         lib.GEMM_SPLITK | lib.GEMM_B_MN, 0.0, 0, 0, st)
        torch.cuda.synchronize()
        return dg, dh, dc

    def persistent():
        dh = torch.zeros(dirs, B, H, device=dev)
        dc = torch.empty(dirs, B, H, device=dev)
        dg = torch.empty(dirs, T, B, 4 * H, dtype=torch.bfloat16, device=dev)
        sync = torch.zeros(256, dtype=torch.int32, device=dev)
        lib.call("vqa_tc_lstm_bwd", lib.ptr(gates), lib.ptr(cs), lib.ptr(dh), lib.ptr(dc), lib.ptr(dqf), lib.ptr(dg),
                 lib.ptr(whh), lib.ptr(q_len), lib.ptr(sync), T, B, H, dirs, st)
        torch.cuda.synchronize()
        return dg, dh, dc

    dg_a, dh_a, dc_a = per_step()
    dg_b, dh_b, dc_b = persistent()

    def err(a, b):
        return float((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-20))
    assert err(dg_b, dg_a) < 1e-2          # bf16 gate gradients; fp32 partial sums arrive in a different order
    assert err(dc_b, dc_a) < 1e-4
    assert err(dh_b, dh_a) < 1e-4


# ------------------------------------------------------------------------------------------------------------------
# persistent LSTM kernels against the ORACLE (not against another kernel of this library)
# ------------------------------------------------------------------------------------------------------------------
def _step_index(pre_tok, q_len, reverse):
    """token-indexed [B,T,X] -> step-indexed [T,B,X]: step s of the reverse direction holds token len-1-s"""
    B, T, _ = pre_tok.shape
    s = torch.arange(T, device=pre_tok.device)[None, :]
    t = (q_len[:, None] - 1 - s).clamp_min(0) if reverse else s.expand(B, T)
    out = pre_tok.gather(1, t[:, :, None].expand_as(pre_tok))
    out = out * (s < q_len[:, None])[:, :, None]
    return out.transpose(0, 1).contiguous()


def _token_index(step_major, q_len, reverse):
    """inverse of _step_index for gradients: [T,B,X] step-indexed -> [B,T,X] token-indexed (zero where t >= len)"""
    x = step_major.transpose(0, 1)
    B, T, _ = x.shape
    t = torch.arange(T, device=x.device)[None, :]
    s = (q_len[:, None] - 1 - t).clamp_min(0) if reverse else t.expand(B, T)
    out = x.gather(1, s[:, :, None].expand_as(x))
    return out * (t < q_len[:, None])[:, :, None]


@pytest.mark.parametrize("ordered", [False, True])
@pytest.mark.parametrize("B,T,H,dirs", [(256, 23, 1024, 2), (1024, 12, 1024, 2), (130, 7, 256, 2), (5, 4, 128, 1), (600, 9, 256, 2)])
def test_persistent_lstm_forward_and_backward_against_the_oracle(B, T, H, dirs, ordered):
    """vqa_tc_lstm_fwd / vqa_tc_lstm_bwd (cooperative tcgen05 kernels) against oracle.lstm_final_cell + autograd
    (reference models/model.py:159-166) on the same bf16-rounded input projections and recurrent weights; the oracle
    rounds what the kernel stores in bf16 (input projection, h) with a straight-through gradient.
    ordered: rows handed over in descending length order (vqa_length_order) through the *_ordered entries, so that whole
    128-row tiles drop out of the late steps; qf / dc_init stay in sample order."""
    from dl_vqa_b200 import lib
    from oracle import vqa_oracle as O
    torch.manual_seed(B + T + H)
    dev = "cuda"
    pre_tok = [(torch.randn(B, T, 4 * H) * 0.8).bfloat16().float() for _ in range(dirs)]
    w_hh = [(torch.randn(4 * H, H) / H ** 0.5).bfloat16().float() for _ in range(dirs)]
    q_len = torch.randint(1, T + 1, (B,))
    q_len[0] = T
    q_len[-1] = 1
    dqf = (torch.randn(B, dirs * H) * 0.1).bfloat16().float()

    # ---- oracle
    sd = {"text.lstm.weight_hh_l0": w_hh[0]}
    if dirs == 2:
        sd["text.lstm.weight_hh_l0_reverse"] = w_hh[1]
    leaves = [p.clone().requires_grad_(True) for p in pre_tok]
    c_n = O.lstm_final_cell(sd, torch.zeros(B, T, 1), q_len, H, dirs == 2, rnd=O.bf16_round_ste, pre=leaves)
    (c_n * dqf).sum().backward()

    # ---- kernels
    st = lib.stream()
    ql = q_len.to(dev)
    gx = torch.stack([_step_index(pre_tok[d].to(dev), ql, d == 1) for d in range(dirs)]).bfloat16().contiguous()
    order = len_rows = None
    if ordered:
        order = torch.empty(B, dtype=torch.int32, device=dev)
        len_rows = torch.empty(B, dtype=torch.int64, device=dev)
        lib.call("vqa_length_order", lib.ptr(ql), lib.ptr(order), lib.ptr(len_rows), B, T, st)
        gx = gx[:, :, order.long()].contiguous()              # row j of the step-indexed buffers = sample order[j]
    rows_len = len_rows if ordered else ql
    cs = torch.empty(dirs, T, B, H, device=dev)
    hs = torch.zeros(dirs, T + 1, B, H, device=dev, dtype=torch.bfloat16)
    qf = torch.empty(B, dirs * H, device=dev, dtype=torch.bfloat16)
    wp = torch.empty(dirs, 4 * H, H, device=dev, dtype=torch.bfloat16)
    whh = torch.stack(w_hh).to(dev)
    for d in range(dirs):
        lib.call("vqa_pack_lstm_whh", lib.ptr(whh[d]), lib.ptr(wp[d]), H, st)
    sync = torch.zeros(dirs, dtype=torch.int32, device=dev)
    lib.call("vqa_tc_lstm_fwd_ordered", lib.ptr(gx), lib.ptr(cs), lib.ptr(hs), lib.ptr(qf), lib.ptr(wp), lib.ptr(rows_len),
             lib.ptr(order), lib.ptr(sync), T, B, H, dirs, st)
    dh = torch.zeros(dirs, B, H, device=dev)
    dc = torch.empty(dirs, B, H, device=dev)
    dg = torch.empty(dirs, T, B, 4 * H, dtype=torch.bfloat16, device=dev)
    sync_b = torch.zeros(256, dtype=torch.int32, device=dev)
    whh_b = whh.bfloat16().contiguous()
    lib.call("vqa_tc_lstm_bwd_ordered", lib.ptr(gx), lib.ptr(cs), lib.ptr(dh), lib.ptr(dc), lib.ptr(dqf.to(dev).bfloat16()), lib.ptr(dg),
             lib.ptr(whh_b), lib.ptr(rows_len), lib.ptr(order), lib.ptr(sync_b), T, B, H, dirs, st)
    torch.cuda.synchronize()
    if ordered:
        o = order.long()
        assert torch.equal(len_rows, ql[o]) and bool((len_rows[:-1] >= len_rows[1:]).all())
        assert torch.equal(torch.sort(o).values, torch.arange(B, device=dev))
        same = len_rows[:-1] == len_rows[1:]
        assert bool((o[1:][same] > o[:-1][same]).all())            # ties keep sample order
        inv = torch.empty_like(o)
        inv[o] = torch.arange(B, device=dev)
        dg = dg[:, :, inv].contiguous()                            # back to sample order for the comparison
        assert bool(torch.isfinite(hs.float()).all())              # frozen rows are copied forward, never left unwritten

    def err(a, b):
        a, b = a.float().cpu(), b.float().cpu()
        return float((a - b).abs().max() / (b.abs().max() + 1e-20))
    assert err(qf, c_n.detach()) < 2e-2, err(qf, c_n.detach())
    for d in range(dirs):
        got = _token_index(dg[d].float(), ql, d == 1)
        e = err(got, leaves[d].grad)
        assert e < 2e-2, (d, e)
        # inactive steps: exactly zero on both sides
        inactive = (torch.arange(T)[None, :] >= q_len[:, None])
        assert float(leaves[d].grad[inactive].abs().max() if inactive.any() else 0.0) == 0.0


@pytest.mark.parametrize("B,IH,IW", [(3, 224, 224), (2, 64, 64), (2, 39, 45), (1, 20, 36), (2, 38, 38)])
def test_tc_conv0_reads_float16_images_bit_identically(B, IH, IW):
    """vqa_tc_conv0_*_x with x_dtype = VQA_F16 (the dtype the reference stores its images in) against the same entries fed
    with the exact float32 widening: forward outputs, masks and the weight / bias gradients must be IDENTICAL bit for
    bit -- staged (TMA, row pitch a multiple of 16 bytes) and direct-load paths, odd widths included."""
    from dl_vqa_b200 import lib
    torch.manual_seed(IH + IW)
    Cin, Cout = 3, 64
    x16 = torch.randn(B, Cin, IH, IW, device="cuda").half()
    x32 = x16.float()
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / 5
    bias = torch.randn(Cout, device="cuda") * 0.1
    PH, PW = (IH - 2) // 2, (IW - 2) // 2
    dpool = torch.randn(B, PH, PW, Cout, device="cuda").bfloat16()
    res = []
    for x, code in ((x32, lib.F32), (x16, lib.F16)):
        out = torch.empty(B, PH, PW, Cout, dtype=torch.bfloat16, device="cuda")
        mask = torch.empty(B, PH, PW, Cout, dtype=torch.uint8, device="cuda")
        lib.call("vqa_tc_conv0_relu_pool_fwd_x", lib.ptr(x), code, lib.ptr(w), lib.ptr(bias), lib.ptr(out), lib.ptr(mask),
                 B, IH, IW, Cin, Cout, lib.stream())
        dw = torch.empty(Cout, Cin, 3, 3, device="cuda")
        db = torch.empty(Cout, device="cuda")
        lib.call("vqa_tc_conv0_bwd_weight_bias_x", lib.ptr(x), code, lib.ptr(dpool), lib.ptr(mask), lib.ptr(dw), lib.ptr(db),
                 B, IH, IW, Cin, Cout, lib.stream())
        torch.cuda.synchronize()
        res.append((out, mask, dw, db))
    a, b = res
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # fp32 atomics across CTAs: the order of the partial sums is not fixed, so compare to rounding of the sum
    assert float((a[2] - b[2]).abs().max()) <= 1e-5 * float(a[2].abs().max())
    assert float((a[3] - b[3]).abs().max()) <= 1e-5 * float(a[3].abs().max())


def _active_kblocks_reference(lens, B, T, s_begin):
    """The oracle's restatement (pinned on the CPU to pack_padded_sequence's batch_sizes, tests/test_oracle.py):
    64-row blocks of steps s_begin..T-1 with any len > step."""
    from oracle import vqa_oracle as O
    assert lens.numel() == B
    return O.lstm_live_row_blocks(lens, T, s_begin)


@pytest.mark.parametrize("ordered", [False, True])
@pytest.mark.parametrize("B,T", [(256, 23), (64, 5), (1024, 23), (128, 1), (192, 9)])
def test_lstm_active_kblocks_lists(B, T, ordered):
    """vqa_lstm_active_kblocks against the definition (a block is live when any of its rows has len > step), for rows in
    sample order and in descending length order, incl. lengths outside [0, T] (clamped) and T = 1 (empty second list)."""
    from dl_vqa_b200 import lib
    torch.manual_seed(B * 31 + T)
    lens = torch.randint(1, T + 1, (B,))
    lens[0], lens[-1] = T + 5, -3
    if B > 64:
        lens[64:128] = torch.randint(1, max(2, T // 3 + 1), (64,))       # one short group in the middle
    if ordered:
        lens = torch.sort(lens, descending=True).values
    d = lens.cuda()
    l0 = torch.full((1 + T * (B // 64),), -7, dtype=torch.int32, device="cuda")
    l1 = torch.full((1 + max(T - 1, 0) * (B // 64),), -7, dtype=torch.int32, device="cuda")
    lib.call("vqa_lstm_active_kblocks", lib.ptr(d), lib.ptr(l0), lib.ptr(l1), B, T, lib.stream())
    torch.cuda.synchronize()
    for lst, s_begin in ((l0, 0), (l1, 1)):
        want = _active_kblocks_reference(lens, B, T, s_begin)
        n = int(lst[0])
        assert n == len(want), (s_begin, n, len(want))
        assert lst[1:1 + n].tolist() == want
        assert bool((lst[1 + n:] == -7).all())                            # nothing written past the list
    # one list only
    l0b = torch.full_like(l0, -7)
    lib.call("vqa_lstm_active_kblocks", lib.ptr(d), lib.ptr(l0b), None, B, T, lib.stream())
    torch.cuda.synchronize()
    assert torch.equal(l0b, l0)
    with pytest.raises(lib.VqaLibraryError):
        lib.call("vqa_lstm_active_kblocks", lib.ptr(d), lib.ptr(l0), lib.ptr(l1), 100, T, lib.stream())


@pytest.mark.parametrize("R,N,K,splitk", [(5888, 4096, 304, True), (5632, 1024, 1024, True), (23552, 4096, 1024, True),
                                          (640, 256, 136, False), (4096, 128, 64, True), (8192, 200, 72, True)])
def test_tc_gemm_kblocks_reads_listed_blocks_only(R, N, K, splitk):
    """vqa_tc_gemm_kblocks: dW[N,K] = sum over the LISTED 64-row blocks of dY^T X.  (1) every unlisted block is zero in dY:
    equals the dense product, bit for bit where the dense kernel does not split K either; (2) unlisted blocks hold data
    (even NaN): they are not read; (3) empty list: dW = 0.  Covers unsplit launches (many tiles) and split-K ones (few
    tiles: the list, not K, is divided among the splits), lists longer than one 32-entry fetch, and a one-entry list."""
    from dl_vqa_b200 import lib
    torch.manual_seed(R + N)
    nb = R // 64
    dY = torch.randn(R, N, device="cuda").bfloat16()
    X = torch.randn(R, K, device="cuda").bfloat16()
    flags = lib.GEMM_OPERANDS_MN | (lib.GEMM_SPLITK if splitk else 0)

    def run(lst, dy=dY):
        l = torch.tensor([len(lst)] + lst + [10 ** 6] * 3, dtype=torch.int32, device="cuda")   # junk past the end is ignored
        dW = torch.zeros(N, K, device="cuda")
        lib.call("vqa_tc_gemm_kblocks", lib.ptr(dy), N, lib.ptr(X), K, lib.ptr(dW), K, N, K, R, flags, lib.ptr(l), lib.stream())
        torch.cuda.synchronize()
        return dW

    def want_of(lst, dy=dY):
        rows = torch.tensor([b * 64 + i for b in lst for i in range(64)], dtype=torch.long, device="cuda")
        return dy[rows].float().t() @ X[rows].float() if len(lst) else torch.zeros(N, K, device="cuda")

    g = torch.Generator().manual_seed(R)
    keep = sorted(torch.randperm(nb, generator=g)[: max(1, (nb * 3) // 5)].tolist())
    for lst in (keep, keep[:1], list(range(nb))):
        want = want_of(lst)
        got = run(lst)
        err = float((got - want).abs().max() / want.abs().max())
        assert err < 1e-4, (len(lst), err)
    # (2) unlisted blocks poisoned
    poisoned = dY.clone()
    mask = torch.ones(nb, dtype=torch.bool)
    mask[keep] = False
    poisoned.view(nb, 64, N)[mask.cuda()] = float("nan")
    got = run(keep, poisoned)
    assert bool(torch.isfinite(got).all())
    assert float((got - want_of(keep)).abs().max() / want_of(keep).abs().max()) < 1e-4
    # (1) zeroed blocks: same numbers as the dense entry
    zeroed = dY.clone()
    zeroed.view(nb, 64, N)[mask.cuda()] = 0
    dense = torch.zeros(N, K, device="cuda")
    lib.call("vqa_tc_gemm", lib.ptr(zeroed), N, 0, lib.ptr(X), K, 0, lib.ptr(dense), lib.F32, K, 0, None, None, 0,
             N, K, R, 1, flags, 0.0, 0, 0, lib.stream())
    torch.cuda.synchronize()
    got = run(keep, zeroed)
    assert float((got - dense).abs().max() / dense.abs().max()) < 2e-6
    # (3) empty list
    assert float(run([]).abs().max()) == 0.0
    # argument checks: the list form exists for the reduction-major operands only
    with pytest.raises(lib.VqaLibraryError):
        lib.call("vqa_tc_gemm_kblocks", lib.ptr(dY), N, lib.ptr(X), K, lib.ptr(dense), K, N, K, R, lib.GEMM_SPLITK,
                 lib.ptr(torch.zeros(4, dtype=torch.int32, device="cuda")), lib.stream())
    with pytest.raises(lib.VqaLibraryError):
        lib.call("vqa_tc_gemm_kblocks", lib.ptr(dY), N, lib.ptr(X), K, lib.ptr(dense), K, N, K, R, flags, None, lib.stream())
