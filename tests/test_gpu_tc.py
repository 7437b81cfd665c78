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
                                   (1000, 64, 72), (4096, 1024, 304), (130, 40, 8)])
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
