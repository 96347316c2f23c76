"""The 3xTF32 tensor-core GEMM (selscan_b200_gemm_3xtf32) against an fp64 reference: every operand-major combination, ragged
sizes, split-K, accumulate; and the Linear autograd form against F.linear.  Bar: the error vs fp64 must stay within 4x the error
of cuBLAS's own fp32 GEMM on the same inputs (it is ~2x in practice), i.e. fp32-level accuracy, not TF32-level (2^-11)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _err(x, ref64):
    return float((x.double() - ref64).abs().max())


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 96, 200), (1000, 384, 96), (77, 48, 36), (384, 96, 7000), (130, 260, 64)])
def test_gemm_matches_fp64(M, N, K, a_mn, b_mn):
    from selscan_b200 import tcgemm

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(M + N + K)
    # leading dimensions must be multiples of 4 floats: pad the stored matrices when a size is not
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(N, K, device="cuda")
    ref64 = A.double() @ B.double().T

    def stored(t, mn):          # (rows, cols) -> storage the kernel accepts, as a view with aligned row stride
        t = t.T if mn else t
        r, c = t.shape
        buf = torch.zeros(r, (c + 3) // 4 * 4, device="cuda")
        buf[:, :c] = t
        return buf[:, :c]

    out = tcgemm.gemm(stored(A, a_mn), stored(B, b_mn), a_mn=a_mn, b_mn=b_mn)
    e_ours, e_blas = _err(out, ref64), _err(A @ B.T, ref64)
    assert e_ours <= 4 * e_blas + 1e-6, (e_ours, e_blas)
    # a single TF32 product would be ~1000x worse
    acc = torch.ones(M, N, device="cuda")
    tcgemm.gemm(stored(A, a_mn), stored(B, b_mn), a_mn=a_mn, b_mn=b_mn, out=acc, accumulate=True)
    assert _err(acc - 1.0, ref64) <= 4 * e_blas + 1e-5


def test_linear_autograd_matches_torch():
    from selscan_b200 import tcgemm

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    x = torch.randn(3, 14, 14, 96, device="cuda", requires_grad=True)
    lin = torch.nn.Linear(96, 384, bias=True).cuda()
    g = torch.randn(3, 14, 14, 384, device="cuda")
    ref = F.linear(x, lin.weight, lin.bias)
    gx, gw, gb = torch.autograd.grad(ref, (x, lin.weight, lin.bias), g)
    assert tcgemm.patch_linears(lin) == 1
    out = lin(x)
    hx, hw, hb = torch.autograd.grad(out, (x, lin.weight, lin.bias), g)
    for a, r, name in ((out, ref, "y"), (hx, gx, "dx"), (hw, gw, "dW"), (hb, gb, "db")):
        torch.testing.assert_close(a, r, rtol=1e-5, atol=2e-5 * float(r.abs().max()), msg=lambda m: f"{name}: {m}")


def test_batched_gemm_with_shared_operands_and_summed_output():
    """bgemm: per-direction weights shared cyclically across images (x_proj), MN-major operands, accumulate, and an output that
    sums over the images that map to it (the weight-gradient form)."""
    from selscan_b200 import tcgemm

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    Bt, K, C, D, L = 3, 4, 40, 64, 200
    W = torch.randn(K, C, D, device="cuda")
    xs = torch.randn(Bt * K, D, L, device="cuda")
    ref = torch.matmul(W.repeat(Bt, 1, 1).double(), xs.double())                       # entry i uses W[i % K]
    out = torch.empty(Bt * K, C, L, device="cuda")
    tcgemm.bgemm(W, xs, out, b_mn=True)
    e_blas = float((torch.matmul(W.repeat(Bt, 1, 1), xs).double() - ref).abs().max())
    assert float((out.double() - ref).abs().max()) <= 4 * e_blas + 1e-6
    # d(xs) += W^T g   (A MN-major, accumulate)
    g = torch.randn(Bt * K, C, L, device="cuda")
    base = torch.randn(Bt * K, D, L, device="cuda")
    ref2 = base.double() + torch.matmul(W.repeat(Bt, 1, 1).double().transpose(1, 2), g.double())
    acc = base.clone()
    tcgemm.bgemm(W, g, acc, a_mn=True, b_mn=True, accumulate=True)
    assert float((acc.double() - ref2).abs().max()) <= 1e-4
    # d W[k] = sum_b g[b, k] xs[b, k]^T   (output shared by the images)
    ref3 = torch.matmul(g.double(), xs.double().transpose(1, 2)).view(Bt, K, C, D).sum(0)
    dW = torch.empty(K, C, D, device="cuda")
    tcgemm.bgemm(g, xs, dW)
    assert float((dW.double() - ref3).abs().max()) <= 2e-5 * float(ref3.abs().max())
