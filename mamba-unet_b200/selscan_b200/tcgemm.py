"""fp32 GEMMs on the tcgen05 tensor cores with the 3xTF32 split (libselscan_b200: selscan_b200_gemm_3xtf32), OPT-IN.

`Linear3xTF32Fn` is the autograd form of `F.linear` for the layers either side of the scan (SS2D.in_proj / out_proj,
/root/reference/code/networks/mamba_sys.py:299,336): forward y = x W^T, dgrad dx = dy W and wgrad dW = dy^T x all run on the
same kernel with K-major / MN-major operand descriptors, no transpose passes.  Results agree with cuBLAS fp32 to ~1e-6 relative
(tests/test_tcgemm_gpu.py); the default model path keeps cuBLAS fp32 so that the reference's arithmetic is untouched."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


def _ok(t):
    return t.data_ptr() % 16 == 0 and t.stride(-1) == 1 and t.stride(-2) % 4 == 0 and t.stride(-2) >= t.shape[-1]


def gemm(a, b, a_mn=False, b_mn=False, out=None, accumulate=False):
    """out (M, N) (+)= A (M, K) @ B (N, K)^T.  a: (M, K) [a_mn=False] or (K, M) [a_mn=True]; b: (N, K) or (K, N) likewise.
    2-D fp32 CUDA tensors with unit inner stride; rows 16-byte aligned."""
    lib = _lib.load()
    M, K = (a.shape[1], a.shape[0]) if a_mn else a.shape
    N, K2 = (b.shape[1], b.shape[0]) if b_mn else b.shape
    if K != K2:
        raise RuntimeError(f"gemm: reduction sizes differ ({K} vs {K2})")
    if not (_ok(a) and _ok(b)):
        a, b = a.contiguous(), b.contiguous()
    if out is None:          # rows of C leave through TMA stores: 16-byte aligned rows
        n4 = (N + 3) // 4 * 4
        out = torch.empty((M, n4), device=a.device, dtype=torch.float32)
        out = out if n4 == N else out[:, :N]
    elif not _ok(out):
        raise RuntimeError("gemm: `out` must have unit inner stride and 16-byte aligned rows")
    with torch.cuda.device(a.device):
        _lib.check(lib.selscan_b200_gemm_3xtf32(a.data_ptr(), a.stride(0), int(a_mn), b.data_ptr(), b.stride(0), int(b_mn),
                                                out.data_ptr(), out.stride(0), M, N, K, 1, 0, 0, 0, int(accumulate), 0, 0, 0,
                                                torch.cuda.current_stream(a.device).cuda_stream), "selscan_b200_gemm_3xtf32")
    return out


def bgemm_ok(*ts):
    """Can these (batch, rows, cols) operands go through the TMA descriptors?"""
    return all(t.is_cuda and t.dtype == torch.float32 and t.dim() == 3 and _ok(t) and (t.shape[0] == 1 or t.stride(0) % 4 == 0)
               for t in ts)


def bgemm(a, b, out, a_mn=False, b_mn=False, accumulate=False, batch=None):
    """out[i % len(out)] (+)= A[i % len(a)] @ B[i % len(b)]^T for i < batch: 3-D (entries, rows, cols) fp32 tensors.
    An operand with fewer entries than `batch` is shared cyclically (weights per direction); an `out` with fewer entries
    receives the SUM over the batch entries that map to it (weight gradients)."""
    lib = _lib.load()
    M, K = (a.shape[2], a.shape[1]) if a_mn else a.shape[1:]
    N, K2 = (b.shape[2], b.shape[1]) if b_mn else b.shape[1:]
    if K != K2 or tuple(out.shape[1:]) != (M, N):
        raise RuntimeError(f"bgemm: shapes do not match: A {tuple(a.shape)} B {tuple(b.shape)} out {tuple(out.shape)}")
    batch = batch or max(a.shape[0], b.shape[0], out.shape[0])
    mods = [0 if t.shape[0] == batch else t.shape[0] for t in (a, b, out)]
    with torch.cuda.device(a.device):
        _lib.check(lib.selscan_b200_gemm_3xtf32(a.data_ptr(), a.stride(1), int(a_mn), b.data_ptr(), b.stride(1), int(b_mn),
                                                out.data_ptr(), out.stride(1), M, N, K, batch, a.stride(0), b.stride(0),
                                                out.stride(0), int(accumulate), mods[0], mods[1], mods[2],
                                                torch.cuda.current_stream(a.device).cuda_stream), "selscan_b200_gemm_3xtf32")
    return out


class Linear3xTF32Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1])
        y = gemm(x2, weight)                                   # (M, K) x (N, K)^T
        if bias is not None:
            y += bias
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not _ok(dy2):
            dy2 = dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy2, weight, b_mn=True).view(ctx.x_shape)      # dy (M, N) x W viewed as (K_in, N)^T: W is stored (N, K_in)
        if ctx.needs_input_grad[1]:
            dw = gemm(dy2, x2, a_mn=True, b_mn=True)                 # dy^T (N, M) x x^T: both stored with the reduction as rows
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = dy2.sum(0)
        return dx, dw, db


def linear(x, weight, bias=None):
    if x.is_cuda and x.dtype == torch.float32 and weight.dtype == torch.float32 and x.shape[-1] % 4 == 0 and weight.shape[0] % 4 == 0:
        return Linear3xTF32Fn.apply(x, weight, bias)
    return F.linear(x, weight, bias)


class Linear(nn.Linear):
    def forward(self, x):
        return linear(x, self.weight, self.bias)


def patch_linears(model):
    """Opt in: route every nn.Linear of `model` through the tensor-core GEMM (parameters / state dict untouched)."""
    n = 0
    for m in model.modules():
        if type(m) is nn.Linear:
            m.__class__ = Linear
            n += 1
    return n
