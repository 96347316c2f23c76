import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from selscan_b200 import tcgemm
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def check(M, N, K, a_mn, b_mn):
    torch.manual_seed(M + N + K)
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    ref64 = (A.double() @ B.double().T)
    a_in = A.T.contiguous() if a_mn else A
    b_in = B.T.contiguous() if b_mn else B
    out = tcgemm.gemm(a_in, b_in, a_mn=a_mn, b_mn=b_mn)
    torch.cuda.synchronize()
    err = (out.double() - ref64).abs().max().item()
    err_blas = ((A @ B.T).double() - ref64).abs().max().item()
    scale = ref64.abs().max().item()
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: max err {err:.3e} (cuBLAS fp32 {err_blas:.3e}), scale {scale:.2f}", flush=True)
for args in [(128, 128, 32, False, False), (128, 128, 96, False, False), (256, 96, 64, False, False), (1000, 384, 96, False, False),
             (128, 128, 32, False, True), (128, 128, 32, True, True), (300, 96, 200, False, True), (384, 96, 7000, True, True),
             (75264, 384, 96, False, False)]:
    check(*args)
