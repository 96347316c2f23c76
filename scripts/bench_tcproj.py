"""The six projection GEMMs of an SS2D block (x_proj / dt_proj forward, their weight and input gradients): tcgemm.bgemm vs cuBLAS."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from selscan_b200 import tcgemm
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
B, K, N = 24, 4, 16
for name, D, L in (("S1", 192, 3136), ("S2", 384, 784), ("S3", 768, 196)):
    R = D // 2 // 16 if D // 2 % 16 == 0 else (D // 2 + 15) // 16
    R = max(R, 1); R = {192: 6, 384: 12, 768: 24, 1536: 48}[D]
    C = R + 2 * N; R4 = (R + 3) // 4 * 4
    xs = torch.randn(B * K, D, L, device=dev); Wx = torch.randn(K, C, D, device=dev); Wd = torch.zeros(K, D, R4, device=dev); Wd[:, :, :R].normal_()
    x_dbl = torch.empty(B * K, C, L, device=dev); dts = torch.empty(B * K, D, L, device=dev)
    dd = torch.randn(B * K, D, L, device=dev); du = torch.randn(B * K, D, L, device=dev); dxd = torch.randn(B * K, C, L, device=dev)
    xs4, dd4, dxd4 = xs.view(B, K, D, L), dd.view(B, K, D, L), dxd.view(B, K, C, L)
    cases = [
        ("x_proj fwd", lambda: tcgemm.bgemm(Wx, xs, x_dbl, b_mn=True), lambda: torch.matmul(Wx.unsqueeze(0), xs4)),
        ("dt_proj fwd", lambda: tcgemm.bgemm(Wd, x_dbl[:, :R4], dts, b_mn=True), lambda: torch.matmul(Wd[:, :, :R].unsqueeze(0), x_dbl.view(B, K, C, L)[:, :, :R])),
        ("d_dt_w", lambda: tcgemm.bgemm(dd, x_dbl[:, :R4], torch.empty(K, D, R4, device=dev)), lambda: torch.matmul(dd4, x_dbl.view(B, K, C, L)[:, :, :R].transpose(-1, -2)).sum(0)),
        ("d_dtr", lambda: tcgemm.bgemm(Wd, dd, torch.empty(B * K, R4, L, device=dev), a_mn=True, b_mn=True), lambda: torch.matmul(Wd[:, :, :R].transpose(-1, -2).unsqueeze(0), dd4)),
        ("d_xproj_w", lambda: tcgemm.bgemm(dxd, xs, torch.empty(K, C, D, device=dev)), lambda: torch.matmul(dxd4, xs4.transpose(-1, -2)).sum(0)),
        ("d_xs +=", lambda: tcgemm.bgemm(Wx, dxd, du, a_mn=True, b_mn=True, accumulate=True), lambda: du.baddbmm_(Wx.transpose(-1, -2).unsqueeze(0).expand(B, K, D, C).reshape(B * K, D, C), dxd)),
    ]
    for cname, ours, ref in cases:
        print(name, "%-12s ours %.4f ms   cuBLAS %.4f ms" % (cname, timeit(ours), timeit(ref)), flush=True)
