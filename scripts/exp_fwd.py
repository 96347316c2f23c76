"""Experiment: forward kernel time with / without checkpoint writes, per stage (B=24).  Not part of the product."""
import os, sys, statistics
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from bench import STAGES, bytes_fwd
from selscan_b200 import ops

def timeit(fn, warm=3, iters=10):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)

dev = torch.device("cuda"); b = 24
for name, d_inner, L, calls in STAGES:
    kd, N, K = 4 * d_inner, 16, 4
    u = torch.randn(b, kd, L, device=dev); dt = 0.5 * torch.randn(b, kd, L, device=dev)
    A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(kd, 1).contiguous()
    Bm = torch.randn(b, K, N, L, device=dev); Cm = torch.randn(b, K, N, L, device=dev)
    D = torch.ones(kd, device=dev); bias = torch.full((kd,), -4.6, device=dev)
    out = torch.empty_like(u); ck = torch.empty(max(ops.ckpt_elems(b, kd, L, N), 4), device=dev)
    t_ck = timeit(lambda: ops.launch_fwd(u, dt, A, Bm, Cm, D, None, bias, True, out, None, None, ck))
    t_no = timeit(lambda: ops.launch_fwd(u, dt, A, Bm, Cm, D, None, bias, True, out, None, None, None))
    t_cp = timeit(lambda: out.copy_(u))
    print(f"{name}: fwd+ckpt {t_ck:.3f} ms  fwd(no ckpt) {t_no:.3f} ms ({bytes_fwd(b,kd,L)/t_no/1e6:.0f} GB/s)  copy u->out {t_cp:.3f} ms ({2*u.numel()*4/t_cp/1e6:.0f} GB/s)")
