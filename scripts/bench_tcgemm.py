"""3xTF32 tensor-core GEMM vs cuBLAS fp32 on the Linear shapes of MambaUnet (batch 24): forward, dgrad, wgrad."""
import json, os, sys
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

rows = []
# (rows M = B*H*W, in features, out features) of the Linear layers
for M, Kin, Nout in [(75264, 96, 384), (75264, 192, 96), (18816, 192, 768), (18816, 384, 192), (4704, 384, 1536), (4704, 768, 384),
                     (1176, 768, 3072), (1176, 1536, 768), (75264, 96, 1536), (18816, 384, 192)]:
    x = torch.randn(M, Kin, device=dev); w = torch.randn(Nout, Kin, device=dev); dy = torch.randn(M, Nout, device=dev)
    for name, ours, ref in (
            ("fwd", lambda: tcgemm.gemm(x, w), lambda: x @ w.T),
            ("dgrad", lambda: tcgemm.gemm(dy, w, b_mn=True), lambda: dy @ w),
            ("wgrad", lambda: tcgemm.gemm(dy, x, a_mn=True, b_mn=True), lambda: dy.T @ x)):
        o, r = ours(), ref()
        err = float((o - r).abs().max() / r.abs().max())
        t_o, t_r = timeit(ours), timeit(ref)
        fl = 2.0 * M * Kin * Nout
        rows.append(dict(M=M, K_in=Kin, N_out=Nout, op=name, ours_ms=round(t_o, 4), cublas_ms=round(t_r, 4), speedup=round(t_r / t_o, 2),
                         ours_tflops=round(fl / t_o / 1e9, 1), rel_diff=float("%.2e" % err)))
        print(rows[-1], flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "tcgemm_bench.json"), "w"), indent=1)
