"""One stage-shaped selective-scan forward + backward through the C ABI (for ncu captures).
    python scripts/run_scan_once.py [stage S1..S4] [batch] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from bench import STAGES
from selscan_b200 import ops

stage = sys.argv[1] if len(sys.argv) > 1 else "S1"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 24
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
d_inner, L = next((d, l) for n, d, l, _ in STAGES if n == stage)
dev = torch.device("cuda")
kd, N, K = 4 * d_inner, 16, 4
torch.manual_seed(0)
t = {"u": ops.empty_rows(b, kd, L, dev).normal_(), "delta": ops.empty_rows(b, kd, L, dev).normal_().mul_(0.5),
     "A": -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(kd, 1).contiguous(),
     "B": torch.randn(b, K, N, L, device=dev), "C": torch.randn(b, K, N, L, device=dev),
     "D": torch.ones(kd, device=dev), "bias": torch.full((kd,), -4.6, device=dev),
     "dout": ops.empty_rows(b, kd, L, dev).normal_(), "out": ops.empty_rows(b, kd, L, dev),
     "ck": torch.empty(max(ops.ckpt_elems(b, kd, L, N), 4), device=dev),
     "du": ops.empty_rows(b, kd, L, dev), "dd": ops.empty_rows(b, kd, L, dev)}
nbc = b * K * N * L
flat = torch.zeros(2 * nbc + kd * N + 2 * kd, device=dev)
dB, dC = flat[:nbc].view(b, K, N, L), flat[nbc:2 * nbc].view(b, K, N, L)
dA = flat[2 * nbc:2 * nbc + kd * N].view(kd, N)
dD, db = flat[2 * nbc + kd * N:2 * nbc + kd * N + kd], flat[2 * nbc + kd * N + kd:]
for _ in range(iters):
    ops.launch_fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], True, t["out"], None, None, t["ck"], None)
    flat.zero_()
    ops.launch_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], t["dout"], None, t["ck"], True,
                   t["du"], t["dd"], dA, dB, dC, dD, None, db)
torch.cuda.synchronize()
print("ok", stage, b, float(t["du"].abs().mean()), float(dB.abs().mean()))
