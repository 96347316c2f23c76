import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from selscan_b200 import tcgemm
M, Kin, Nout = 75264, 96, 384
x = torch.randn(M, Kin, device="cuda"); w = torch.randn(Nout, Kin, device="cuda"); dy = torch.randn(M, Nout, device="cuda")
for _ in range(2):
    tcgemm.gemm(x, w); tcgemm.gemm(dy, w, b_mn=True); tcgemm.gemm(dy, x, a_mn=True, b_mn=True)
torch.cuda.synchronize()
