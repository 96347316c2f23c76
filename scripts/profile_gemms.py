"""Which GEMMs of a MambaUnet training step are far from their flop / byte bound?  (torch profiler with shapes)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
from selscan_b200 import workloads as wl
from selscan_b200.vssm import DiceLoss, MambaUnet

dev = torch.device("cuda")
if os.environ.get("BLAS"):
    torch.backends.cuda.preferred_blas_library(os.environ["BLAS"])
    print("blas:", torch.backends.cuda.preferred_blas_library())
torch.manual_seed(0)
model = MambaUnet(num_classes=4).to(dev).train()
opt = wl.make_sgd(model)
dice = DiceLoss(4)
x = torch.rand(24, 1, 224, 224, device=dev)
y = torch.randint(0, 4, (24, 224, 224), device=dev)
for _ in range(3):
    wl.supervised_step(model, opt, dice, x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    wl.supervised_step(model, opt, dice, x, y)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.key in ("aten::mm", "aten::bmm", "aten::baddbmm_", "aten::addmm", "aten::baddbmm"):
        shp = [s for s in e.input_shapes if s]
        rows.append((e.device_time_total / 1e3, e.count, e.key, shp))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total GEMM ms per step: %.2f" % tot)
for ms, n, k, shp in rows[:45]:
    # flops / bytes estimate from the two matrix operands
    mats = [s for s in shp if len(s) >= 2][-2:]
    a, b = mats
    if len(a) == 2:
        M, K = a; N = b[1]; batch = 1
    else:
        batch, M, K = a[-3], a[-2], a[-1]; N = b[-1]
    fl = 2.0 * batch * M * N * K
    by = 4.0 * batch * (M * K + K * N + M * N)
    each = ms / n
    print("%7.3f ms  x%-3d %-14s %-44s  %6.1f us each  %5.1f TF/s  %6.0f GB/s" % (ms, n, k, str(shp)[:44], each * 1e3, fl / each / 1e9, by / each / 1e6))
