"""Where does a MambaUnet training step spend its GPU time?  (torch profiler, top kernels)  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
from selscan_b200 import workloads as wl
from selscan_b200.vssm import DiceLoss, MambaUnet

dev = torch.device("cuda")
torch.manual_seed(0)
model = MambaUnet(num_classes=4).to(dev).train()
if os.environ.get("TC"):
    from selscan_b200 import ss2d, tcgemm
    tcgemm.patch_linears(model)
    ss2d.TC_PROJ = True
opt = wl.make_sgd(model)
dice = DiceLoss(4)
x = torch.rand(24, 1, 224, 224, device=dev)
y = torch.randint(0, 4, (24, 224, 224), device=dev)
for _ in range(3):
    wl.supervised_step(model, opt, dice, x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        wl.supervised_step(model, opt, dice, x, y)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
