"""Where does ss2d.MIRROR lose its time?  Per-op CUDA time of one SS2D block fwd+bwd (stage-1 shape, batch 24), MIRROR on / off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from torch.profiler import profile, ProfilerActivity
from selscan_b200 import ss2d
from selscan_b200.vssm import SS2D

torch.manual_seed(0)
d_model, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (96, 56)
blk = SS2D(d_model).cuda()
x = torch.randn(24, H, H, d_model, device="cuda", requires_grad=True)
g = torch.randn(24, H, H, d_model, device="cuda")
for mir in (False, True):
    ss2d.MIRROR = mir
    for _ in range(3):
        blk(x).backward(g)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            blk(x).backward(g)
        torch.cuda.synchronize()
    print("==== MIRROR", mir)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
