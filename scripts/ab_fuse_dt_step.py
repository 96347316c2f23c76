"""Same-process A/B of the supervised MambaUnet training step (batch 24, fp32) with ss2d.FUSE_DT on / off, plus peak memory."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from selscan_b200 import ss2d, workloads
from selscan_b200.vssm import DiceLoss, MambaUnet

torch.manual_seed(0)
model = MambaUnet(num_classes=4).cuda().train()
opt, dice = workloads.make_sgd(model), DiceLoss(4)
x = torch.rand(24, 1, 224, 224, device="cuda")
y = torch.randint(0, 4, (24, 224, 224), device="cuda")
res = {}
for rep in range(3):
    for mode in (True, False):
        ss2d.FUSE_DT = mode
        for _ in range(4):
            workloads.supervised_step(model, opt, dice, x, y)
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(15):
            workloads.supervised_step(model, opt, dice, x, y)
        e1.record()
        torch.cuda.synchronize()
        r = res.setdefault("fuse_dt=%s" % mode, {"ms": [], "peak_gb": []})
        r["ms"].append(round(e0.elapsed_time(e1) / 15, 3))
        r["peak_gb"].append(round(torch.cuda.max_memory_allocated() / 2**30, 3))
print(json.dumps(res))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"))
