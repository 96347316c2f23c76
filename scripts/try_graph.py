"""CUDA-graph capture of the MambaUnet steps: does it capture, is the replayed loss equal to the eager one, how fast."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from selscan_b200 import workloads as wl
from selscan_b200.vssm import DiceLoss, MambaUnet
dev = torch.device("cuda")
torch.manual_seed(0)
B = int(os.environ.get("B", 24))
model = MambaUnet(num_classes=4, drop_path_rate=0.0).to(dev).train()
import copy
model2 = copy.deepcopy(model)
opt, opt2 = wl.make_sgd(model), wl.make_sgd(model2)
dice = DiceLoss(4)
x = torch.rand(B, 1, 224, 224, device=dev); y = torch.randint(0, 4, (B, 224, 224), device=dev)
g = wl.GraphedStep(lambda a, b: wl.supervised_step(model, opt, dice, a, b), x, y, warmup=3)
for _ in range(3):
    wl.supervised_step(model2, opt2, dice, x, y)
for i in range(3):
    lg = float(g(x, y)); le = float(wl.supervised_step(model2, opt2, dice, x, y))
    print("step", i, "graph loss", lg, "eager loss", le)
def t(fn, n=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("graph ms", t(lambda: g(x, y)), "eager ms", t(lambda: wl.supervised_step(model2, opt2, dice, x, y)))
