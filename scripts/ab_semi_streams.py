"""Same-process A/B of the semi-supervised dual-network step (2 x MambaUnet, batch 16) on one vs two CUDA streams, eager and as a
CUDA graph; also checks that both give the same loss and parameters after a few steps."""
import copy, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from selscan_b200 import workloads as wl
from selscan_b200.vssm import DiceLoss, MambaUnet

dev = "cuda"
torch.manual_seed(0)
x = torch.rand(16, 1, 224, 224, device=dev)
y = torch.randint(0, 4, (16, 224, 224), device=dev)
dice, cw = DiceLoss(4), wl.consistency_weight(3000)
base = [MambaUnet(num_classes=4).to(dev).train() for _ in range(2)]


def timed(fn, n=12, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3)


res, finals = {}, {}
for mode in ("one_stream", "two_streams"):
    side = torch.cuda.Stream() if mode == "two_streams" else None
    m1, m2 = copy.deepcopy(base[0]), copy.deepcopy(base[1])
    o1, o2 = wl.make_sgd(m1), wl.make_sgd(m2)
    losses = [float(wl.semi_step(m1, m2, o1, o2, dice, x, y, 8, cw, side)) for _ in range(3)]
    finals[mode] = (losses, [p.detach().clone() for p in list(m1.parameters())[:6] + list(m2.parameters())[:6]])
    res[mode + "_eager_ms"] = timed(lambda: wl.semi_step(m1, m2, o1, o2, dice, x, y, 8, cw, side))
    g = wl.GraphedStep(lambda a, b: wl.semi_step(m1, m2, o1, o2, dice, a, b, 8, cw, side), x, y)
    res[mode + "_graph_ms"] = timed(lambda: g(x, y))
    del g, m1, m2, o1, o2
    torch.cuda.empty_cache()
la, lb = finals["one_stream"][0], finals["two_streams"][0]
res["losses_one_stream"], res["losses_two_streams"] = la, lb
res["max_param_diff"] = max(float((a - b).abs().max()) for a, b in zip(finals["one_stream"][1], finals["two_streams"][1]))
print(json.dumps(res))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
