"""Time the four SS2D edge kernels at the stage shapes of MambaUnet (batch 24) against their minimum HBM traffic.
Not part of the product."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from selscan_b200 import ss2d

dev = torch.device("cuda")
B = int(os.environ.get("EDGE_BATCH", 24))
STAGES = {"S1": (192, 56), "S2": (384, 28), "S3": (768, 14), "S4": (1536, 7)}
if os.environ.get("EDGE_STAGES"):
    STAGES = {k: v for k, v in STAGES.items() if k in os.environ["EDGE_STAGES"].split(",")}
PROFILE = bool(os.environ.get("EDGE_PROFILE"))      # one launch of each kernel per stage (for ncu)
PEAK = 6550.7


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n


rows = []
for name, (D, H) in STAGES.items():
    W, L = H, H * H
    unit = B * D * L * 4 / 1e6          # MB of one (B, D, L) fp32 tensor
    # rotate over enough buffer sets to defeat L2 at the small stages
    nset = max(1, int(400 / (unit * 8)) + 1)
    sets = []
    for _ in range(nset):
        xz = torch.randn(B, H, W, 2 * D, device=dev)
        cw, cb = torch.randn(D, 1, 3, 3, device=dev) * 0.3, torch.randn(D, device=dev) * 0.1
        gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
        xs = ss2d.edge_in_fwd(xz, D, cw, cb)
        out, xhat, rstd = ss2d.edge_out_fwd(xs, H, W, xz.data_ptr() + 4 * D, 2 * D, gam, bet, 1e-5, True)
        sets.append(dict(xz=xz, cw=cw, cb=cb, gam=gam, bet=bet, xs=xs, out=out, xhat=xhat, rstd=rstd, dxz=torch.empty_like(xz)))
    it = [0]

    def nxt():
        it[0] = (it[0] + 1) % nset
        return sets[it[0]]

    def f_in_fwd():
        s = nxt(); ss2d.edge_in_fwd(s["xz"], D, s["cw"], s["cb"])

    def f_in_bwd():
        s = nxt(); ss2d.edge_in_bwd(s["xs"], s["xz"], D, s["cw"], s["cb"], s["dxz"])

    def f_out_fwd():
        s = nxt(); ss2d.edge_out_fwd(s["xs"], H, W, s["xz"].data_ptr() + 4 * D, 2 * D, s["gam"], s["bet"], 1e-5, True)

    def f_out_bwd():
        s = nxt(); ss2d.edge_out_bwd(s["out"], H, W, s["xz"].data_ptr() + 4 * D, 2 * D, s["xhat"], s["rstd"], s["gam"], s["bet"],
                                     s["dxz"].data_ptr() + 4 * D, 2 * D)

    for kname, fn, units in (("in_fwd", f_in_fwd, 5), ("in_bwd", f_in_bwd, 6), ("out_fwd", f_out_fwd, 7), ("out_bwd", f_out_bwd, 8)):
        if PROFILE:
            fn()
            torch.cuda.synchronize()
            continue
        ms = timeit(fn)
        gbps = units * unit / 1e3 / (ms / 1e3)
        rows.append(dict(stage=name, kernel=kname, ms=round(ms, 4), min_mb=round(units * unit, 1), gbps=round(gbps, 1),
                         frac_of_peak=round(gbps / PEAK, 3)))
        print(rows[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "edges_bench.json"), "w"), indent=1)
