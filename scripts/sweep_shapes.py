"""BASELINE config 5: selective-scan shape sweep across the 4 stages and batch 1..192 -> HBM roofline curve data.
Writes gpurun_out/roofline_sweep.json.  Inputs rotate through a pool larger than L2 for small working sets."""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from bench import STAGES, bytes_bwd, bytes_fwd, measured_peak
from selscan_b200 import ops

dev = torch.device("cuda")
peak, _ = measured_peak()
rows = []
for name, d_inner, L, _ in STAGES:
    kd, N, K = 4 * d_inner, 16, 4
    for b in (1, 2, 4, 8, 16, 24, 48, 96, 192):
        per_set = 4 * (8 * b * kd * L + 6 * b * K * N * L) + 4 * b * kd * ((L + 7) // 8) * 16
        if per_set > 24e9:
            continue
        nsets = max(1, min(8, int(512e6 // per_set) + 1))          # pool > L2 (126 MB) for small shapes
        sets = []
        for _ in range(nsets):
            t = {"u": ops.empty_rows(b, kd, L, dev).normal_(), "delta": ops.empty_rows(b, kd, L, dev).normal_().mul_(0.5),
                 "A": -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(kd, 1).contiguous(),
                 "B": torch.randn(b, K, N, L, device=dev), "C": torch.randn(b, K, N, L, device=dev),
                 "D": torch.ones(kd, device=dev), "bias": torch.full((kd,), -4.6, device=dev),
                 "dout": ops.empty_rows(b, kd, L, dev).normal_(), "out": ops.empty_rows(b, kd, L, dev),
                 "ck": torch.empty(max(ops.ckpt_elems(b, kd, L, N), 4), device=dev),
                 "du": ops.empty_rows(b, kd, L, dev), "dd": ops.empty_rows(b, kd, L, dev)}
            nbc = b * K * N * L
            t["flat"] = torch.zeros(2 * nbc + kd * N + 2 * kd, device=dev)
            t["dB"], t["dC"] = t["flat"][:nbc].view(b, K, N, L), t["flat"][nbc:2 * nbc].view(b, K, N, L)
            t["dA"] = t["flat"][2 * nbc:2 * nbc + kd * N].view(kd, N)
            t["dD"], t["db"] = t["flat"][2 * nbc + kd * N:2 * nbc + kd * N + kd], t["flat"][2 * nbc + kd * N + kd:]
            n_ws = ops.fwd_workspace_elems(b, kd, L, N, K)
            t["ws"] = torch.empty(n_ws, device=dev) if n_ws > 0 else None   # small batches: segmented forward
            sets.append(t)

        def fwd(t):
            ops.launch_fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], True, t["out"], None, None, t["ck"], t["ws"])

        def bwd(t):
            t["flat"].zero_()
            ops.launch_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], t["dout"], None, t["ck"], True,
                           t["du"], t["dd"], t["dA"], t["dB"], t["dC"], t["dD"], None, t["db"])

        def timeit(fn, iters=12):
            for i in range(3):
                fn(sets[i % nsets])
            torch.cuda.synchronize()
            ts = []
            for i in range(iters):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(sets[i % nsets]); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return statistics.median(ts)

        f, w = timeit(fwd), timeit(bwd)
        bf, bb = bytes_fwd(b, kd, L), bytes_bwd(b, kd, L)
        rows.append({"stage": name, "batch": b, "KD": kd, "L": L, "fwd_ms": round(f, 4), "bwd_ms": round(w, 4),
                     "fwd_gbps": round(bf / f / 1e6, 1), "bwd_gbps": round(bb / w / 1e6, 1),
                     "fwdbwd_gbps": round((bf + bb) / (f + w) / 1e6, 1), "frac_of_peak": round((bf + bb) / (f + w) / 1e6 / peak, 4),
                     "rotating_sets": nsets, "fwd_segmented": sets[0]["ws"] is not None})
        print(rows[-1], flush=True)
        del sets
        torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"peak_gbps": peak, "rows": rows}, open(os.path.join(ROOT, "gpurun_out", "roofline_sweep.json"), "w"), indent=1)
