"""Time the REFERENCE's own CUDA kernels (oracle/_ref, sm_100a rebuild) next to ours on the B200, per stage shape.

BASELINE INFRASTRUCTURE: python oracle/bench_ref_cuda.py [--batch 24] [--out file.json]
Same tensors for both, CUDA events, every call streams >= 234 MB so L2 is defeated by the working set.
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mamba-unet_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from bench import STAGES, bytes_bwd, bytes_fwd, measured_peak  # noqa: E402
from oracle import ref_cuda  # noqa: E402
from selscan_b200 import ops  # noqa: E402


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=24)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    dev = torch.device("cuda")
    peak, _ = measured_peak()
    res = {"batch": a.batch, "peak_gbps": peak, "stages": {}}
    g = torch.Generator(device=dev).manual_seed(0)
    for name, d_inner, L, calls in STAGES:
        kd, N, K, b = 4 * d_inner, 16, 4, a.batch
        u = torch.randn(b, kd, L, device=dev, generator=g)
        dt = 0.5 * torch.randn(b, kd, L, device=dev, generator=g)
        A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(kd, 1).contiguous()
        Bm = torch.randn(b, K, N, L, device=dev, generator=g)
        Cm = torch.randn(b, K, N, L, device=dev, generator=g)
        D = torch.ones(kd, device=dev)
        bias = torch.full((kd,), -4.6, device=dev)
        dout = torch.randn(b, kd, L, device=dev, generator=g)
        bf, bb = bytes_fwd(b, kd, L), bytes_bwd(b, kd, L)
        row = {}
        if ref_cuda.available():
            out_r, x_r = ref_cuda.ref_fwd(u, dt, A, Bm, Cm, D, None, bias, True)[:2]
            f = timeit(lambda: ref_cuda.ref_fwd(u, dt, A, Bm, Cm, D, None, bias, True))
            w = timeit(lambda: ref_cuda.ref_bwd(u, dt, A, Bm, Cm, D, None, bias, dout, x_r, None, True))
            row["reference"] = {"fwd_ms": f, "bwd_ms": w, "fwd_gbps": bf / f / 1e6, "bwd_gbps": bb / w / 1e6,
                                "fwdbwd_gbps": (bf + bb) / (f + w) / 1e6, "frac": (bf + bb) / (f + w) / 1e6 / peak}
        # ours: rows with a 16-byte aligned pitch, as the public op / forward_core_b200 allocate them (L = 49 -> pitch 52)
        pad = lambda t: ops.empty_rows(b, kd, L, dev).copy_(t)
        u, dt, dout = pad(u), pad(dt), pad(dout)
        out = ops.empty_rows(b, kd, L, dev)
        ck = torch.empty(max(ops.ckpt_elems(b, kd, L, N), 4), device=dev)
        du, dd = ops.empty_rows(b, kd, L, dev), ops.empty_rows(b, kd, L, dev)
        dA, dB, dC = torch.zeros(kd, N, device=dev), torch.zeros_like(Bm), torch.zeros_like(Cm)
        dD, db = torch.zeros(kd, device=dev), torch.zeros(kd, device=dev)
        f = timeit(lambda: ops.launch_fwd(u, dt, A, Bm, Cm, D, None, bias, True, out, None, None, ck))

        def bwd():
            for t in (dA, dB, dC, dD, db):
                t.zero_()
            ops.launch_bwd(u, dt, A, Bm, Cm, D, None, bias, dout, None, ck, True, du, dd, dA, dB, dC, dD, None, db)

        w = timeit(bwd)
        row["ours"] = {"fwd_ms": f, "bwd_ms": w, "fwd_gbps": bf / f / 1e6, "bwd_gbps": bb / w / 1e6,
                       "fwdbwd_gbps": (bf + bb) / (f + w) / 1e6, "frac": (bf + bb) / (f + w) / 1e6 / peak}
        if "reference" in row:
            row["speedup_fwdbwd"] = (row["reference"]["fwd_ms"] + row["reference"]["bwd_ms"]) / (f + w)
        res["stages"][name] = row
        print(name, json.dumps(row), flush=True)
        del u, dt, Bm, Cm, dout, out, ck, du, dd
        torch.cuda.empty_cache()
    if a.out:
        with open(a.out, "w") as fh:
            json.dump(res, fh, indent=1)


if __name__ == "__main__":
    main()
