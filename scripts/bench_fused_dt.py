"""dt_proj inside the scan kernels vs a materialised step tensor: kernel times per stage (batch 24) and the GEMM it replaces.
    python scripts/bench_fused_dt.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from bench import DT_RANK, STAGES
from selscan_b200 import ops

dev, b, K, N = "cuda", 24, 4, 16


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


res = {}
for name, D, L, _ in STAGES:
    R, KD = DT_RANK[name], K * D
    if not ops.dt_fusable(b, KD, L, N, K, R):
        res[name] = {"fusable": False, "dt_rank": R}
        continue
    torch.manual_seed(0)
    x_dbl = torch.randn(b, K, R + 2 * N, L, device=dev)
    dt_w = torch.randn(K, D, R, device=dev) * R ** -0.5
    u, dout = torch.randn(b, KD, L, device=dev), torch.randn(b, KD, L, device=dev)
    A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(KD, 1).contiguous()
    Dp, bias = torch.ones(KD, device=dev), torch.full((KD,), -4.6, device=dev)
    Bv, Cv, dt_x = x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], x_dbl[:, :, :R]
    delta = torch.empty(b, K, D, L, device=dev)
    gemm = timeit(lambda: torch.matmul(dt_w.unsqueeze(0), dt_x, out=delta))
    out, du, dd = (torch.empty(b, KD, L, device=dev) for _ in range(3))
    ck = torch.empty(max(ops.ckpt_elems(b, KD, L, N), 4), device=dev)
    nbc = b * K * N * L
    flat = torch.zeros(2 * nbc + KD * N + 2 * KD, device=dev)
    dB, dC = flat[:nbc].view(b, K, N, L), flat[nbc:2 * nbc].view(b, K, N, L)
    dA = flat[2 * nbc:2 * nbc + KD * N].view(KD, N)
    dD, db = flat[2 * nbc + KD * N:2 * nbc + KD * N + KD], flat[2 * nbc + KD * N + KD:]
    row = {"fusable": True, "dt_rank": R, "dt_proj_gemm_ms": round(gemm, 4)}
    for tag, kw, dl in (("materialised", {}, delta.view(b, KD, L)), ("fused", dict(dt_w=dt_w.view(KD, R), dt_x=dt_x), None)):
        f = timeit(lambda: ops.launch_fwd(u, dl, A, Bv, Cv, Dp, None, bias, True, out, None, None, ck, None, **kw))
        w = timeit(lambda: ops.launch_bwd(u, dl, A, Bv, Cv, Dp, None, bias, dout, None, ck, True, du, dd, dA, dB, dC, dD, None, db, **kw))
        row[tag] = {"fwd_ms": round(f, 4), "bwd_ms": round(w, 4)}
    row["net_ms_per_call"] = round(row["fused"]["fwd_ms"] + row["fused"]["bwd_ms"] - row["materialised"]["fwd_ms"] - row["materialised"]["bwd_ms"] - gemm, 4)
    res[name] = row
    print(name, json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
