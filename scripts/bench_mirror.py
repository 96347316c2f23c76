"""Mirrored direction pairs: scan kernels (mirror vs plain) and edge kernels (2 vs 4 planes) per stage, batch 24."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from bench import DT_RANK, STAGES
from selscan_b200 import ops, ss2d

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
    return round(e0.elapsed_time(e1) / iters, 4)


res = {}
for name, D, L, _ in STAGES[:3]:
    H = int(L ** 0.5)
    R, KD = DT_RANK[name], K * D
    torch.manual_seed(0)
    x_dbl = torch.randn(b, K, R + 2 * N, L, device=dev)
    delta = torch.randn(b, KD, L, device=dev) * 0.5
    A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(KD, 1).contiguous()
    Dp, bias = torch.ones(KD, device=dev), torch.full((KD,), -4.6, device=dev)
    Bv, Cv = x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:]
    ck = torch.empty(max(ops.ckpt_elems(b, KD, L, N), 4), device=dev)
    nbc = b * K * N * L
    flat = torch.zeros(2 * nbc + KD * N + 2 * KD, device=dev)
    dB, dC = flat[:nbc].view(b, K, N, L), flat[nbc:2 * nbc].view(b, K, N, L)
    dA = flat[2 * nbc:2 * nbc + KD * N].view(KD, N)
    dD, db = flat[2 * nbc + KD * N:2 * nbc + KD * N + KD], flat[2 * nbc + KD * N + KD:]
    dd = torch.empty(b, KD, L, device=dev)
    row = {}
    for mir in (False, True):
        rows = KD // 2 if mir else KD
        u, dout = torch.randn(b, rows, L, device=dev), torch.randn(b, rows, L, device=dev)
        out, du = torch.zeros(b, rows, L, device=dev), torch.zeros(b, rows, L, device=dev)
        f = timeit(lambda: ops.launch_fwd(u, delta, A, Bv, Cv, Dp, None, bias, True, out, None, None, ck, None, mirror_pairs=mir))
        w = timeit(lambda: ops.launch_bwd(u, delta, A, Bv, Cv, Dp, None, bias, dout, None, ck, True, du, dd, dA, dB, dC, dD, None, db, mirror_pairs=mir))
        z = timeit(lambda: (out.zero_(), du.zero_())) if mir else 0.0
        row["mirror" if mir else "plain"] = {"fwd_ms": f, "bwd_ms": w, "memsets_ms": z}
    xz = torch.randn(b, H, H, 2 * D, device=dev)
    cw, cb = torch.randn(D, 1, 3, 3, device=dev) * 0.3, torch.randn(D, device=dev) * 0.1
    gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    for n in (4, 2):
        xs = ss2d.edge_in_fwd(xz, D, cw, cb, n)
        o, xhat, rstd = ss2d.edge_out_fwd(xs, H, H, xz.data_ptr() + 4 * D, 2 * D, gam, bet, 1e-5, True)
        dxz = torch.empty_like(xz)
        row[f"edges_{n}planes"] = {
            "in_fwd": timeit(lambda: ss2d.edge_in_fwd(xz, D, cw, cb, n)),
            "in_bwd": timeit(lambda: ss2d.edge_in_bwd(xs, xz, D, cw, cb, dxz)),
            "out_fwd": timeit(lambda: ss2d.edge_out_fwd(xs, H, H, xz.data_ptr() + 4 * D, 2 * D, gam, bet, 1e-5, True)),
            "out_bwd": timeit(lambda: ss2d.edge_out_bwd(o, H, H, xz.data_ptr() + 4 * D, 2 * D, xhat, rstd, gam, bet, dxz.data_ptr() + 4 * D, 2 * D, n))}
    res[name] = row
    print(name, json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
