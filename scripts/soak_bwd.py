"""Soak test of the tiled kernels' hand-over logic: many launches of stage-shaped calls, outputs without cross-CTA reductions
(out, du, ddelta) must be bit-identical run to run; atomically accumulated gradients must agree to round-off.
    python scripts/soak_bwd.py [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import torch
from selscan_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda")
bad = 0
for (b, kd, L, K) in [(24, 768, 3136, 4), (24, 1536, 784, 4), (24, 3072, 196, 4), (24, 6144, 49, 4), (5, 256, 1001, 4), (3, 64, 77, 1)]:
    N = 16
    torch.manual_seed(b * 1000 + L)
    t = {"u": ops.empty_rows(b, kd, L, dev).normal_(), "delta": ops.empty_rows(b, kd, L, dev).normal_().mul_(0.5),
         "A": -torch.rand(kd, N, device=dev) * 4 - 0.1, "B": torch.randn(b, K, N, L, device=dev), "C": torch.randn(b, K, N, L, device=dev),
         "D": torch.randn(kd, device=dev), "bias": torch.randn(kd, device=dev) - 3, "dout": ops.empty_rows(b, kd, L, dev).normal_(),
         "out": ops.empty_rows(b, kd, L, dev), "ck": torch.empty(max(ops.ckpt_elems(b, kd, L, N), 4), device=dev),
         "du": ops.empty_rows(b, kd, L, dev), "dd": ops.empty_rows(b, kd, L, dev)}
    nbc = b * K * N * L
    flat = torch.zeros(2 * nbc + kd * N + 2 * kd, device=dev)
    dB, dC = flat[:nbc].view(b, K, N, L), flat[nbc:2 * nbc].view(b, K, N, L)
    dA = flat[2 * nbc:2 * nbc + kd * N].view(kd, N)
    dD, db = flat[2 * nbc + kd * N:2 * nbc + kd * N + kd], flat[2 * nbc + kd * N + kd:]
    ref = None
    for it in range(iters):
        t["out"].fill_(float("nan")); t["du"].fill_(float("nan")); t["dd"].fill_(float("nan")); t["ck"].fill_(float("nan"))
        ops.launch_fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], True, t["out"], None, None, t["ck"], None)
        flat.zero_()
        ops.launch_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], t["dout"], None, t["ck"], True,
                       t["du"], t["dd"], dA, dB, dC, dD, None, db)
        torch.cuda.synchronize()
        cur = (t["out"][..., :L].clone(), t["du"][..., :L].clone(), t["dd"][..., :L].clone(), flat.clone())
        assert all(torch.isfinite(c).all() for c in cur), f"non-finite output, shape {(b, kd, L)}, iteration {it}"
        if ref is None:
            ref = cur
            continue
        for name, a_, r_ in zip(("out", "du", "ddelta"), cur[:3], ref[:3]):
            if not torch.equal(a_, r_):
                bad += 1
                print("MISMATCH", name, (b, kd, L), it, float((a_ - r_).abs().max()))
        rel = float((cur[3] - ref[3]).abs().max() / ref[3].abs().max())
        if rel > 1e-5:
            bad += 1
            print("MISMATCH accumulated grads", (b, kd, L), it, rel)
    print("shape", (b, kd, L), "ok" if bad == 0 else "BAD", flush=True)
print("soak:", "PASS" if bad == 0 else f"FAIL ({bad})", ops.bwd_kernel_name())
sys.exit(1 if bad else 0)
