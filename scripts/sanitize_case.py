"""Tiny fwd+bwd through the public op on shapes that hit the tiled kernels (dpg 64) and the generic ones (dpg 3, z), for
compute-sanitizer.  Checks against the oracle too."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "mamba-unet_b200")]
import numpy as np, torch
import oracle
from mamba_ssm.ops.selective_scan_interface import selective_scan_fn

for (batch, dim, L, N, G, has_z) in [(2, 256, 41, 16, 4, False), (1, 128, 100, 16, 2, False), (2, 6, 19, 8, 2, True), (1, 4, 9, 40, 1, False)]:
    inp = oracle.make_inputs(batch, dim, L, N, G, dist="M", seed=1, has_z=has_z)
    t = {k: (torch.from_numpy(v).cuda().requires_grad_(k != "dout") if v is not None else None) for k, v in inp.items()}
    out, last = selective_scan_fn(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["z"], t["delta_bias"], True, True)
    out.backward(t["dout"])
    torch.cuda.synchronize()
    ref = oracle.oracle_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"], inp["delta_bias"], True)
    g = oracle.oracle_bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"], inp["delta_bias"], inp["dout"], True)
    assert np.allclose(out.detach().cpu().numpy(), ref, rtol=1e-4, atol=1e-4), "out"
    assert np.allclose(t["u"].grad.cpu().numpy(), g["du"], rtol=1e-3, atol=1e-3), "du"
    assert np.allclose(t["B"].grad.cpu().numpy(), g["dB"], rtol=1e-3, atol=1e-3), "dB"
    print("ok", (batch, dim, L, N, G, has_z))
