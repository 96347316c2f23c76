"""GPU parity against the REFERENCE's own CUDA kernels (selective_scan_cuda rebuilt unmodified for sm_100a into
oracle/_ref/ by oracle/build_ref.py), on identical inputs, for the four Mamba-UNet stage shapes.

Both implementations are also compared with the fp64 oracle so that the report says whose error is larger
("require new <= stated tol and report the ratio to the reference CUDA's own error").
Skipped when oracle/_ref/ was not built (it only builds where /root/reference exists)."""
import numpy as np
import pytest
import torch

from oracle import ref_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_cuda.available(), reason="oracle/_ref not built")]


@pytest.mark.parametrize("D,L", [(192, 3136), (384, 784), (768, 196), (1536, 49)])
def test_matches_reference_kernels(oracle, D, L):
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn

    batch, K, N = 2, 4, 16
    inp = oracle.make_inputs(batch, K * D, L, N, K, dist="M", seed=123)
    t = {k: torch.from_numpy(v).cuda() for k, v in inp.items() if v is not None}
    out_r, x_r = ref_cuda.ref_fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["delta_bias"], True)[:2]
    g_r = ref_cuda.ref_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["delta_bias"], t["dout"], x_r,
                           None, True)
    leaves = {k: t[k].clone().requires_grad_() for k in ("u", "delta", "A", "B", "C", "D", "delta_bias")}
    out = selective_scan_fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"], None,
                            leaves["delta_bias"], True)
    out.backward(t["dout"])
    ours = [leaves[k].grad for k in ("u", "delta", "A", "B", "C", "D", "delta_bias")]
    names = ["du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"]

    ref64 = oracle.oracle_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None, inp["delta_bias"],
                              True, precision=64)
    g64 = oracle.oracle_bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None, inp["delta_bias"],
                            inp["dout"], True, precision=64)

    def err(a, ref):
        return float(np.abs(a.astype(np.float64) - ref).max() / max(1.0, np.abs(ref).max()))

    report = {"out": (err(out.detach().cpu().numpy(), ref64), err(out_r.cpu().numpy(), ref64))}
    torch.testing.assert_close(out.detach(), out_r, rtol=1e-4, atol=1e-5 * max(1.0, float(out_r.abs().max())))
    for nm, mine, theirs in zip(names, ours, g_r[:7]):
        scale = max(1.0, float(theirs.abs().max()))
        report[nm] = (err(mine.cpu().numpy(), g64[nm]), err(theirs.float().cpu().numpy().reshape(g64[nm].shape), g64[nm]))
        torch.testing.assert_close(mine, theirs.reshape(mine.shape).float(), rtol=1e-3, atol=1e-4 * scale, msg=lambda m: f"{nm}: {m}")
    print(f"\n[D={D} L={L}] max scaled error vs fp64 oracle (ours, reference CUDA):")
    for k, (e1, e2) in report.items():
        print(f"   {k:12s} {e1:.3e}  {e2:.3e}")
        assert e1 <= 5 * e2 + 1e-6, (k, e1, e2)


def test_reference_path_model_matches_product_model():
    """The baseline arm of bench.py (`reference_cuda.mambaunet`: oracle/ref_model.py = the reference kernels + the ATen chain of
    forward_corev0 + torch LayerNorm) and the product model compute the same function: same parameters, same input, forward and
    input gradient.  Guards the baseline numbers against a mis-stated chain."""
    import copy

    from oracle import ref_model
    from selscan_b200.vssm import VSSM

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(21)
    ours = VSSM(depths=(1, 1, 1, 1), dims=(32, 64, 128, 256), drop_path_rate=0.0).cuda().eval()
    theirs = ref_model.to_reference_path(copy.deepcopy(ours))
    x = torch.rand(2, 3, 64, 64, device="cuda")
    g = torch.randn(2, 4, 64, 64, device="cuda")
    outs = []
    for m in (ours, theirs):
        xi = x.clone().requires_grad_()
        out = m(xi)
        (out * g).sum().backward()
        outs.append((out.detach(), xi.grad))
    torch.testing.assert_close(outs[0][0], outs[1][0], rtol=1e-3, atol=2e-4)
    torch.testing.assert_close(outs[0][1], outs[1][1], rtol=2e-3, atol=2e-4 * float(outs[1][1].abs().max()))
