"""GPU parity at the BASELINE.json sizes and layouts that the smaller cases do not reach ("parity holes"):

  * batch 24 at stage 1 and stage 4: EVERY output and gradient against the fp64 oracle, including the ones accumulated over all 24
    images with atomics (dA, dD, ddelta_bias) and the full dB / dC (reference: selective_scan_bwd_kernel.cuh:298-329,467-477);
  * B / C passed as the strided views of x_dbl that SS2D.forward_core really produces (code/networks/mamba_sys.py:406-415:
    stride(-1) = R + 2N, R = 6 / 12 / 24 / 48 -> pitch 152 / 176 / 224 / 320 bytes) at the four real stage shapes, on the tiled
    kernels (channels per group 192 ... 1536).

Tolerances as everywhere: forward rtol 1e-4 / atol 1e-5 x scale, gradients rtol 1e-3 / atol 1e-4 x scale.
"""
import numpy as np
import pytest
import torch

from test_parity_gpu import BWD_ATOL, BWD_RTOL, FWD_ATOL, FWD_RTOL, _t, check_all, close, run_ours

pytestmark = pytest.mark.gpu

STAGES = {"S1": (192, 3136, 6), "S2": (384, 784, 12), "S3": (768, 196, 24), "S4": (1536, 49, 48)}


def _refs(oracle, inp):
    ref_out, ref_last = oracle.oracle_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                                          inp["delta_bias"], True, return_last_state=True, precision=64)
    ref_g = oracle.oracle_bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"], inp["delta_bias"],
                              inp["dout"], True, precision=64)
    return ref_out, ref_last, ref_g


@pytest.mark.parametrize("stage,dist", [("S4", "M"), ("S4", "T"), ("S1", "M")])
def test_batch24_all_gradients(oracle, stage, dist):
    D, L, _ = STAGES[stage]
    inp = oracle.make_inputs(24, 4 * D, L, 16, 4, dist=dist, seed=2400 + L, has_z=False, has_D=True, has_bias=True)
    out, last, grads = run_ours(inp, True)
    check_all(out, last, grads, *_refs(oracle, inp))


@pytest.mark.parametrize("stage", ["S1", "S2", "S3", "S4"])
def test_xdbl_strided_bc_real_stage_shapes(oracle, stage):
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn
    from selscan_b200 import ops

    D, L, R = STAGES[stage]
    batch, K, N = 2, 4, 16
    inp = oracle.make_inputs(batch, K * D, L, N, K, dist="M", seed=77 + R)
    xdbl = torch.randn(batch, K, L, R + 2 * N, device="cuda")
    xdbl[..., R:R + N] = torch.from_numpy(inp["B"]).cuda().permute(0, 1, 3, 2)
    xdbl[..., R + N:] = torch.from_numpy(inp["C"]).cuda().permute(0, 1, 3, 2)
    xdbl.requires_grad_()
    Bv = xdbl[..., R:R + N].permute(0, 1, 3, 2)
    Cv = xdbl[..., R + N:].permute(0, 1, 3, 2)
    assert Bv.stride(-1) == R + 2 * N and Cv.stride(-1) == R + 2 * N and Bv.stride(-2) == 1
    # rows of u / delta / dout with the padded pitch the model path uses (stage 4: 49 -> 52)
    u, dt = (ops.empty_rows(batch, K * D, L, "cuda").copy_(torch.from_numpy(inp[k]).cuda()).requires_grad_() for k in ("u", "delta"))
    A, Dp, bias = (_t(inp[k]) for k in ("A", "D", "delta_bias"))
    out, last = selective_scan_fn(u, dt, A, Bv, Cv, Dp, z=None, delta_bias=bias, delta_softplus=True, return_last_state=True)
    out.backward(torch.from_numpy(inp["dout"]).cuda())
    torch.cuda.synchronize()
    ref_out, ref_last, ref_g = _refs(oracle, inp)
    close(out.detach().cpu().numpy(), ref_out, FWD_RTOL, FWD_ATOL, "out")
    close(last.cpu().numpy(), ref_last, FWD_RTOL, FWD_ATOL, "last_state")
    gx = xdbl.grad.cpu().numpy()
    close(np.transpose(gx[..., R:R + N], (0, 1, 3, 2)), ref_g["dB"], BWD_RTOL, BWD_ATOL, "dB")
    close(np.transpose(gx[..., R + N:], (0, 1, 3, 2)), ref_g["dC"], BWD_RTOL, BWD_ATOL, "dC")
    assert float(np.abs(gx[..., :R]).max()) == 0.0          # the dt_rank columns of x_dbl are not inputs of the scan
    for name, t in (("du", u), ("ddelta", dt), ("dA", A), ("dD", Dp), ("ddelta_bias", bias)):
        close(t.grad.cpu().numpy(), ref_g[name], BWD_RTOL, BWD_ATOL, name)


@pytest.mark.parametrize("batch", [1, 3])
def test_small_batch_takes_tiled_backward(oracle, batch):
    """batch 1 has no batch stride to speak of: it must still be eligible for the tiled kernels (and match the oracle)."""
    inp = oracle.make_inputs(batch, 256, 333, 16, 2, dist="T", seed=5 + batch)
    out, last, grads = run_ours(inp, True)
    check_all(out, last, grads, *_refs(oracle, inp))
