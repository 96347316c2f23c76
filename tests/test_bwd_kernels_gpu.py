"""GPU parity of the tiled backward kernel (selscan_bwd_ws.cu: warp-specialised) and of its hand-over to the generic kernel
(selscan_bwd.cu: mixed B/C position strides, SELSCAN_B200_GENERIC=1), through the public op -> C ABI, against the fp64 oracle.  Shapes here all have channels-per-group % 64 == 0 and L > 8, i.e. they take the tiled path
(replaces /root/reference/mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489).

Tolerances as in test_parity_gpu.py: forward rtol 1e-4 / atol 1e-5 x scale, gradients rtol 1e-3 / atol 1e-4 x scale.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from test_parity_gpu import BWD_ATOL, BWD_RTOL, FWD_ATOL, FWD_RTOL, _t, check_all, close, run_ours

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_refs(oracle, inp, softplus=True):
    ref_out, ref_last = oracle.oracle_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                                          inp["delta_bias"], softplus, return_last_state=True, precision=64)
    ref_g = oracle.oracle_bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"], inp["delta_bias"],
                              inp["dout"], softplus, precision=64)
    return ref_out, ref_last, ref_g


def test_default_backward_kernel_is_warp_specialised():
    from selscan_b200 import ops

    assert ops.bwd_kernel_name() == "selscan_bwd_ws_kernel"


# chunk boundaries of the backward walk (chunks of 8 positions, halves of 4): partial last chunks, partial halves, 2..N chunks
@pytest.mark.parametrize("L", [9, 12, 13, 16, 17, 23, 24, 25, 50, 100, 257])
def test_tiled_backward_uneven_lengths(oracle, L):
    inp = oracle.make_inputs(2, 128, L, 16, 2, dist="M", seed=300 + L, has_z=False, has_D=True, has_bias=True)
    out, last, grads = run_ours(inp, True)
    check_all(out, last, grads, *_oracle_refs(oracle, inp))


@pytest.mark.parametrize("N", [1, 5, 8, 12, 16])
@pytest.mark.parametrize("softplus,has_D,has_bias", [(True, True, True), (False, False, False), (True, False, True)])
def test_tiled_backward_states_and_optionals(oracle, N, softplus, has_D, has_bias):
    inp = oracle.make_inputs(3, 64, 77, N, 1, dist="T", seed=40 + N, has_z=False, has_D=has_D, has_bias=has_bias)
    out, last, grads = run_ours(inp, softplus)
    check_all(out, last, grads, *_oracle_refs(oracle, inp, softplus))


def test_tiled_backward_softplus_threshold(oracle):
    """delta + bias beyond +-20: the softplus cut-off (fwd_kernel.cuh:153-156) and its derivative (bwd_kernel.cuh:446-450)."""
    inp = oracle.make_inputs(2, 64, 40, 16, 1, dist="T", seed=7, has_z=False, has_D=True, has_bias=True)
    rng = np.random.default_rng(3)
    inp["delta"] = (inp["delta"] + rng.choice([-30.0, -21.0, 0.0, 19.5, 20.5, 35.0], size=inp["delta"].shape)).astype(np.float32)
    inp["A"] = (inp["A"] * 0.02).astype(np.float32)   # keep exp(delta * A) in range for delta ~ 35
    out, last, grads = run_ours(inp, True)
    check_all(out, last, grads, *_oracle_refs(oracle, inp))


def _xdbl_views(inp, R, mixed):
    """B / C as permuted views of an x_dbl-like (batch, K, L, R + 2N) tensor (stride(-1) = R + 2N); mixed: C stays (N, L)-major."""
    B_np, C_np = inp["B"], inp["C"]
    batch, K, N, L = B_np.shape
    xdbl = torch.randn(batch, K, L, R + 2 * N, device="cuda")
    xdbl[..., R:R + N] = torch.from_numpy(B_np).cuda().permute(0, 1, 3, 2)
    xdbl[..., R + N:] = torch.from_numpy(C_np).cuda().permute(0, 1, 3, 2)
    xdbl.requires_grad_()
    Bv = xdbl[..., R:R + N].permute(0, 1, 3, 2)
    Cv = _t(C_np) if mixed else xdbl[..., R + N:].permute(0, 1, 3, 2)
    return xdbl, Bv, Cv


@pytest.mark.parametrize("mixed", [False, True])
def test_tiled_backward_xdbl_layout_and_mixed_strides(oracle, mixed):
    """l-major B/C (the layout SS2D really produces) on the tiled path; with different position strides for B and C the
    warp-specialised kernel is not eligible and the generic kernel runs."""
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn

    batch, K, D, L, N, R = 2, 2, 64, 83, 16, 6
    inp = oracle.make_inputs(batch, K * D, L, N, K, dist="M", seed=55)
    xdbl, Bv, Cv = _xdbl_views(inp, R, mixed)
    assert Bv.stride(-1) == R + 2 * N and (Cv.stride(-1) == 1) == mixed
    u, dt, A, Dp, bias = (_t(inp[k]) for k in ("u", "delta", "A", "D", "delta_bias"))
    out = selective_scan_fn(u, dt, A, Bv, Cv, Dp, z=None, delta_bias=bias, delta_softplus=True)
    out.backward(torch.from_numpy(inp["dout"]).cuda())
    ref_out, _, ref_g = _oracle_refs(oracle, inp)
    close(out.detach().cpu().numpy(), ref_out, FWD_RTOL, FWD_ATOL, "out")
    gx = xdbl.grad.cpu().numpy()
    close(np.transpose(gx[..., R:R + N], (0, 1, 3, 2)), ref_g["dB"], BWD_RTOL, BWD_ATOL, "dB")
    gC = Cv.grad.cpu().numpy() if mixed else np.transpose(gx[..., R + N:], (0, 1, 3, 2))
    close(gC, ref_g["dC"], BWD_RTOL, BWD_ATOL, "dC")
    for name, t in (("du", u), ("ddelta", dt), ("dA", A), ("dD", Dp), ("ddelta_bias", bias)):
        close(t.grad.cpu().numpy(), ref_g[name], BWD_RTOL, BWD_ATOL, name)


_CHILD = r"""
import sys, numpy as np, torch
sys.path[:0] = [{root!r}, {pkg!r}]
import oracle
from mamba_ssm.ops.selective_scan_interface import selective_scan_fn
from selscan_b200 import ops
inp = oracle.make_inputs(2, 768, 200, 16, 4, dist="M", seed=11)
t = {{k: (torch.from_numpy(v).cuda().requires_grad_(k != "dout") if v is not None else None) for k, v in inp.items()}}
out = selective_scan_fn(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], z=None, delta_bias=t["delta_bias"], delta_softplus=True)
out.backward(t["dout"])
torch.cuda.synchronize()
np.savez(sys.argv[1], kernel=ops.bwd_kernel_name(), **{{k: t[k].grad.cpu().numpy() for k in ("u", "delta", "A", "B", "C", "D", "delta_bias")}})
"""


def test_both_tiled_backward_kernels_agree(oracle, tmp_path):
    """The same seeded stage-1-like call (192 channels per group) through the warp-specialised tiled kernel and the generic
    kernel (SELSCAN_B200_GENERIC=1), each in its own process (the choice is made once per process): both within tolerance of the
    oracle and within 2e-5 of each other."""
    inp = oracle.make_inputs(2, 768, 200, 16, 4, dist="M", seed=11)
    _, _, ref_g = _oracle_refs(oracle, inp)
    res = {}
    for mode in ("ws", "generic"):
        path = str(tmp_path / f"g_{mode}.npz")
        env = dict(os.environ, SELSCAN_B200_GENERIC="1" if mode == "generic" else "0")
        code = _CHILD.format(root=ROOT, pkg=os.path.join(ROOT, "mamba-unet_b200"))
        subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=600)
        with np.load(path) as f:
            res[mode] = {k: f[k] for k in f.files}
        for k, name in (("u", "du"), ("delta", "ddelta"), ("A", "dA"), ("B", "dB"), ("C", "dC"), ("D", "dD"), ("delta_bias", "ddelta_bias")):
            close(res[mode][k].reshape(ref_g[name].shape), ref_g[name], BWD_RTOL, BWD_ATOL, f"{mode}:{name}")
    for k in ("u", "delta", "A", "B", "C", "D", "delta_bias"):
        a, b = res["ws"][k].astype(np.float64), res["generic"][k].astype(np.float64)
        scale = max(1.0, float(np.abs(b).max()))
        assert float(np.abs(a - b).max()) <= 2e-5 * scale, k


def test_tiled_backward_repeatable_up_to_atomics(oracle):
    """du and ddelta have no cross-CTA reduction: two runs are bit-identical; the atomically accumulated gradients agree to
    round-off."""
    inp = oracle.make_inputs(4, 256, 120, 16, 4, dist="T", seed=91, has_z=False, has_D=True, has_bias=True)
    _, _, g1 = run_ours(inp, True)
    _, _, g2 = run_ours(inp, True)
    assert np.array_equal(g1["du"], g2["du"]) and np.array_equal(g1["ddelta"], g2["ddelta"])
    for k in ("dA", "dB", "dC", "dD", "ddelta_bias"):
        close(g1[k], g2[k], 1e-5, 1e-6, k)


def test_persistent_forward_many_items(oracle):
    """Shapes with more work items than resident CTAs of the persistent forward (selscan_fwd_tma.cu: a CTA walks items i, i + grid, ...
    with the producer running ahead across item boundaries; parameters of the next item are prefetched): outputs, last state and --
    through the saved states it writes -- the backward still match the oracle."""
    import torch
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn
    for tag, (batch, dim, L, G) in {"a": (24, 256, 200, 4), "b": (30, 128, 49, 2), "c": (20, 64, 13, 1), "d": (40, 1536, 49, 4)}.items():
        inp = oracle.make_inputs(batch, dim, L, 16, G, dist="M", seed=17)
        t = {k: (torch.from_numpy(v).cuda().requires_grad_(k != "dout") if v is not None else None) for k, v in inp.items()}
        out, last = selective_scan_fn(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], z=None, delta_bias=t["delta_bias"],
                                      delta_softplus=True, return_last_state=True)
        out.backward(t["dout"])
        torch.cuda.synchronize()
        ref_out, ref_last, ref_g = _oracle_refs(oracle, inp)
        close(out.detach().cpu().numpy(), ref_out, FWD_RTOL, FWD_ATOL, tag + ":out")
        close(last.cpu().numpy(), ref_last, FWD_RTOL, FWD_ATOL, tag + ":last_state")
        close(t["u"].grad.cpu().numpy(), ref_g["du"], BWD_RTOL, BWD_ATOL, tag + ":du")


def test_soak_handover_logic():
    """scripts/soak_bwd.py: repeated launches at the four stage shapes (batch 24) and two ragged ones; out / du / ddelta must be
    bit-identical run to run and never NaN (buffers are poisoned before every launch) -- a race in the mbarrier hand-over of the
    warp-specialised kernels would show up here."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "soak_bwd.py"), "8"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "soak: PASS" in r.stdout


@pytest.mark.parametrize("batch,dim,L,G", [(2, 128, 9, 2), (2, 128, 50, 2), (1, 64, 257, 1), (2, 768, 200, 4), (3, 256, 49, 4)])
def test_tiled_backward_with_z_gate(oracle, batch, dim, L, G):
    """SiLU(z) gating on the tiled kernels (reference: selective_scan_fwd_kernel.cuh:280-298, selective_scan_bwd_kernel.cuh:171-207):
    out * silu(z) forward; dz and the gated dout in the helper warps of the warp-specialised backward.  All gradients incl. dz."""
    inp = oracle.make_inputs(batch, dim, L, 16, G, dist="T", seed=900 + L, has_z=True, has_D=True, has_bias=True)
    out, last, grads = run_ours(inp, True)
    assert grads["dz"] is not None
    check_all(out, last, grads, *_oracle_refs(oracle, inp))


def test_accumulator_alignment_fallback():
    """dB / dC leave as TMA reduce-adds only where their rows are 16-byte aligned (selscan_bwd_ws.cu `tma_bc`); a caller may hand
    over accumulators that are not (views at odd offsets; the ABI asks for alignment of `ckpt` only): the kernel then takes scalar
    REDs, with the same results."""
    from selscan_b200 import ops

    torch.manual_seed(5)
    dev, b, kd, L, K, N = "cuda", 2, 256, 200, 4, 16
    u, delta = torch.randn(b, kd, L, device=dev), torch.randn(b, kd, L, device=dev) * 0.5
    A = -torch.rand(kd, N, device=dev) - 0.5
    B, C = torch.randn(b, K, N, L, device=dev), torch.randn(b, K, N, L, device=dev)
    D, bias, dout = torch.randn(kd, device=dev), torch.randn(kd, device=dev) - 2, torch.randn(b, kd, L, device=dev)
    nbc, n_ck = b * K * N * L, ops.ckpt_elems(b, kd, L, N)
    res = []
    for off in (0, 1):       # element offset of the dB | dC blocks: 0 = aligned, 1 = 4 bytes off
        out, du, dd = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
        ck = torch.empty(max(n_ck, 4), device=dev)
        flat = torch.zeros(2 * nbc + 8, device=dev)
        dB, dC = flat[off:off + nbc].view(b, K, N, L), flat[off + nbc + 4:off + 2 * nbc + 4].view(b, K, N, L)
        small = torch.zeros(kd * N + 2 * kd, device=dev)
        dA, dD, db = small[:kd * N].view(kd, N), small[kd * N:kd * N + kd], small[kd * N + kd:]
        ops.launch_fwd(u, delta, A, B, C, D, None, bias, True, out, None, None, ck, None)
        ops.launch_bwd(u, delta, A, B, C, D, None, bias, dout, None, ck, True, du, dd, dA, dB, dC, dD, None, db)
        res.append([t.clone() for t in (out, du, dd, dA, dB, dC, dD, db)])
    for name, a, r in zip(("out", "du", "ddelta", "dA", "dB", "dC", "dD", "ddelta_bias"), res[1], res[0]):
        scale = max(1.0, float(r.abs().max()))
        torch.testing.assert_close(a, r, rtol=BWD_RTOL, atol=BWD_ATOL * scale, msg=lambda m: f"{name}: {m}")
