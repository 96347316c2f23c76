"""Replay on the GPU of the call the UNCHANGED reference model makes into the op (fixtures: tests/golden/ss2d_call_*.npz, traced
from /root/reference/code/networks/mamba_sys.py:396-436,527-540 by tests/golden/make_golden_call_trace.py):

  * `selective_scan_fn` from this repo's `mamba_ssm.ops.selective_scan_interface`, with the reference's exact positional / keyword
    arguments and tensors rebuilt with the reference's exact strides and storage offsets (x_dbl is laid out (k, b, l, c):
    B / C arrive with batch stride < group stride and position stride R + 2N), against the reference's own result;
  * the SS2D block with the fixture's parameters through `forward_b200` / `forward_core_b200` (what `patch_ss2d` installs on the
    reference class), forward and backward, against the reference block's output and gradients.
"""
import os
import types

import numpy as np
import pytest
import torch

from test_parity_gpu import BWD_ATOL, BWD_RTOL, FWD_ATOL, FWD_RTOL, close

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["ss2d_call_d32_7x7", "ss2d_call_d64_14x14"]


def _load(name):
    with np.load(os.path.join(HERE, "golden", name + ".npz"), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def _strided(g, key):
    """The argument exactly as the reference passed it: same shape, strides and offset inside a storage of the same size."""
    val = torch.from_numpy(g[f"arg.{key}.value"])
    base = torch.zeros(int(g[f"arg.{key}.storage_numel"]), dtype=val.dtype, device="cuda")
    view = base.as_strided(tuple(val.shape), tuple(int(s) for s in g[f"arg.{key}.stride"]), int(g[f"arg.{key}.offset"]))
    view.copy_(val.cuda())
    assert view.stride() == tuple(int(s) for s in g[f"arg.{key}.stride"])
    return view


@pytest.mark.parametrize("name", CASES)
def test_replay_reference_call(name):
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn

    g = _load(name)
    t = {k: _strided(g, k).requires_grad_() for k in ("u", "delta", "A", "B", "C", "D", "delta_bias")}
    assert t["B"].stride(0) < t["B"].stride(1) and t["B"].stride(3) > 1        # the (k, b, l, c) layout of the reference's einsum
    out = selective_scan_fn(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], z=None, delta_bias=t["delta_bias"],
                            delta_softplus=bool(int(g["kw.delta_softplus"])), return_last_state=bool(int(g["kw.return_last_state"])))
    assert out.dtype == torch.float32 and tuple(out.shape) == g["op_out"].shape          # mamba_sys.py:427
    close(out.detach().cpu().numpy(), g["op_out"], FWD_RTOL, FWD_ATOL, "out")
    out.sum().backward()                                                                   # gradients flow to strided leaves
    assert all(torch.isfinite(v.grad).all() for v in t.values())


class _RefLikeSS2D(torch.nn.Module):
    """The attributes of the reference SS2D module (mamba_sys.py:267-338) that the installed forwards read, with the fixture's
    parameters; built WITHOUT this repo's model code (stock nn.Linear / nn.Conv2d / nn.LayerNorm, as in the reference)."""

    def __init__(self, sd):
        super().__init__()
        D2, d_model = sd["in_proj.weight"].shape
        D = D2 // 2
        self.d_model, self.d_inner, self.d_state = d_model, D, sd["A_logs"].shape[1]
        self.dt_rank = sd["dt_projs_weight"].shape[2]
        self.in_proj = torch.nn.Linear(d_model, 2 * D, bias="in_proj.bias" in sd)
        self.conv2d = torch.nn.Conv2d(D, D, 3, padding=1, groups=D, bias=True)
        self.out_norm = torch.nn.LayerNorm(D)
        self.out_proj = torch.nn.Linear(D, d_model, bias="out_proj.bias" in sd)
        for k in ("x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds"):
            setattr(self, k, torch.nn.Parameter(torch.empty(sd[k].shape)))
        self.dropout = None
        self.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("fused", [True, False])
def test_patched_block_matches_reference_block(name, fused):
    from selscan_b200 import ss2d

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    g = _load(name)
    m = _RefLikeSS2D({k[3:]: v for k, v in g.items() if k.startswith("sd.")}).cuda()
    m.forward_core = types.MethodType(ss2d.forward_core_b200, m)       # what `self.forward_core = self.forward_corev0` binds after patch_ss2d
    x = torch.from_numpy(g["x"]).cuda().requires_grad_()
    if fused:
        y = ss2d.forward_b200(m, x)
    else:  # only the core replaced: SS2D.forward itself unchanged (mamba_sys.py:527-540)
        xz = m.in_proj(x)
        xh, z = xz.chunk(2, dim=-1)
        xh = torch.nn.functional.silu(m.conv2d(xh.permute(0, 3, 1, 2).contiguous()))
        y = m.out_proj(m.forward_core(xh) * torch.nn.functional.silu(z))
    close(y.detach().cpu().numpy(), g["y"], 1e-3, 1e-4, "y")
    (y * torch.from_numpy(g["dy"]).cuda()).sum().backward()
    close(x.grad.cpu().numpy(), g["dx"], 2e-3, 2e-4, "dx")
    for n, p in m.named_parameters():
        close(p.grad.cpu().numpy(), g["grad." + n], 2e-3, 5e-4, "grad." + n)
