"""Generate tests/golden/*.npz from the REFERENCE's own selective_scan_ref (CPU, autograd backward).

Runs only in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

It loads /root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py BY FILE PATH with stub modules
for its top-level native imports (causal_conv1d, causal_conv1d_cuda, selective_scan_cuda -- lines 9-11 of
that file), calls selective_scan_ref (:86-152) on seeded inputs, backpropagates a seeded dout through
torch autograd, and stores inputs + outputs + gradients as small fp32 fixtures.  The fixtures travel to
the GPU box; the reference does not.  Cases mirror mamba/tests/ops/test_selective_scan.py:17-88 (dim 4,
dstate 8, groups 1/2, z, D, bias, softplus) plus the Mamba-UNet call pattern (N=16, G=4, z=None, uneven L).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py"
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.oracle import make_inputs  # noqa: E402


def load_reference():
    for name in ("causal_conv1d", "causal_conv1d_cuda", "selective_scan_cuda"):
        m = types.ModuleType(name)
        m.causal_conv1d_fn = None
        sys.modules.setdefault(name, m)
    spec = importlib.util.spec_from_file_location("_ref_selective_scan_interface", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# name: (batch, dim, L, N, G, dist, seed, has_z, has_D, has_bias, softplus, squeeze_BC)
CASES = {
    "ref_test_g1_L128": (2, 4, 128, 8, 1, "T", 0, True, True, True, True, True),
    "ref_test_g2_L256": (2, 4, 256, 8, 2, "T", 1, True, True, True, True, False),
    "unet_like_L49": (2, 32, 49, 16, 4, "M", 2, False, True, True, True, False),
    "unet_like_L196": (1, 16, 196, 16, 4, "M", 3, False, True, True, True, False),
    "unet_like_T_L100": (2, 8, 100, 16, 4, "T", 4, False, True, True, True, False),
    "plain_noD_nobias_nosoftplus_L37": (1, 6, 37, 16, 2, "T", 5, False, False, False, False, False),
    "z_noD_L64": (1, 4, 64, 16, 1, "T", 6, True, False, True, True, True),
    "big_delta_cutoff_L33": (1, 4, 33, 16, 2, "T", 7, False, True, True, True, False),
}


def main():
    ref = load_reference()
    torch.set_num_threads(4)
    for name, (batch, dim, L, N, G, dist, seed, has_z, has_D, has_bias, softplus, squeeze) in CASES.items():
        inp = make_inputs(batch, dim, L, N, G, dist=dist, seed=seed, has_z=has_z, has_D=has_D,
                          has_bias=has_bias)
        if name.startswith("big_delta"):
            # push some (delta + bias) past the softplus threshold 20 (fwd_kernel.cuh:155) and far below 0
            inp["delta"][:, :, ::5] = 25.0
            inp["delta"][:, :, 1::7] = -30.0
            inp["A"] *= 0.01
        t = {k: (torch.from_numpy(v.copy()).requires_grad_(k != "dout") if v is not None else None)
             for k, v in inp.items()}
        Bt, Ct = t["B"], t["C"]
        if squeeze:  # (batch, N, L) form, selective_scan_interface.py:125-126,138-139
            Bt = Bt.detach()[:, 0].clone().requires_grad_()
            Ct = Ct.detach()[:, 0].clone().requires_grad_()
        out, last = ref.selective_scan_ref(t["u"], t["delta"], t["A"], Bt, Ct, t["D"], z=t["z"],
                                           delta_bias=t["delta_bias"], delta_softplus=softplus,
                                           return_last_state=True)
        out.backward(t["dout"])
        save = {k: v for k, v in inp.items() if v is not None}
        if squeeze:
            save["B"] = save["B"][:, 0]
            save["C"] = save["C"][:, 0]
        save.update(out=out.detach().numpy(), last_state=last.detach().numpy(),
                    du=t["u"].grad.numpy(), ddelta=t["delta"].grad.numpy(), dA=t["A"].grad.numpy(),
                    dB=Bt.grad.numpy(), dC=Ct.grad.numpy(), delta_softplus=np.array(int(softplus)))
        if has_D:
            save["dD"] = t["D"].grad.numpy()
        if has_z:
            save["dz"] = t["z"].grad.numpy()
        if has_bias:
            save["ddelta_bias"] = t["delta_bias"].grad.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **{k: np.asarray(v) for k, v in save.items()})
        print(f"{name}: wrote {os.path.getsize(path) / 1024:.1f} KiB  |out|max={np.abs(save['out']).max():.3f}")


if __name__ == "__main__":
    main()
