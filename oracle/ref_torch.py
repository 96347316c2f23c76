"""PyTorch-CPU port of the reference's selective_scan_ref -- TEST / BASELINE INFRASTRUCTURE ONLY.

Restates /root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py:86-152 with the same tensor program
(materialised (B, D, L, N) decay and drive tensors, a Python loop over the sequence, autograd for the backward),
because that program -- not just its result -- is what "the reference's CPU path" costs.  bench.py's
cpu_baseline / --impl reference legs time it on the GPU box's host cores; tests check it against the C oracle.
"""
import torch
import torch.nn.functional as F


def selective_scan_ref_torch(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                             return_last_state=False):
    dtype_in = u.dtype
    u = u.float()
    delta = delta.float()
    if delta_bias is not None:                                   # :104-105
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:                                           # :106-107
        delta = F.softplus(delta)
    batch, dim, dstate = u.shape[0], A.shape[0], A.shape[1]
    B = B.float()
    C = C.float()
    x = A.new_zeros((batch, dim, dstate))
    ys = []
    deltaA = torch.exp(torch.einsum("bdl,dn->bdln", delta, A))  # :121
    if B.dim() == 3:                                             # :125-126
        deltaB_u = torch.einsum("bdl,bnl,bdl->bdln", delta, B, u)
    else:                                                        # :128-129
        B = B.repeat_interleave(dim // B.shape[1], dim=1)
        deltaB_u = torch.einsum("bdl,bdnl,bdl->bdln", delta, B, u)
    if C.dim() == 4:                                             # :130-131
        C = C.repeat_interleave(dim // C.shape[1], dim=1)
    last_state = None
    for i in range(u.shape[2]):                                  # :133-146
        x = deltaA[:, :, i] * x + deltaB_u[:, :, i]
        if C.dim() == 3:
            y = torch.einsum("bdn,bn->bd", x, C[:, :, i])
        else:
            y = torch.einsum("bdn,bdn->bd", x, C[:, :, :, i])
        if i == u.shape[2] - 1:
            last_state = x
        ys.append(y)
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u * D.view(-1, 1)              # :148
    if z is not None:                                            # :149-150
        out = out * F.silu(z.float())
    out = out.to(dtype=dtype_in)
    return out if not return_last_state else (out, last_state)
