"""The reference's SS2D data path on the GPU -- BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference model sources cannot travel to the GPU box, but its *path* can be timed there: this module re-routes the
SS2D blocks of a `selscan_b200.vssm.MambaUnet` (same architecture / parameters as the reference model, golden-checked in
tests/test_vssm_gpu.py) through

  * the reference's OWN CUDA kernels (`oracle/_ref/selective_scan_cuda.so`, the unmodified sources rebuilt for sm_100a),
    called the way its autograd wrapper calls them (/root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py:14-74), and
  * the ATen chain of `SS2D.forward_corev0` / `SS2D.forward` (/root/reference/code/networks/mamba_sys.py:396-436, :527-540):
    stack / transpose / flip / cat, two einsums, flips / transposes / adds, nn.LayerNorm, nn.Conv2d, F.silu,

with torch's own LayerNorm everywhere and the reference's DiceLoss (per-class `.item()`, code/utils/losses.py:355-368).
That is BASELINE.md rows B2 / B4 / B5 / B6: "MambaUnet with the reference op on this GPU".  bench.py's `reference_cuda`
block and tests/test_vs_reference_cuda_gpu.py are the only users.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ref_cuda


class RefSelectiveScanFn(torch.autograd.Function):
    """selective_scan_interface.py:14-74 on the rebuilt reference extension (real A, variable B / C, optional D / z / bias)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False):
        u, delta = (t if t.stride(-1) == 1 else t.contiguous() for t in (u, delta))          # :19-22
        D = D.contiguous() if D is not None else None                                         # :23-24
        B, C = (t if t.stride(-1) == 1 else t.contiguous() for t in (B, C))                   # :25-28
        if z is not None and z.stride(-1) != 1:                                               # :29-30
            z = z.contiguous()
        ctx.squeeze = (B.dim() == 3, C.dim() == 3)                                            # :31-36
        B = B.unsqueeze(1) if B.dim() == 3 else B
        C = C.unsqueeze(1) if C.dim() == 3 else C
        out, x, *rest = ref_cuda.ref_fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus)  # :37
        ctx.delta_softplus, ctx.has_z = delta_softplus, z is not None
        if z is None:                                                                         # :41-47
            ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
            return out
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, x, out)
        return rest[0]

    @staticmethod
    def backward(ctx, dout):
        if not ctx.has_z:                                                                     # :50-56
            u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
            z = out = None
        else:
            u, delta, A, B, C, D, z, delta_bias, x, out = ctx.saved_tensors
        if dout.stride(-1) != 1:                                                              # :57-58
            dout = dout.contiguous()
        du, ddelta, dA, dB, dC, dD, dbias, *rest = ref_cuda.ref_bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out,
                                                                    ctx.delta_softplus)       # :62-65
        dz = rest[0] if ctx.has_z else None
        dB = dB.squeeze(1) if ctx.squeeze[0] else dB                                          # :67-68
        dC = dC.squeeze(1) if ctx.squeeze[1] else dC
        return (du, ddelta, dA, dB, dC, dD if D is not None else None, dz, dbias if delta_bias is not None else None, None)


def ref_selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False):
    return RefSelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus)


def forward_core_reference(self, x):
    """The tensor program of SS2D.forward_corev0 (mamba_sys.py:396-436) with the reference scan; `self` is an SS2D module."""
    B, D, H, W = x.shape
    L, K, R, N = H * W, 4, self.dt_rank, self.d_state
    rows = x.view(B, -1, L)
    cols = x.transpose(2, 3).contiguous().view(B, -1, L)
    fw = torch.stack([rows, cols], dim=1).view(B, 2, -1, L)                                   # :403
    xs = torch.cat([fw, fw.flip(-1)], dim=1)                                                  # :404  (b, k, d, l)
    x_dbl = torch.einsum("b k d l, k c d -> b k c l", xs.view(B, K, -1, L), self.x_proj_weight)        # :406
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)                                        # :408
    dts = torch.einsum("b k r l, k d r -> b k d l", dts.view(B, K, -1, L), self.dt_projs_weight)        # :409
    out_y = ref_selective_scan_fn(
        xs.float().view(B, -1, L), dts.contiguous().float().view(B, -1, L), -torch.exp(self.A_logs.float()).view(-1, N),
        Bs.float().view(B, K, -1, L), Cs.float().view(B, K, -1, L), self.Ds.float().view(-1), None,
        self.dt_projs_bias.float().view(-1), True).view(B, K, -1, L)                          # :411-426
    inv = out_y[:, 2:4].flip(-1).view(B, 2, -1, L)                                            # :429
    wh = out_y[:, 1].view(B, -1, W, H).transpose(2, 3).contiguous().view(B, -1, L)            # :430
    invwh = inv[:, 1].view(B, -1, W, H).transpose(2, 3).contiguous().view(B, -1, L)           # :431
    y = out_y[:, 0] + inv[:, 0] + wh + invwh                                                  # :432
    y = y.transpose(1, 2).contiguous().view(B, H, W, -1)                                      # :433
    return F.layer_norm(y, (y.shape[-1],), self.out_norm.weight, self.out_norm.bias, self.out_norm.eps).to(x.dtype)   # :434


def ss2d_forward_reference(self, x):
    """SS2D.forward (mamba_sys.py:527-540) with stock torch ops around forward_core_reference."""
    xz = self.in_proj(x)
    xh, z = xz.chunk(2, dim=-1)
    xh = F.silu(self.conv2d(xh.permute(0, 3, 1, 2).contiguous()))
    y = forward_core_reference(self, xh) * F.silu(z)
    return self.out_proj(y)


class RefDiceLoss(nn.Module):
    """code/utils/losses.py:332-368: per-class soft Dice with the `.item()` host read of every class (4 syncs per call)."""

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, probs, target):
        onehot = torch.cat([(target == i).float() for i in range(self.n_classes)], dim=1)
        loss, self.class_dice = 0.0, []
        for i in range(self.n_classes):
            s, t = probs[:, i], onehot[:, i]
            d = 1 - (2 * (s * t).sum() + 1e-5) / ((s * s).sum() + (t * t).sum() + 1e-5)
            self.class_dice.append(1.0 - d.item())
            loss = loss + d
        return loss / self.n_classes


def to_reference_path(model):
    """Re-route every SS2D block and LayerNorm of a selscan_b200.vssm model through the reference path (in place)."""
    from selscan_b200 import vssm

    for m in model.modules():
        if isinstance(m, vssm.SS2D):
            m.forward = ss2d_forward_reference.__get__(m)
        elif isinstance(m, nn.LayerNorm):
            m.forward = nn.LayerNorm.forward.__get__(m)
    return model
