"""Host-side mirror of the reference op interface for the selective scan.

Mirrors /root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py:
  SelectiveScanFn   :14-74   (autograd forward/backward around the native kernels)
  selective_scan_fn :77-83
  selective_scan_ref:86-152  (the semantic spec, plain PyTorch; exported because callers import it)
Same names, argument meaning, return values and error class (RuntimeError); the native calls go to
libselscan_b200.so through its C ABI (include/selscan_b200.h) instead of the pybind `selective_scan_cuda`.
selective_scan_fn has no fallback: without a CUDA tensor or without the library it raises.
"""

import torch
import torch.nn.functional as F

from . import _lib


def _p(t):
    return None if t is None else t.data_ptr()


def _check_inputs(u, delta, A, B, C, D, z, delta_bias):
    """Shape / dtype / device rules of selective_scan.cpp:233-305 (TORCH_CHECK -> RuntimeError)."""
    if not u.is_cuda:
        raise RuntimeError("selective_scan_fn: u must be a CUDA tensor (this build has no CPU fallback; "
                           "use selective_scan_ref for a CPU reference)")
    for name, t in (("delta", delta), ("A", A), ("B", B), ("C", C), ("D", D), ("z", z), ("delta_bias", delta_bias)):
        if t is not None and t.device != u.device:
            raise RuntimeError(f"selective_scan_fn: {name} must be on the same device as u")
    if u.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        raise RuntimeError(f"selective_scan_fn: unsupported input type {u.dtype}")
    if delta.dtype != u.dtype:
        raise RuntimeError("selective_scan_fn: delta must have the same dtype as u")
    if A.is_complex():
        raise RuntimeError("selective_scan_fn: complex A is not supported by the B200 build (Mamba-UNet uses real A)")
    if A.dtype != torch.float32:
        raise RuntimeError("selective_scan_fn: A must be float32")
    if u.dim() != 3 or delta.shape != u.shape:
        raise RuntimeError(f"selective_scan_fn: u and delta must both be (batch, dim, seqlen); got {tuple(u.shape)} "
                           f"and {tuple(delta.shape)}")
    batch, dim, seqlen = u.shape
    if A.dim() != 2 or A.shape[0] != dim:
        raise RuntimeError(f"selective_scan_fn: A must be (dim, dstate) = ({dim}, N); got {tuple(A.shape)}")
    dstate = A.shape[1]
    if dstate > 256:
        raise RuntimeError("selective_scan only supports state dimension <= 256")
    for name, t in (("B", B), ("C", C)):
        if t.dim() == 2:   # constant (dim, dstate) weights (selective_scan.cpp:238-246): fp32, one row per channel
            if tuple(t.shape) != (dim, dstate) or t.dtype != torch.float32:
                raise RuntimeError(f"selective_scan_fn: constant {name} must be float32 of shape ({dim}, {dstate}); got "
                                   f"{t.dtype} {tuple(t.shape)}")
            continue
        if t.dim() not in (3, 4):
            raise RuntimeError(f"selective_scan_fn: {name} must have 2, 3 or 4 dimensions")
        if t.dtype != u.dtype:
            raise RuntimeError(f"selective_scan_fn: {name} must have the same dtype as u")
        want = (batch, dstate, seqlen) if t.dim() == 3 else (batch, t.shape[1], dstate, seqlen)
        if tuple(t.shape) != want:
            raise RuntimeError(f"selective_scan_fn: {name} has shape {tuple(t.shape)}, expected {want}")
    for name, t in (("D", D), ("delta_bias", delta_bias)):
        if t is not None:
            if t.dtype != torch.float32:
                raise RuntimeError(f"selective_scan_fn: {name} must be float32")
            if tuple(t.shape) != (dim,):
                raise RuntimeError(f"selective_scan_fn: {name} must have shape ({dim},); got {tuple(t.shape)}")
    if z is not None and (z.dtype != u.dtype or z.shape != u.shape):
        raise RuntimeError("selective_scan_fn: z must have the shape and dtype of u")
    return batch, dim, seqlen, dstate


def empty_rows(batch, dim, seqlen, device):
    """(batch, dim, seqlen) fp32 with the row pitch rounded up to 4 floats: rows stay 16-byte aligned for uneven seqlen
    (stage 4 of Mamba-UNet has L = 49), which is what the TMA-staged kernels need.  Contiguous when seqlen % 4 == 0."""
    pitch = (seqlen + 3) // 4 * 4
    buf = torch.empty((batch, dim, pitch), device=device, dtype=torch.float32)
    return buf if pitch == seqlen else buf[:, :, :seqlen]


def _rowmajor(t):
    """Unit stride along seqlen (selective_scan_interface.py:19-22,29-30) AND 16-byte aligned rows: a tensor that only misses
    the alignment (e.g. a contiguous L = 49 tensor) is copied once into a padded-pitch buffer so that it can take the tiled kernels."""
    if (t.stride(-1) == 1 and t.data_ptr() % 16 == 0 and t.stride(1) % 4 == 0 and (t.shape[0] == 1 or t.stride(0) % 4 == 0)):
        return t
    out = empty_rows(t.shape[0], t.shape[1], t.shape[2], t.device)
    out.copy_(t)
    return out


def _strides2(t):
    return (0, 0) if t is None else (t.stride(0), t.stride(1))


def _require_f32_cuda(**tensors):
    """The C ABI takes raw pointers and element strides: a tensor of another dtype (e.g. a half produced under autocast) or on
    another device must never reach it silently."""
    dev = None
    for name, t in tensors.items():
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_cuda:
            raise RuntimeError(f"selscan_b200: {name} must be a float32 CUDA tensor, got {t.dtype} on {t.device}")
        dev = dev or t.device
        if t.device != dev:
            raise RuntimeError(f"selscan_b200: {name} is on {t.device}, expected {dev}")


def _dt_fields(dt_w, dt_x, mirror_pairs=False):
    """Tail of both argument structs: the fused dt_proj inputs (dt_w (dim, R) rows, dt_x (batch, G, R, L) unit stride along L) and
    the mirrored-pairs flag."""
    return dict(_dt_only(dt_w, dt_x), mirror_pairs=int(bool(mirror_pairs)))


def _dt_only(dt_w, dt_x):
    if dt_w is None:
        return dict(dt_w=None, dt_x=None, dt_w_d_stride=0, dt_x_batch_stride=0, dt_x_group_stride=0, dt_x_r_stride=0, dt_rank=0)
    if dt_w.dim() != 2 or dt_w.stride(1) != 1 or dt_x.dim() != 4 or dt_x.stride(3) != 1 or dt_x.shape[2] != dt_w.shape[1]:
        raise RuntimeError("selscan_b200: dt_w must be (dim, R) with unit stride along R and dt_x (batch, G, R, L) with unit stride along L")
    return dict(dt_w=_p(dt_w), dt_x=_p(dt_x), dt_w_d_stride=dt_w.stride(0), dt_x_batch_stride=dt_x.stride(0),
                dt_x_group_stride=dt_x.stride(1), dt_x_r_stride=dt_x.stride(2), dt_rank=dt_w.shape[1])


def dt_fusable(batch, dim, seqlen, dstate, ngroups, dt_rank):
    """True when calls of these sizes may pass dt_w / dt_x instead of a materialised delta (selscan_b200_dt_fusable)."""
    return bool(_lib.load().selscan_b200_dt_fusable(batch, dim, seqlen, dstate, ngroups, dt_rank))


def mirror_ok(batch, dim, seqlen, dstate, ngroups):
    """True when calls of these sizes may use mirror_pairs (selscan_b200_mirror_ok)."""
    return bool(_lib.load().selscan_b200_mirror_ok(batch, dim, seqlen, dstate, ngroups))


def launch_fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus, out, out_z=None, last_state=None, ckpt=None,
               workspace=None, dt_w=None, dt_x=None, mirror_pairs=False):
    """One selscan_b200_fwd call on the current stream.  All tensors fp32 CUDA; B, C 4-D (batch, G, N, L);
    u/delta/z/out unit-stride along seqlen; outputs preallocated by the caller (the library never allocates).
    dt_w / dt_x: fused dt_proj (delta may then be None; see dt_fusable).
    mirror_pairs: groups (2j, 2j+1) share the rows of `u` / `out` (which then have dim / 2 rows); the odd group walks them back to
    front and both ADD their outputs into `out`, which the caller zero-initialises (see include/selscan_b200.h, mirror_ok)."""
    lib = _lib.load()
    _require_f32_cuda(u=u, delta=delta, A=A, B=B, C=C, D=D, z=z, delta_bias=delta_bias, out=out, out_z=out_z,
                      last_state=last_state, ckpt=ckpt, workspace=workspace, dt_w=dt_w, dt_x=dt_x)
    batch, dim, seqlen = u.shape
    if mirror_pairs:
        dim *= 2                                   # `dim` counts the channels of all groups; u / out carry one row per pair
    if delta is None:
        delta = u if dt_w is not None else None   # strides only; the pointer is passed as NULL below
    a = _lib.FwdArgs(
        batch=batch, dim=dim, seqlen=seqlen, dstate=A.shape[1], ngroups=B.shape[1],
        delta_softplus=int(bool(delta_softplus)),
        u=_p(u), delta=None if dt_w is not None else _p(delta), A=_p(A), B=_p(B), C=_p(C), D=_p(D), z=_p(z), delta_bias=_p(delta_bias),
        u_batch_stride=u.stride(0), u_d_stride=u.stride(1),
        delta_batch_stride=delta.stride(0), delta_d_stride=delta.stride(1),
        A_d_stride=A.stride(0), A_n_stride=A.stride(1),
        B_batch_stride=B.stride(0), B_group_stride=B.stride(1), B_n_stride=B.stride(2), B_l_stride=B.stride(3),
        C_batch_stride=C.stride(0), C_group_stride=C.stride(1), C_n_stride=C.stride(2), C_l_stride=C.stride(3),
        z_batch_stride=_strides2(z)[0], z_d_stride=_strides2(z)[1],
        out=_p(out), out_batch_stride=out.stride(0), out_d_stride=out.stride(1),
        out_z=_p(out_z), out_z_batch_stride=_strides2(out_z)[0], out_z_d_stride=_strides2(out_z)[1],
        last_state=_p(last_state), ckpt=_p(ckpt), workspace=_p(workspace), **_dt_fields(dt_w, dt_x, mirror_pairs))
    with torch.cuda.device(u.device):
        _lib.check(lib.selscan_b200_fwd(a, torch.cuda.current_stream(u.device).cuda_stream), "selscan_b200_fwd")


def launch_bwd(u, delta, A, B, C, D, z, delta_bias, dout, out, ckpt, delta_softplus,
               du, ddelta, dA, dB, dC, dD=None, dz=None, ddelta_bias=None, dt_w=None, dt_x=None, mirror_pairs=False):
    """One selscan_b200_bwd call on the current stream.  dA, dB, dC, dD, ddelta_bias must be zero-initialised.
    dt_w / dt_x: fused dt_proj, as in launch_fwd; ddelta is then the gradient w.r.t. the raw step dt_w . dt_x.
    mirror_pairs: as in launch_fwd; u / dout / du have one row per pair and du must be zero-initialised too."""
    lib = _lib.load()
    _require_f32_cuda(u=u, delta=delta, A=A, B=B, C=C, D=D, z=z, delta_bias=delta_bias, dout=dout, out=out, ckpt=ckpt, du=du,
                      ddelta=ddelta, dA=dA, dB=dB, dC=dC, dD=dD, dz=dz, ddelta_bias=ddelta_bias, dt_w=dt_w, dt_x=dt_x)
    batch, dim, seqlen = u.shape
    if mirror_pairs:
        dim *= 2
    if delta is None:
        delta = u if dt_w is not None else None
    a = _lib.BwdArgs(
        batch=batch, dim=dim, seqlen=seqlen, dstate=A.shape[1], ngroups=B.shape[1],
        delta_softplus=int(bool(delta_softplus)),
        u=_p(u), delta=None if dt_w is not None else _p(delta), A=_p(A), B=_p(B), C=_p(C), D=_p(D), z=_p(z), delta_bias=_p(delta_bias),
        dout=_p(dout), out=_p(out), ckpt=_p(ckpt),
        u_batch_stride=u.stride(0), u_d_stride=u.stride(1),
        delta_batch_stride=delta.stride(0), delta_d_stride=delta.stride(1),
        A_d_stride=A.stride(0), A_n_stride=A.stride(1),
        B_batch_stride=B.stride(0), B_group_stride=B.stride(1), B_n_stride=B.stride(2), B_l_stride=B.stride(3),
        C_batch_stride=C.stride(0), C_group_stride=C.stride(1), C_n_stride=C.stride(2), C_l_stride=C.stride(3),
        z_batch_stride=_strides2(z)[0], z_d_stride=_strides2(z)[1],
        dout_batch_stride=dout.stride(0), dout_d_stride=dout.stride(1),
        out_batch_stride=_strides2(out)[0], out_d_stride=_strides2(out)[1],
        du_batch_stride=du.stride(0), du_d_stride=du.stride(1),
        ddelta_batch_stride=ddelta.stride(0), ddelta_d_stride=ddelta.stride(1),
        dz_batch_stride=_strides2(dz)[0], dz_d_stride=_strides2(dz)[1],
        du=_p(du), ddelta=_p(ddelta), dz=_p(dz), dA=_p(dA), dB=_p(dB), dC=_p(dC), dD=_p(dD),
        ddelta_bias=_p(ddelta_bias), **_dt_fields(dt_w, dt_x, mirror_pairs))
    with torch.cuda.device(u.device):
        _lib.check(lib.selscan_b200_bwd(a, torch.cuda.current_stream(u.device).cuda_stream), "selscan_b200_bwd")


def bwd_kernel_name():
    """Kernel the tiled backward path launches on this device ("selscan_bwd_tm_kernel" unless overridden / unusable)."""
    return _lib.load().selscan_b200_bwd_kernel().decode()


def ckpt_elems(batch, dim, seqlen, dstate):
    return int(_lib.load().selscan_b200_ckpt_elems(batch, dim, seqlen, dstate))


def fwd_workspace_elems(batch, dim, seqlen, dstate, ngroups):
    """Floats of scratch that let a small-batch forward split the sequence into concurrent segments (0: not useful)."""
    return int(_lib.load().selscan_b200_fwd_workspace_elems(batch, dim, seqlen, dstate, ngroups))


class SelectiveScanFn(torch.autograd.Function):
    """selective_scan_interface.py:14-74, on the sm_100a kernels."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                return_last_state=False):
        batch, dim, seqlen, dstate = _check_inputs(u, delta, A, B, C, D, z, delta_bias)
        _lib.load()
        in_dtype = u.dtype
        if in_dtype != torch.float32:
            # fp16 / bf16 I/O: the kernels compute in fp32 exactly as the reference's do
            # (input_t -> float on load, selective_scan_common.h:148-176); widen here, round once on return.
            u, delta, B, C = u.float(), delta.float(), B.float(), C.float()
            z = z.float() if z is not None else None
        u, delta = _rowmajor(u), _rowmajor(delta)
        z = _rowmajor(z) if z is not None else None
        if D is not None:
            D = D.contiguous()
        if delta_bias is not None:
            delta_bias = delta_bias.contiguous()
        ctx.squeeze_B = B.dim() == 3
        ctx.squeeze_C = C.dim() == 3
        if ctx.squeeze_B:
            B = B.unsqueeze(1)  # selective_scan_interface.py:31-36
        if ctx.squeeze_C:
            C = C.unsqueeze(1)
        # constant (dim, dstate) B / C (selective_scan.cpp:238-246): a zero-stride broadcast view with one group per channel --
        # the kernels take any strides.  Mamba-UNet never passes them; this keeps the reference's API surface.
        ctx.const_B = B.dim() == 2
        ctx.const_C = C.dim() == 2
        if ctx.const_B:
            B = B.float().contiguous().view(1, dim, dstate, 1).expand(batch, dim, dstate, seqlen)
        if ctx.const_C:
            C = C.float().contiguous().view(1, dim, dstate, 1).expand(batch, dim, dstate, seqlen)
        ctx.rep_B = ctx.rep_C = 1
        if B.shape[1] != C.shape[1]:   # different groupings (e.g. constant B with variable C): use the finer one for both
            g = max(B.shape[1], C.shape[1])
            if g % B.shape[1] or g % C.shape[1]:
                raise RuntimeError(f"selective_scan_fn: B has {B.shape[1]} groups and C has {C.shape[1]}: incompatible")
            ctx.rep_B, ctx.rep_C = g // B.shape[1], g // C.shape[1]
            if ctx.rep_B > 1:
                B = B.repeat_interleave(ctx.rep_B, dim=1)
            if ctx.rep_C > 1:
                C = C.repeat_interleave(ctx.rep_C, dim=1)
        ngroups = B.shape[1]
        if C.shape[1] != ngroups or dim % ngroups != 0:
            raise RuntimeError(f"selective_scan_fn: B and C must share a group count dividing dim={dim}; "
                               f"got {B.shape[1]} and {C.shape[1]}")
        needs_grad = any(ctx.needs_input_grad)
        with torch.cuda.device(u.device):
            out = empty_rows(batch, dim, seqlen, u.device)
            out_z = empty_rows(batch, dim, seqlen, u.device) if z is not None else None
            last_state = (torch.empty((batch, dim, dstate), device=u.device, dtype=torch.float32)
                          if return_last_state else None)
            ckpt = None
            if needs_grad:  # saved scan states, the role of the reference's `x` (selective_scan.cpp:313)
                ckpt = torch.empty((max(ckpt_elems(batch, dim, seqlen, dstate), 4),), device=u.device,
                                   dtype=torch.float32)
            n_ws = fwd_workspace_elems(batch, dim, seqlen, dstate, ngroups) if z is None else 0
            ws = torch.empty((n_ws,), device=u.device, dtype=torch.float32) if n_ws > 0 else None
            launch_fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus, out, out_z, last_state, ckpt, ws)
        ctx.delta_softplus = bool(delta_softplus)
        ctx.has_z = z is not None
        ctx.in_dtype = in_dtype
        if needs_grad:
            if not ctx.has_z:
                ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, ckpt)
            else:
                ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, ckpt, out)
        ret = out_z if ctx.has_z else out
        if in_dtype != torch.float32:
            ret = ret.to(in_dtype)
        if not return_last_state:
            return ret
        ctx.mark_non_differentiable(last_state)
        return ret, last_state

    @staticmethod
    def backward(ctx, dout, *args):
        if not ctx.has_z:
            u, delta, A, B, C, D, delta_bias, ckpt = ctx.saved_tensors
            z = out = None
        else:
            u, delta, A, B, C, D, z, delta_bias, ckpt, out = ctx.saved_tensors
        dout = _rowmajor(dout.float())
        batch, dim, seqlen = u.shape
        dstate, ngroups = A.shape[1], B.shape[1]
        with torch.cuda.device(u.device):
            du = empty_rows(batch, dim, seqlen, u.device)
            ddelta = empty_rows(batch, dim, seqlen, u.device)
            dz = empty_rows(batch, dim, seqlen, u.device) if z is not None else None
            # reductions are accumulated in fp32 with atomics: zero-initialised, as selective_scan.cpp:458-466 -- one memset
            nbc = batch * ngroups * dstate * seqlen
            flat = torch.zeros((2 * nbc + dim * dstate + 2 * dim,), device=u.device, dtype=torch.float32)
            dB = flat[:nbc].view(batch, ngroups, dstate, seqlen)
            dC = flat[nbc:2 * nbc].view(batch, ngroups, dstate, seqlen)
            dA = flat[2 * nbc:2 * nbc + dim * dstate].view(dim, dstate)
            dD = flat[2 * nbc + dim * dstate:2 * nbc + dim * dstate + dim] if D is not None else None
            dbias = flat[2 * nbc + dim * dstate + dim:] if delta_bias is not None else None
            launch_bwd(u, delta, A, B, C, D, z, delta_bias, dout, out, ckpt, ctx.delta_softplus,
                       du, ddelta, dA, dB, dC, dD, dz, dbias)
        if ctx.rep_B > 1:   # gradients of a coarser grouping: sum over the channels that shared a group
            dB = dB.view(batch, ngroups // ctx.rep_B, ctx.rep_B, dstate, seqlen).sum(2)
        if ctx.rep_C > 1:
            dC = dC.view(batch, ngroups // ctx.rep_C, ctx.rep_C, dstate, seqlen).sum(2)
        dB = dB.sum(dim=(0, 3)) if ctx.const_B else (dB.squeeze(1) if ctx.squeeze_B else dB)  # selective_scan_interface.py:67-68
        dC = dC.sum(dim=(0, 3)) if ctx.const_C else (dC.squeeze(1) if ctx.squeeze_C else dC)
        if ctx.in_dtype != torch.float32:
            du, ddelta = du.to(ctx.in_dtype), ddelta.to(ctx.in_dtype)
            dB, dC = dB.to(ctx.in_dtype), dC.to(ctx.in_dtype)
            dz = dz.to(ctx.in_dtype) if dz is not None else None
        return (du, ddelta, dA, dB, dC, dD, dz, dbias, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """if return_last_state is True, returns (out, last_state); last_state is (batch, dim, dstate) and
    carries no gradient (selective_scan_interface.py:77-83)."""
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                       return_last_state=False):
    """Plain-PyTorch statement of the op (any device), the role of selective_scan_interface.py:86-152.

    u, delta: (B, D, L); A: (D, N) real; B, C: (B, N, L) or (B, G, N, L); D, delta_bias: (D,); z: (B, D, L).
    Exported because callers import it next to selective_scan_fn (code/networks/mamba_sys.py:17-20);
    selective_scan_fn never calls it."""
    dtype_in = u.dtype
    u, delta = u.float(), delta.float()
    if delta_bias is not None:
        delta = delta + delta_bias[..., None].float()
    if delta_softplus:
        delta = F.softplus(delta)
    if A.is_complex():
        raise RuntimeError("selective_scan_ref (B200 build): complex A is not provided")
    batch, dim, seqlen = u.shape
    dstate = A.shape[1]
    B, C = B.float(), C.float()
    if B.dim() == 2:   # constant (dim, dstate)
        B = B.view(1, dim, dstate, 1).expand(batch, dim, dstate, seqlen)
    if C.dim() == 2:
        C = C.view(1, dim, dstate, 1).expand(batch, dim, dstate, seqlen)
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    B = B.repeat_interleave(dim // B.shape[1], dim=1)  # channel d -> group d // (dim / G)
    C = C.repeat_interleave(dim // C.shape[1], dim=1)
    decay = torch.exp(delta.unsqueeze(-1) * A.view(1, dim, 1, dstate))           # (B, D, L, N)
    drive = (delta * u).unsqueeze(-1) * B.permute(0, 1, 3, 2)                    # (B, D, L, N)
    x = u.new_zeros((batch, dim, dstate))
    ys = []
    for l in range(seqlen):
        x = decay[:, :, l] * x + drive[:, :, l]
        ys.append((x * C[:, :, :, l]).sum(-1))
    y = torch.stack(ys, dim=2) if ys else u.new_zeros((batch, dim, 0))
    out = y if D is None else y + u * D.view(1, dim, 1).float()
    if z is not None:
        out = out * F.silu(z.float())
    out = out.to(dtype_in)
    return out if not return_last_state else (out, x)
