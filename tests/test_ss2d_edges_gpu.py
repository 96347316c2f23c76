"""The SS2D edge kernels against the plain-torch chains they replace
(/root/reference/code/networks/mamba_sys.py:533-534 + :403-404 and :429-434 + :536), forward and backward, through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(2, 8, 7, 7), (1, 5, 4, 9), (2, 192, 56, 56), (3, 64, 14, 14), (2, 40, 28, 28), (1, 384, 7, 7), (2, 33, 9, 5),
          (1, 1536, 7, 7), (1, 768, 14, 14)]


def _fp32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _close(got, ref, rtol, atol_rel, name):
    scale = max(1.0, float(ref.abs().max()))
    torch.testing.assert_close(got, ref, rtol=rtol, atol=atol_rel * scale, msg=lambda m: f"{name}: {m}")


@pytest.mark.parametrize("B,D,H,W", SHAPES)
def test_prologue_conv_silu_cross_scan(B, D, H, W):
    from selscan_b200.ss2d import cross_scan_torch, edge_in_bwd, edge_in_fwd

    _fp32()
    torch.manual_seed(B * 1000 + D + H)
    xz = torch.randn(B, H, W, 2 * D, device="cuda")
    w = (0.4 * torch.randn(D, 1, 3, 3, device="cuda")).requires_grad_()
    b = (0.2 * torch.randn(D, device="cuda")).requires_grad_()
    xh = xz[..., :D].detach().clone().requires_grad_()
    ref = cross_scan_torch(F.silu(F.conv2d(xh.permute(0, 3, 1, 2).contiguous(), w, b, padding=1, groups=D)))
    xs = edge_in_fwd(xz, D, w.detach(), b.detach())
    assert xs.shape == ref.shape
    _close(xs, ref.detach(), 1e-5, 1e-6, "xs")
    g = torch.randn_like(ref)
    gx_ref, gw_ref, gb_ref = torch.autograd.grad(ref, (xh, w, b), g)
    L = H * W
    pitch = (L + 3) // 4 * 4
    gp = torch.empty(B, 4, D, pitch, device="cuda")[..., :L]
    gp.copy_(g)
    d_xz = torch.full_like(xz, 7.0)
    gw, gb = edge_in_bwd(gp, xz, D, w.detach(), b.detach(), d_xz)
    assert torch.equal(d_xz[..., D:], torch.full_like(d_xz[..., D:], 7.0))      # the z half belongs to the other edge
    _close(d_xz[..., :D], gx_ref, 1e-4, 1e-5, "dx")
    _close(gw, gw_ref, 1e-3, 1e-4, "dconv_w")
    _close(gb, gb_ref, 1e-3, 1e-4, "dconv_b")


@pytest.mark.parametrize("gated", [True, False])
@pytest.mark.parametrize("B,D,H,W", SHAPES)
def test_epilogue_merge_layernorm_gate(B, D, H, W, gated):
    from selscan_b200.ss2d import cross_merge_torch, edge_out_bwd, edge_out_fwd

    _fp32()
    torch.manual_seed(B * 1000 + D + W)
    L = H * W
    pitch = (L + 3) // 4 * 4
    ys = torch.empty(B, 4, D, pitch, device="cuda")[..., :L]
    ys.copy_(torch.randn(B, 4, D, L, device="cuda") + 0.3)
    xz = torch.randn(B, H, W, 2 * D, device="cuda")
    gamma = (1.0 + 0.3 * torch.randn(D, device="cuda")).requires_grad_()
    beta = (0.2 * torch.randn(D, device="cuda")).requires_grad_()
    ys_r = ys.detach().clone().requires_grad_()
    z_r = xz[..., D:].detach().clone().requires_grad_()
    y = cross_merge_torch(ys_r, H, W).transpose(1, 2).contiguous().view(B, H, W, D)
    ref = F.layer_norm(y, (D,), gamma, beta, 1e-5)
    if gated:
        ref = ref * F.silu(z_r)
    esz = xz.element_size()
    zptr = xz.data_ptr() + D * esz if gated else None
    out, xhat, rstd = edge_out_fwd(ys, H, W, zptr, 2 * D, gamma.detach(), beta.detach(), 1e-5, True)
    _close(out, ref.detach(), 1e-4, 1e-5, "out")
    out2, xh2, rs2 = edge_out_fwd(ys, H, W, zptr, 2 * D, gamma.detach(), beta.detach(), 1e-5, False)   # inference: nothing saved
    assert xh2 is None and rs2 is None and torch.equal(out, out2)
    g = torch.randn_like(ref)
    grads = torch.autograd.grad(ref, (ys_r, gamma, beta) + ((z_r,) if gated else ()), g)
    d_xz = torch.full_like(xz, 7.0)
    d_ys, d_gamma, d_beta = edge_out_bwd(g, H, W, zptr, 2 * D, xhat, rstd, gamma.detach(), beta.detach(),
                                         d_xz.data_ptr() + D * esz if gated else None, 2 * D)
    assert torch.equal(d_xz[..., :D], torch.full_like(d_xz[..., :D], 7.0))
    _close(d_ys, grads[0], 1e-3, 1e-5, "dys")
    _close(d_gamma, grads[1], 1e-3, 1e-4, "dgamma")
    _close(d_beta, grads[2], 1e-3, 1e-4, "dbeta")
    if gated:
        _close(d_xz[..., D:], grads[3], 1e-3, 1e-5, "dz")


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("B,d_model,H,W", [(2, 16, 8, 8), (1, 20, 7, 7), (2, 96, 56, 56), (2, 24, 5, 9), (3, 192, 14, 14)])
def test_fused_ss2d_block_matches_separate_kernels(B, d_model, H, W, tc, monkeypatch):
    """SS2D.forward through SS2DFusedFn (hand-written backward) == the same block through conv2d / CrossScan / autograd;
    tc: x_proj / dt_proj and their gradients through the 3xTF32 tensor-core GEMM instead of cuBLAS fp32."""
    from selscan_b200 import ss2d
    from selscan_b200.vssm import SS2D

    monkeypatch.setattr(ss2d, "TC_PROJ", tc)
    _fp32()
    torch.manual_seed(d_model + H)
    blk = SS2D(d_model).cuda()
    with torch.no_grad():
        blk.out_norm.weight.add_(0.2 * torch.randn_like(blk.out_norm.weight))
        blk.out_norm.bias.add_(0.2 * torch.randn_like(blk.out_norm.bias))
        blk.Ds.add_(0.3 * torch.randn_like(blk.Ds))
    x = torch.randn(B, H, W, d_model, device="cuda")
    g = torch.randn(B, H, W, d_model, device="cuda")
    res = []
    for fused in (True, False):
        blk.fused = fused
        blk.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_()
        out = blk(xi)
        out.backward(g)
        res.append((out.detach(), xi.grad, {k: p.grad.clone() for k, p in blk.named_parameters()}))
    _close(res[0][0], res[1][0], 1e-4, 1e-5, "out")
    _close(res[0][1], res[1][1], 1e-3, 1e-4, "dx")
    assert set(res[0][2]) == {k for k, _ in blk.named_parameters()}
    for k in res[1][2]:
        _close(res[0][2][k], res[1][2][k], 2e-3, 2e-4, k)
    with torch.no_grad():
        blk.fused = True
        _close(blk(x), res[1][0], 1e-4, 1e-5, "inference out")


def test_large_plane_takes_the_separate_kernels():
    """A plane too large for the edge kernels' shared memory (80 x 80) still runs on the GPU through CrossScan / CrossMerge and
    torch's conv / LayerNorm; same result as the reference chain of a smaller-plane-capable block."""
    from selscan_b200 import ss2d
    from selscan_b200.vssm import SS2D

    _fp32()
    torch.manual_seed(3)
    blk = SS2D(16).cuda()
    assert not ss2d.fused_supported(blk, 80, 80) and ss2d.fused_supported(blk, 56, 56)
    x = torch.randn(1, 80, 80, 16, device="cuda", requires_grad=True)
    out = blk(x)
    blk.fused = False
    ref = blk(x)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-6)
    out.sum().backward()
    assert torch.isfinite(x.grad).all()


def test_patch_ss2d_installs_forward():
    from selscan_b200.ss2d import forward_b200, forward_core_b200, patch_ss2d

    class Dummy:
        pass

    patch_ss2d(Dummy)
    assert Dummy.forward is forward_b200 and Dummy.forward_corev0 is forward_core_b200


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("fused", [True, False])
def test_ss2d_block_under_autocast(dtype, fused):
    """The reference runs SS2D under autocast (mamba_sys.py:865) and forward_corev0 casts xs / dts / Bs / Cs to float for the scan
    (:411-427).  Here in_proj / out_proj may run in half precision, everything between them (projections, scan, LayerNorm) must
    stay fp32: outputs and gradients of the autocast run agree with the fp32 run to half-precision round-off, and nothing reads a
    half buffer as fp32."""
    from selscan_b200.vssm import SS2D

    _fp32()
    torch.manual_seed(11)
    blk = SS2D(32).cuda()
    blk.fused = fused
    x = torch.randn(2, 14, 14, 32, device="cuda")
    g = torch.randn(2, 14, 14, 32, device="cuda")
    res = []
    for ac in (False, True):
        blk.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_()
        with torch.autocast("cuda", dtype=dtype, enabled=ac):
            out = blk(xi)
        out.float().backward(g)
        res.append((out.detach().float(), xi.grad.float(), {k: p.grad.float().clone() for k, p in blk.named_parameters()}))
    tol = 4e-2 if dtype == torch.bfloat16 else 8e-3
    for a, b, nm in [(res[1][0], res[0][0], "out"), (res[1][1], res[0][1], "dx")] + [(res[1][2][k], res[0][2][k], k) for k in res[0][2]]:
        assert torch.isfinite(a).all(), nm
        scale = float(b.abs().max())
        assert float((a - b).abs().max()) <= tol * max(scale, 1e-3), (nm, float((a - b).abs().max()), scale)


def test_cross_scan_merge_planes_beyond_shared_memory():
    """Token planes above ~225 x 225 do not fit the plane kernels' shared memory: CrossScan / CrossMerge (and with them
    forward_core_b200) take the plain-torch chains instead of raising."""
    from selscan_b200.ss2d import CrossMerge, CrossScan, cross_merge_torch, cross_scan_torch

    torch.manual_seed(5)
    x = torch.randn(1, 3, 230, 231, device="cuda", requires_grad=True)
    xs = CrossScan.apply(x)
    assert torch.equal(xs, cross_scan_torch(x.detach()))
    y = CrossMerge.apply(xs, 230, 231)
    torch.testing.assert_close(y, cross_merge_torch(xs.detach(), 230, 231))
    y.sum().backward()
    torch.testing.assert_close(x.grad, torch.full_like(x, 4.0))


@pytest.mark.parametrize("batch,D,L,R", [(10, 192, 3136, 6), (4, 384, 784, 12), (6, 64, 200, 5), (5, 128, 96, 9)])
def test_fused_dt_proj_in_the_scan_kernels(batch, D, L, R):
    """dt_proj inside the scan (mamba_sys.py:409 fused into the kernels): the C ABI called with dt_w / dt_x and
    no delta tensor gives the same out, saved states and gradients as with delta = dt_w . dt_x materialised by a GEMM."""
    from selscan_b200 import ops

    torch.manual_seed(R * 100 + L)
    K, N = 4, 16
    KD, C = K * D, R + 2 * N
    dev = "cuda"
    assert ops.dt_fusable(batch, KD, L, N, K, R)
    x_dbl = torch.randn(batch, K, C, L, device=dev)
    dt_w = torch.randn(K, D, R, device=dev) * R ** -0.5
    u, dout = torch.randn(batch, KD, L, device=dev), torch.randn(batch, KD, L, device=dev)
    A = -torch.rand(KD, N, device=dev) * 4 - 0.1
    Dp, bias = torch.randn(KD, device=dev), torch.randn(KD, device=dev) - 2
    Bv, Cv, dt_x = x_dbl[:, :, R:R + N], x_dbl[:, :, R + N:], x_dbl[:, :, :R]
    delta = torch.matmul(dt_w.unsqueeze(0), dt_x).view(batch, KD, L)
    res = []
    for fused in (False, True):
        out = torch.empty(batch, KD, L, device=dev)
        ck = torch.empty(max(ops.ckpt_elems(batch, KD, L, N), 4), device=dev)
        du, dd = torch.empty_like(out), torch.empty_like(out)
        nbc = batch * K * N * L
        flat = torch.zeros(2 * nbc + KD * N + 2 * KD, device=dev)
        dB, dC = flat[:nbc].view(batch, K, N, L), flat[nbc:2 * nbc].view(batch, K, N, L)
        dA = flat[2 * nbc:2 * nbc + KD * N].view(KD, N)
        dD, db = flat[2 * nbc + KD * N:2 * nbc + KD * N + KD], flat[2 * nbc + KD * N + KD:]
        kw = dict(dt_w=dt_w.view(KD, R), dt_x=dt_x) if fused else {}
        ops.launch_fwd(u, None if fused else delta, A, Bv, Cv, Dp, None, bias, True, out, None, None, ck, None, **kw)
        ops.launch_bwd(u, None if fused else delta, A, Bv, Cv, Dp, None, bias, dout, None, ck, True, du, dd, dA, dB, dC, dD, None, db, **kw)
        torch.cuda.synchronize()
        res.append(dict(out=out, ck=ck, du=du, ddelta=dd, flat=flat))
    for k in res[0]:
        a, b = res[1][k], res[0][k]
        scale = max(1.0, float(b.abs().max()))
        torch.testing.assert_close(a, b, rtol=1e-4 if k in ("out", "ck") else 1e-3, atol=(1e-5 if k in ("out", "ck") else 1e-4) * scale,
                                   msg=lambda m: f"{k}: {m}")
    assert not ops.dt_fusable(batch, KD, L, N, K, 13) and not ops.dt_fusable(batch, KD, L + 1, N, K, R)


@pytest.mark.parametrize("B,d_model,H,W", [(10, 96, 56, 56), (4, 192, 28, 28)])
def test_ss2d_block_with_and_without_fused_dt_proj(B, d_model, H, W, monkeypatch):
    """The SS2D block (SS2DFusedFn) allocates no (B, 4D, L) step tensor at stages 1 and 2 and computes the same thing."""
    from selscan_b200 import ss2d
    from selscan_b200.vssm import SS2D

    _fp32()
    torch.manual_seed(d_model)
    blk = SS2D(d_model).cuda()
    x = torch.randn(B, H, W, d_model, device="cuda")
    g = torch.randn(B, H, W, d_model, device="cuda")
    res = []
    monkeypatch.setattr(ss2d, "FUSE_DT_MAX_RANK", 12)
    for fuse in (True, False):
        monkeypatch.setattr(ss2d, "FUSE_DT", fuse)
        blk.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_()
        out = blk(xi)
        out.backward(g)
        res.append((out.detach(), xi.grad, {k: p.grad.clone() for k, p in blk.named_parameters()}))
    _close(res[0][0], res[1][0], 1e-4, 1e-5, "out")
    _close(res[0][1], res[1][1], 1e-3, 1e-4, "dx")
    for k in res[1][2]:
        _close(res[0][2][k], res[1][2][k], 2e-3, 2e-4, k)


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("B,d_model,H,W", [(10, 96, 56, 56), (4, 192, 28, 28), (3, 384, 14, 14), (24, 32, 12, 20)])
def test_ss2d_block_with_and_without_mirrored_scan_directions(B, d_model, H, W, tc, monkeypatch):
    """MIRROR: the prologue / epilogue move two scan-order planes instead of four and the scan kernels walk the reversed directions
    themselves (mamba_sys.py:404, :429 inside the kernels).  Same outputs, same gradients for every parameter (which come back in the
    reference's direction order)."""
    from selscan_b200 import ops, ss2d
    from selscan_b200.vssm import SS2D

    _fp32()
    monkeypatch.setattr(ss2d, "TC_PROJ", tc)
    torch.manual_seed(d_model + W)
    blk = SS2D(d_model).cuda()
    with torch.no_grad():   # make the four directions differ in every parameter
        for p in (blk.A_logs, blk.Ds, blk.dt_projs_bias, blk.out_norm.weight, blk.out_norm.bias):
            p.add_(0.2 * torch.randn_like(p))
    assert ops.mirror_ok(B, 8 * d_model, H * W, 16, 4)
    x = torch.randn(B, H, W, d_model, device="cuda")
    g = torch.randn(B, H, W, d_model, device="cuda")
    res = []
    for mir in (True, False):      # (the default "auto" picks one of the two per block)
        monkeypatch.setattr(ss2d, "MIRROR", mir)
        blk.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_()
        out = blk(xi)
        out.backward(g)
        res.append((out.detach(), xi.grad, {k: p.grad.clone() for k, p in blk.named_parameters()}))
    _close(res[0][0], res[1][0], 1e-4, 1e-5, "out")
    _close(res[0][1], res[1][1], 1e-3, 1e-4, "dx")
    for k in res[1][2]:
        _close(res[0][2][k], res[1][2][k], 2e-3, 2e-4, k)
