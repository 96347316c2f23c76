"""Mirrored direction pairs inside the scan kernels (include/selscan_b200.h: mirror_pairs; the reversed half of SS2D's CrossScan /
CrossMerge, code/networks/mamba_sys.py:404 `torch.flip(x_hwwh, dims=[-1])` and :429 `torch.flip(out_y[:, 2:4], dims=[-1])`):
groups (2j, 2j+1) share one set of u rows, the odd group walks them back to front, outputs and du of a pair are summed in source
order.  Checked against the SAME kernels run the plain way on explicitly flipped copies (which are checked against the oracle
elsewhere), forward and every gradient, with and without dt_proj fused."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(ops, batch, P, D, L, R, fused, mirror, t):
    """One forward + backward through the C ABI.  mirror: u / dout given per pair, delta / dt_x / B / C per group in source order.
    Otherwise everything per group in scan order.  Returns out, du, ddelta, dB, dC, dA, dD, dbias."""
    dev, N, G = "cuda", 16, 2 * P
    KD = G * D
    rows = P * D if mirror else KD
    out = torch.zeros(batch, rows, L, device=dev)
    du = torch.zeros(batch, rows, L, device=dev)
    dd = torch.empty(batch, KD, L, device=dev)
    ck = torch.empty(max(ops.ckpt_elems(batch, KD, L, N), 4), device=dev)
    nbc = batch * G * N * L
    flat = torch.zeros(2 * nbc + KD * N + 2 * KD, device=dev)
    dB, dC = flat[:nbc].view(batch, G, N, L), flat[nbc:2 * nbc].view(batch, G, N, L)
    dA = flat[2 * nbc:2 * nbc + KD * N].view(KD, N)
    dD, db = flat[2 * nbc + KD * N:2 * nbc + KD * N + KD], flat[2 * nbc + KD * N + KD:]
    kw = dict(mirror_pairs=mirror)
    delta = None
    if fused:
        kw.update(dt_w=t["dt_w"].view(KD, R), dt_x=t["dt_x"])
    else:
        delta = torch.matmul(t["dt_w"].unsqueeze(0), t["dt_x"]).view(batch, KD, L)
    ops.launch_fwd(t["u"], delta, t["A"], t["B"], t["C"], t["D"], None, t["bias"], True, out, None, None, ck, None, **kw)
    ops.launch_bwd(t["u"], delta, t["A"], t["B"], t["C"], t["D"], None, t["bias"], t["dout"], None, ck, True, du, dd, dA, dB, dC, dD,
                   None, db, **kw)
    torch.cuda.synchronize()
    return dict(out=out, du=du, ddelta=dd.view(batch, G, D, L), dB=dB.clone(), dC=dC.clone(), dA=dA.clone(), dD=dD.clone(), dbias=db.clone())


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("batch,P,D,L,R", [(10, 2, 192, 3136, 6), (8, 2, 384, 784, 12), (6, 2, 128, 100, 5), (5, 1, 64, 76, 3),
                                             (24, 2, 768, 196, 12), (7, 2, 64, 12, 4), (9, 2, 64, 36, 6), (3, 3, 128, 260, 2)])
def test_mirror_pairs_match_explicit_flips(batch, P, D, L, R, fused):
    from selscan_b200 import ops

    N, G = 16, 2 * P
    KD = G * D
    assert ops.mirror_ok(batch, KD, L, N, G)
    assert not ops.mirror_ok(batch, KD, L + 1, N, G) and not ops.mirror_ok(batch, KD // G * 3, L, N, 3)   # seqlen % 4, odd group count
    if fused and (L % 4 or not ops.dt_fusable(batch, KD, L, N, G, R)):
        pytest.skip("dt_proj not fusable at these sizes")
    torch.manual_seed(L + R)
    dev = "cuda"
    Cw = R + 2 * N
    x_dbl = torch.randn(batch, G, Cw, L, device=dev)                    # per group, SOURCE order
    src = dict(u=torch.randn(batch, P * D, L, device=dev), dout=torch.randn(batch, P * D, L, device=dev),
               dt_w=torch.randn(G, D, R, device=dev) * R ** -0.5, dt_x=x_dbl[:, :, :R], B=x_dbl[:, :, R:R + N], C=x_dbl[:, :, R + N:],
               A=-torch.rand(KD, N, device=dev) * 4 - 0.1, D=torch.randn(KD, device=dev), bias=torch.randn(KD, device=dev) - 2)
    got = _run(ops, batch, P, D, L, R, fused, True, src)

    odd = torch.arange(G, device=dev) % 2 == 1

    def scan_order(x):      # (batch, G, ., L) source order -> scan order: odd groups flipped
        return torch.where(odd.view(1, G, 1, 1), x.flip(-1), x).contiguous()

    def per_group(x):       # (batch, P*D, L) per pair -> (batch, G, D, L), the odd copy flipped
        x = x.view(batch, P, 1, D, L).expand(batch, P, 2, D, L).reshape(batch, G, D, L)
        return scan_order(x)

    xd = scan_order(x_dbl)
    ref_in = dict(src, u=per_group(src["u"]).view(batch, KD, L), dout=per_group(src["dout"]).view(batch, KD, L),
                  dt_x=xd[:, :, :R], B=xd[:, :, R:R + N], C=xd[:, :, R + N:])
    ref = _run(ops, batch, P, D, L, R, False, False, ref_in)

    def fold(x):            # (batch, KD, L) scan order per group -> per pair, source order, summed
        x = scan_order(x.view(batch, G, D, L)).view(batch, P, 2, D, L)
        return (x[:, :, 0] + x[:, :, 1]).reshape(batch, P * D, L)

    want = dict(out=fold(ref["out"]), du=fold(ref["du"]), ddelta=scan_order(ref["ddelta"]), dB=scan_order(ref["dB"]),
                dC=scan_order(ref["dC"]), dA=ref["dA"], dD=ref["dD"], dbias=ref["dbias"])
    for k, w in want.items():
        scale = max(1.0, float(w.abs().max()))
        fwd = k == "out"
        torch.testing.assert_close(got[k], w, rtol=1e-4 if fwd else 1e-3, atol=(1e-5 if fwd else 1e-4) * scale, msg=lambda m: f"{k}: {m}")
