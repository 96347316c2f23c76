"""Short-row LayerNorm kernels (caller side of SS2D: ln_1 and the patch merging / expanding norms) against torch's
F.layer_norm, forward and backward, through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,D", [(1, 16), (37, 96), (24 * 56 * 56, 96), (1000, 192), (333, 200), (4704, 384), (77, 768),
                                    (1176, 1536), (5, 33), (64, 1), (9, 1600)])
def test_layernorm_matches_torch(rows, D):
    from selscan_b200.layernorm import layer_norm

    torch.manual_seed(rows + D)
    x = (torch.randn(rows, D, device="cuda") * 2 + 0.5).requires_grad_()
    w = (1 + 0.3 * torch.randn(D, device="cuda")).requires_grad_()
    b = (0.2 * torch.randn(D, device="cuda")).requires_grad_()
    g = torch.randn(rows, D, device="cuda")
    ref = F.layer_norm(x, (D,), w, b, 1e-5)
    gref = torch.autograd.grad(ref, (x, w, b), g)
    out = layer_norm(x, w, b, 1e-5)
    got = torch.autograd.grad(out, (x, w, b), g)
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(got[0], gref[0], rtol=1e-4, atol=1e-5)
    for a, r, name in ((got[1], gref[1], "dweight"), (got[2], gref[2], "dbias")):
        torch.testing.assert_close(a, r, rtol=1e-4, atol=1e-5 * max(1.0, float(r.abs().max())), msg=lambda m: f"{name}: {m}")


def test_layernorm_module_and_patch():
    from selscan_b200.layernorm import LayerNorm, patch_layernorms

    torch.manual_seed(0)
    ref = torch.nn.LayerNorm(96).cuda()
    mine = LayerNorm(96).cuda()
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(2, 14, 14, 96, device="cuda").permute(0, 2, 1, 3)      # non-contiguous input
    torch.testing.assert_close(mine(x), ref(x), rtol=1e-5, atol=1e-5)
    with torch.no_grad():
        torch.testing.assert_close(mine(x), ref(x), rtol=1e-5, atol=1e-5)
    seq = torch.nn.Sequential(torch.nn.Linear(8, 96), torch.nn.LayerNorm(96)).cuda()
    want = seq(torch.ones(3, 8, device="cuda"))
    assert patch_layernorms(seq) == 1 and type(seq[1]) is LayerNorm
    torch.testing.assert_close(seq(torch.ones(3, 8, device="cuda")), want, rtol=1e-5, atol=1e-5)
