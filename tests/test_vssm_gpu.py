"""Caller context on the GPU: selscan_b200.vssm.VSSM (from-scratch model, SS2D core = forward_core_b200 on the sm_100a kernels) must
reproduce the forward and backward of the REFERENCE's own VSSM (golden fixture made by tests/golden/make_golden_model.py from the
unchanged reference model + reference selective_scan_ref on CPU), with the reference's state dict loaded strictly."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_vssm_matches_reference_model():
    from selscan_b200.vssm import VSSM

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = load_golden("model_vssm_small")
    model = VSSM(depths=(1, 1, 1, 1), dims=(16, 32, 64, 128), drop_path_rate=0.0)
    sd = {k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}
    model.load_state_dict(sd, strict=True)          # same parameter names and shapes as the reference
    model = model.cuda().eval()
    x = torch.from_numpy(g["x"]).cuda().requires_grad_()
    out = model(x)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out"], rtol=1e-3, atol=2e-4)
    (out * torch.from_numpy(g["dout"]).cuda()).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g["dx"], rtol=2e-3, atol=2e-4 * float(np.abs(g["dx"]).max()))
    params = dict(model.named_parameters())
    checked = 0
    for k, ref in g.items():
        if not k.startswith("grad."):
            continue
        got = params[k[5:]].grad.cpu().numpy()
        scale = max(1e-6, float(np.abs(ref).max()))
        np.testing.assert_allclose(got, ref, rtol=2e-3, atol=5e-4 * scale, err_msg=k)
        checked += 1
    assert checked >= 20


def test_mambaunet_train_step_runs_and_is_finite():
    """One supervised step of the benchmark workload shape (train_fully_supervised_2D_VIM.py:152-160) at batch 2."""
    from selscan_b200.vssm import DiceLoss, MambaUnet

    torch.manual_seed(0)
    model = MambaUnet(num_classes=4).cuda().train()
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)
    x = torch.rand(2, 1, 224, 224, device="cuda")
    y = torch.randint(0, 4, (2, 224, 224), device="cuda")
    out = model(x)
    assert out.shape == (2, 4, 224, 224)
    loss = 0.5 * (torch.nn.functional.cross_entropy(out, y) + DiceLoss(4)(torch.softmax(out, 1), y.unsqueeze(1)))
    opt.zero_grad()
    loss.backward()
    opt.step()
    assert torch.isfinite(loss)
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    assert sum(p.grad is not None for p in model.parameters()) == len(list(model.parameters()))


def test_training_step_captured_as_cuda_graph_matches_eager():
    """workloads.GraphedStep: every kernel of the step (TMA descriptors included) must be capturable, and the replayed
    step must produce the same losses as the eager step on an identical copy of the model."""
    import copy

    from selscan_b200 import workloads as wl
    from selscan_b200.vssm import DiceLoss, MambaUnet

    torch.manual_seed(0)
    m1 = MambaUnet(num_classes=4, depths=(1, 1, 1, 1), dims=(32, 64, 128, 256), drop_path_rate=0.0).cuda().train()
    m2 = copy.deepcopy(m1)
    o1, o2 = wl.make_sgd(m1), wl.make_sgd(m2)
    dice = DiceLoss(4)
    x = torch.rand(2, 1, 64, 64, device="cuda")
    y = torch.randint(0, 4, (2, 64, 64), device="cuda")
    g = wl.GraphedStep(lambda a, b: wl.supervised_step(m1, o1, dice, a, b), x, y, warmup=2)
    for _ in range(2):
        wl.supervised_step(m2, o2, dice, x, y)
    for _ in range(3):
        lg = float(g(x, y).detach())
        le = float(wl.supervised_step(m2, o2, dice, x, y).detach())
        assert abs(lg - le) <= 1e-5 * max(1.0, abs(le)), (lg, le)


def test_semi_step_on_two_streams_matches_one_stream():
    """workloads.semi_step(side=stream): the second network's forward (and, through autograd, its backward) runs on a second stream,
    concurrently with the first one's.  Same losses and parameters as the single-stream step (drop path off: its random masks
    are drawn in issue order, which differs), eager and captured as a CUDA graph."""
    import copy

    from selscan_b200 import workloads as wl
    from selscan_b200.vssm import DiceLoss, MambaUnet

    torch.manual_seed(3)
    kw = dict(num_classes=4, depths=(1, 1, 1, 1), dims=(32, 64, 128, 256), drop_path_rate=0.0)
    a0, b0 = MambaUnet(**kw).cuda().train(), MambaUnet(**kw).cuda().train()
    dice, cw = DiceLoss(4), wl.consistency_weight(3000)
    x = torch.rand(4, 1, 64, 64, device="cuda")
    y = torch.randint(0, 4, (4, 64, 64), device="cuda")
    runs = {}
    for mode in ("one", "two", "two_graph"):
        a, b = copy.deepcopy(a0), copy.deepcopy(b0)
        oa, ob = wl.make_sgd(a), wl.make_sgd(b)
        side = None if mode == "one" else torch.cuda.Stream()
        step = lambda xx, yy: wl.semi_step(a, b, oa, ob, dice, xx, yy, 2, cw, side)   # noqa: E731
        if mode == "two_graph":
            step = wl.GraphedStep(step, x, y, warmup=2)      # two eager steps first (lazy optimizer state), then the capture
            losses = [None, None] + [float(step(x, y).detach()) for _ in range(3)]
        else:
            losses = [float(step(x, y).detach()) for _ in range(5)]
        torch.cuda.synchronize()
        runs[mode] = (losses, [p.detach().clone() for p in list(a.parameters()) + list(b.parameters())])
    for mode in ("two", "two_graph"):
        for l1, l2 in zip(runs["one"][0], runs[mode][0]):
            assert l2 is None or abs(l1 - l2) <= 2e-5 * max(1.0, abs(l1)), (mode, runs["one"][0], runs[mode][0])
        for p1, p2 in zip(runs["one"][1], runs[mode][1]):
            torch.testing.assert_close(p2, p1, rtol=1e-3, atol=2e-5)


@pytest.mark.parametrize("B,D,H,W", [(2, 5, 7, 7), (1, 3, 4, 9), (2, 8, 56, 56), (3, 4, 14, 14)])
def test_cross_scan_merge_kernels(B, D, H, W):
    """The plane kernels against the plain-torch statement of CrossScan / CrossMerge, forward and backward (bit-exact:
    pure data movement for the scan, a fixed 4-term sum for the merge)."""
    from selscan_b200.ss2d import CrossMerge, CrossScan, cross_merge_torch, cross_scan_torch

    torch.manual_seed(0)
    x = torch.randn(B, D, H, W, device="cuda", requires_grad=True)
    xs = CrossScan.apply(x)
    ref = cross_scan_torch(x)
    assert torch.equal(xs, ref)
    g = torch.randn_like(ref)
    gx, = torch.autograd.grad(xs, x, g)
    gx_ref, = torch.autograd.grad(ref, x, g)
    torch.testing.assert_close(gx, gx_ref, rtol=1e-6, atol=1e-6)
    ys = torch.randn(B, 4, D, H * W, device="cuda", requires_grad=True)
    y = CrossMerge.apply(ys, H, W)
    y_ref = cross_merge_torch(ys, H, W)
    torch.testing.assert_close(y, y_ref, rtol=1e-6, atol=1e-6)
    gy = torch.randn_like(y_ref)
    g1, = torch.autograd.grad(y, ys, gy)
    g2, = torch.autograd.grad(y_ref, ys, gy)
    assert torch.equal(g1, g2)
