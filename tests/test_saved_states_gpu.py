"""The forward's saved scan states ("checkpoints": the state after every 8th position, include/selscan_b200.h `ckpt`) against a plain
fp64 recurrence -- the backward restarts from them, so the gradient tests cover them too; this one names the culprit directly.
Covers the tiled forward's staging tiles + TMA stores: full tiles, a partial last tile, and the interval a sequence ends in."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("b,kd,L", [(2, 256, 784), (3, 256, 100), (2, 512, 49), (1, 256, 3136)])
def test_saved_states_match_recurrence(b, kd, L):
    from selscan_b200 import ops

    torch.manual_seed(L)
    dev, K, N = "cuda", 4, 16
    u = torch.randn(b, kd, L, device=dev)
    delta = torch.randn(b, kd, L, device=dev) * 0.5
    A = -torch.rand(kd, N, device=dev) - 0.5
    B, C = torch.randn(b, K, N, L, device=dev), torch.randn(b, K, N, L, device=dev)
    D, bias = torch.ones(kd, device=dev), torch.full((kd,), -1.0, device=dev)
    out = torch.empty_like(u)
    n_ck = (L + 7) // 8 - 1
    ck = torch.full((max(ops.ckpt_elems(b, kd, L, N), 4),), float("nan"), device=dev)
    ops.launch_fwd(u, delta, A, B, C, D, None, bias, True, out, None, None, ck, None)
    got = ck[: b * kd * n_ck * N].view(b, kd, n_ck, N)
    assert not torch.isnan(got).any()
    # fp64 recurrence, all rows at once
    dl = torch.nn.functional.softplus(delta.double() + bias.double()[None, :, None])
    Bd = B.double().repeat_interleave(kd // K, dim=1)                    # (b, kd, N, L)
    x = torch.zeros(b, kd, N, device=dev, dtype=torch.float64)
    ref = torch.empty(b, kd, n_ck, N, device=dev, dtype=torch.float64)
    for l in range(L):
        x = torch.exp(dl[:, :, l, None] * A.double()[None]) * x + (dl[:, :, l] * u[:, :, l].double())[..., None] * Bd[..., l]
        if (l + 1) % 8 == 0 and (l + 1) // 8 - 1 < n_ck:
            ref[:, :, (l + 1) // 8 - 1] = x
    scale = max(1.0, float(ref.abs().max()))
    torch.testing.assert_close(got.double(), ref, rtol=1e-4, atol=1e-5 * scale)
