"""Pin the C oracle (oracle/selscan_oracle.c) against the reference's own selective_scan_ref.

The golden fixtures were produced by /root/reference's selective_scan_ref + torch autograd
(tests/golden/make_golden.py); the reference itself computes in fp32, so the fp64 oracle is expected to
agree with it to fp32 round-off of the reference, and the fp32 oracle to a few ulps more.
"""
import numpy as np
import pytest

from conftest import golden_names, load_golden

GRADS = ["du", "ddelta", "dA", "dB", "dC", "dD", "dz", "ddelta_bias"]


def _run(orc, g, precision):
    sp = bool(int(g["delta_softplus"]))
    out, last = orc.oracle_fwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g.get("D"), g.get("z"),
                               g.get("delta_bias"), sp, return_last_state=True, precision=precision)
    grads = orc.oracle_bwd(g["u"], g["delta"], g["A"], g["B"], g["C"], g.get("D"), g.get("z"),
                           g.get("delta_bias"), g["dout"], sp, precision=precision)
    return out, last, grads


@pytest.mark.parametrize("precision", [64, 32])
@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(oracle, name, precision):
    g = load_golden(name)
    out, last, grads = _run(oracle, g, precision)
    # forward: north-star tolerance (rtol 1e-4 / atol 1e-5), scaled by the output magnitude for atol
    scale = max(1.0, float(np.abs(g["out"]).max()))
    np.testing.assert_allclose(out, g["out"], rtol=1e-4, atol=1e-5 * scale)
    np.testing.assert_allclose(last, g["last_state"], rtol=1e-4, atol=1e-5 * scale)
    for k in GRADS:
        if k not in g:
            assert grads[k] is None or k == "dz", k
            continue
        got = grads[k]
        if k in ("dB", "dC") and g[k].ndim == 3:  # squeezed 3-D B/C (selective_scan_interface.py:67-68)
            got = got[:, 0]
        gs = max(1.0, float(np.abs(g[k]).max()))
        np.testing.assert_allclose(got, g[k], rtol=1e-3, atol=1e-4 * gs, err_msg=k)


def test_oracle_f32_vs_f64_close(oracle):
    inp = oracle.make_inputs(2, 16, 300, 16, 4, dist="M", seed=11)
    o64 = oracle.oracle_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None,
                            inp["delta_bias"], True, precision=64)
    o32 = oracle.oracle_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None,
                            inp["delta_bias"], True, precision=32)
    np.testing.assert_allclose(o32, o64, rtol=1e-4, atol=1e-5)


def test_oracle_gradients_by_finite_difference(oracle):
    """Independent check of the analytic backward: central differences through the fp64 forward."""
    inp = oracle.make_inputs(1, 4, 12, 4, 2, dist="T", seed=3, has_z=True)
    args = dict(delta_softplus=True, precision=64)

    def loss(**over):
        a = {**inp, **over}
        out = oracle.oracle_fwd(a["u"], a["delta"], a["A"], a["B"], a["C"], a["D"], a["z"], a["delta_bias"],
                                **args)
        return float((out.astype(np.float64) * inp["dout"]).sum())

    grads = oracle.oracle_bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                              inp["delta_bias"], inp["dout"], True, precision=64)
    rng = np.random.default_rng(0)
    eps = 1e-2  # inputs are fp32; the forward is evaluated in fp64 but returns fp32, keep the step large
    for key, gk in [("u", "du"), ("delta", "ddelta"), ("A", "dA"), ("B", "dB"), ("C", "dC"), ("D", "dD"),
                    ("z", "dz"), ("delta_bias", "ddelta_bias")]:
        base = inp[key]
        for _ in range(4):
            idx = tuple(rng.integers(0, s) for s in base.shape)
            p, m = base.copy(), base.copy()
            p[idx] += eps
            m[idx] -= eps
            fd = (loss(**{key: p}) - loss(**{key: m})) / (float(p[idx]) - float(m[idx]))
            assert abs(fd - grads[gk][idx]) <= 2e-2 * max(1.0, abs(fd)), (key, idx, fd, grads[gk][idx])
