"""numpy/ctypes front-end of oracle/selscan_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The C file restates /root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py:86-152
(selective_scan_ref) and the backward identities of
/root/reference/mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:278-296.  It is pinned against
the reference's own selective_scan_ref through tests/golden/ (see tests/golden/make_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libselscan_oracle.so")
_lib = None

_fp = ctypes.POINTER(ctypes.c_float)


def build_oracle(force=False):
    """Compile oracle/selscan_oracle.c with gcc (a few hundred ms).  Idempotent."""
    src = os.path.join(_HERE, "selscan_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []))
    return _SO


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        lib = ctypes.CDLL(_SO)
        lib.selscan_oracle_fwd.restype = ctypes.c_int
        lib.selscan_oracle_fwd.argtypes = [_fp] * 8 + [ctypes.c_int] * 6 + [_fp, _fp, ctypes.c_int]
        lib.selscan_oracle_bwd.restype = ctypes.c_int
        lib.selscan_oracle_bwd.argtypes = [_fp] * 9 + [ctypes.c_int] * 6 + [_fp] * 8 + [ctypes.c_int]
        lib.selscan_oracle_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def oracle_threads():
    return _load().selscan_oracle_threads()


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_fp)


def _canon_bc(X, batch):
    """(batch, N, L) -> (batch, 1, N, L); selective_scan_interface.py:31-36."""
    X = _c(X)
    if X.ndim == 3:
        X = X.reshape(batch, 1, X.shape[1], X.shape[2])
    return X


def oracle_fwd(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
               return_last_state=False, precision=64):
    """Same argument meaning as selective_scan_ref (selective_scan_interface.py:86-100); numpy in/out."""
    u, delta, A = _c(u), _c(delta), _c(A)
    batch, dim, L = u.shape
    N = A.shape[1]
    B, C = _canon_bc(B, batch), _canon_bc(C, batch)
    G = B.shape[1]
    assert B.shape == (batch, G, N, L) and C.shape == (batch, G, N, L), (B.shape, C.shape)
    D, z, delta_bias = _c(D), _c(z), _c(delta_bias)
    out = np.empty((batch, dim, L), np.float32)
    last = np.empty((batch, dim, N), np.float32) if return_last_state else None
    rc = _load().selscan_oracle_fwd(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(z), _p(delta_bias),
                                    int(bool(delta_softplus)), batch, dim, L, N, G, _p(out), _p(last),
                                    precision)
    if rc != 0:
        raise ValueError("selscan_oracle_fwd: bad arguments")
    return (out, last) if return_last_state else out


def oracle_bwd(u, delta, A, B, C, D, z, delta_bias, dout, delta_softplus=False, precision=64):
    """Gradients of sum(out * dout) w.r.t. every input; returns a dict of numpy arrays.

    dB / dC come back in the 4-D (batch, G, N, L) layout (the caller squeezes for 3-D inputs, as
    selective_scan_interface.py:67-68 does)."""
    u, delta, A, dout = _c(u), _c(delta), _c(A), _c(dout)
    batch, dim, L = u.shape
    N = A.shape[1]
    B, C = _canon_bc(B, batch), _canon_bc(C, batch)
    G = B.shape[1]
    D, z, delta_bias = _c(D), _c(z), _c(delta_bias)
    du = np.empty_like(u)
    ddelta = np.empty_like(u)
    dA = np.empty((dim, N), np.float32)
    dB = np.empty_like(B)
    dC = np.empty_like(C)
    dD = np.empty((dim,), np.float32)
    dbias = np.empty((dim,), np.float32)
    dz = np.empty_like(u) if z is not None else None
    rc = _load().selscan_oracle_bwd(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(z), _p(delta_bias),
                                    _p(dout), int(bool(delta_softplus)), batch, dim, L, N, G, _p(du),
                                    _p(ddelta), _p(dA), _p(dB), _p(dC), _p(dD), _p(dz), _p(dbias),
                                    precision)
    if rc != 0:
        raise ValueError("selscan_oracle_bwd: bad arguments")
    return {"du": du, "ddelta": ddelta, "dA": dA, "dB": dB, "dC": dC,
            "dD": dD if D is not None else None, "dz": dz,
            "ddelta_bias": dbias if delta_bias is not None else None}


def make_inputs(batch, dim, L, N=16, G=4, dist="T", seed=0, has_z=False, has_D=True, has_bias=True):
    """Seeded synthetic inputs.

    dist "T": the reference test's distributions (mamba/tests/ops/test_selective_scan.py:58-88):
              u,B,C,z ~ N(0,1); delta ~ 0.5*U(0,1); A ~ -0.5*U(0,1); delta_bias ~ 0.5*U(0,1); D ~ N(0,1).
    dist "M": model-like (code/networks/mamba_sys.py:353-361,369-375,385-394): A = -(1..N) per row, D = 1,
              delta_bias = softplus^-1(exp(U(log 1e-3, log 1e-1))), delta ~ 0.5*N(0,1), u,B,C ~ N(0,1).
    Returns a dict of float32 numpy arrays (z / D / delta_bias may be None) plus dout ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    f = np.float32
    u = rng.standard_normal((batch, dim, L)).astype(f)
    Bm = rng.standard_normal((batch, G, N, L)).astype(f)
    Cm = rng.standard_normal((batch, G, N, L)).astype(f)
    dout = rng.standard_normal((batch, dim, L)).astype(f)
    z = rng.standard_normal((batch, dim, L)).astype(f) if has_z else None
    if dist == "T":
        delta = (0.5 * rng.random((batch, dim, L))).astype(f)
        A = (-0.5 * rng.random((dim, N))).astype(f)
        bias = (0.5 * rng.random((dim,))).astype(f)
        D = rng.standard_normal((dim,)).astype(f)
    elif dist == "M":
        delta = (0.5 * rng.standard_normal((batch, dim, L))).astype(f)
        A = np.tile(-np.arange(1, N + 1, dtype=f), (dim, 1))
        dt = np.exp(rng.random((dim,)) * (np.log(1e-1) - np.log(1e-3)) + np.log(1e-3))
        bias = (dt + np.log(-np.expm1(-dt))).astype(f)
        D = np.ones((dim,), f)
    else:
        raise ValueError(dist)
    return {"u": u, "delta": delta, "A": A, "B": Bm, "C": Cm, "D": D if has_D else None, "z": z,
            "delta_bias": bias if has_bias else None, "dout": dout}
