"""Loader for the REFERENCE's own selective_scan_cuda extension rebuilt for sm_100a (oracle/_ref/, built by
oracle/build_ref.py in the build container) -- TEST / BASELINE INFRASTRUCTURE ONLY.

It is the GPU-side second oracle (parity on identical inputs) and the "reference kernels on B200" timing baseline.
The calls below are the ones the reference's autograd wrapper makes
(/root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py:37 and :62-65).
"""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "selective_scan_cuda.so")
_mod = None


def available():
    return os.path.exists(SO)


def load():
    global _mod
    if _mod is None:
        import torch  # noqa: F401  (libtorch symbols must be loaded first)

        spec = importlib.util.spec_from_file_location("selective_scan_cuda", SO)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
    return _mod


def ref_fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus):
    """-> (out, x[, out_z]) exactly as selective_scan_cuda.fwd returns them."""
    return load().fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus)


def ref_bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out, delta_softplus):
    """-> (du, ddelta, dA, dB, dC, dD, ddelta_bias[, dz])."""
    return load().bwd(u, delta, A, B, C, D, z, delta_bias, dout, x, out, None, delta_softplus, False)
