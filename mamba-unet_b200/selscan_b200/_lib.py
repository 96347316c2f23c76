"""ctypes binding of libselscan_b200.so (C ABI declared in include/selscan_b200.h).

There is no fallback: if the library is missing or fails to load, every op raises RuntimeError.
"""
import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# SELSCAN_B200_LIB: another build of the same library (A/B timing of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("SELSCAN_B200_LIB") or os.path.join(_PKG, "lib", "libselscan_b200.so")

ABI_VERSION = 7
CKPT_INTERVAL = 8
STATE_PAD = 16

_i32, _i64, _ptr = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p


# fused dt_proj inputs, the tail of both argument structs
_DT_FIELDS = [("dt_w", _ptr), ("dt_x", _ptr), ("dt_w_d_stride", _i64), ("dt_x_batch_stride", _i64), ("dt_x_group_stride", _i64),
              ("dt_x_r_stride", _i64), ("dt_rank", _i32), ("mirror_pairs", _i32)]


class FwdArgs(ctypes.Structure):
    """struct selscan_fwd_args -- field order must match include/selscan_b200.h (tests/test_abi.py checks)."""
    _fields_ = (
        [(n, _i32) for n in ("batch", "dim", "seqlen", "dstate", "ngroups", "delta_softplus")]
        + [(n, _ptr) for n in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias")]
        + [(n, _i64) for n in (
            "u_batch_stride", "u_d_stride", "delta_batch_stride", "delta_d_stride", "A_d_stride", "A_n_stride",
            "B_batch_stride", "B_group_stride", "B_n_stride", "B_l_stride",
            "C_batch_stride", "C_group_stride", "C_n_stride", "C_l_stride",
            "z_batch_stride", "z_d_stride")]
        + [("out", _ptr), ("out_batch_stride", _i64), ("out_d_stride", _i64),
           ("out_z", _ptr), ("out_z_batch_stride", _i64), ("out_z_d_stride", _i64),
           ("last_state", _ptr), ("ckpt", _ptr), ("workspace", _ptr)]
        + _DT_FIELDS
    )


class BwdArgs(ctypes.Structure):
    """struct selscan_bwd_args."""
    _fields_ = (
        [(n, _i32) for n in ("batch", "dim", "seqlen", "dstate", "ngroups", "delta_softplus")]
        + [(n, _ptr) for n in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias", "dout", "out", "ckpt")]
        + [(n, _i64) for n in (
            "u_batch_stride", "u_d_stride", "delta_batch_stride", "delta_d_stride", "A_d_stride", "A_n_stride",
            "B_batch_stride", "B_group_stride", "B_n_stride", "B_l_stride",
            "C_batch_stride", "C_group_stride", "C_n_stride", "C_l_stride",
            "z_batch_stride", "z_d_stride", "dout_batch_stride", "dout_d_stride",
            "out_batch_stride", "out_d_stride", "du_batch_stride", "du_d_stride",
            "ddelta_batch_stride", "ddelta_d_stride", "dz_batch_stride", "dz_d_stride")]
        + [(n, _ptr) for n in ("du", "ddelta", "dz", "dA", "dB", "dC", "dD", "ddelta_bias")]
        + _DT_FIELDS
    )


_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library once; raise RuntimeError (never fall back) if it is unavailable."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libselscan_b200.so not found at {LIB_PATH}: build it with `python mamba-unet_b200/build.py` "
                "(there is no CPU or PyTorch fallback for selective_scan_fn)")
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise RuntimeError(f"cannot load {LIB_PATH}: {e}") from e
        lib.selscan_b200_abi_version.restype = ctypes.c_int
        lib.selscan_b200_last_error.restype = ctypes.c_char_p
        lib.selscan_b200_bwd_kernel.restype = ctypes.c_char_p
        lib.selscan_b200_ckpt_elems.restype = ctypes.c_int64
        lib.selscan_b200_ckpt_elems.argtypes = [_i32] * 4
        lib.selscan_b200_fwd_workspace_elems.restype = ctypes.c_int64
        lib.selscan_b200_fwd_workspace_elems.argtypes = [_i32] * 5
        lib.selscan_b200_dt_fusable.restype = ctypes.c_int
        lib.selscan_b200_dt_fusable.argtypes = [_i32] * 6
        lib.selscan_b200_mirror_ok.restype = ctypes.c_int
        lib.selscan_b200_mirror_ok.argtypes = [_i32] * 5
        lib.selscan_b200_fwd.restype = ctypes.c_int
        lib.selscan_b200_fwd.argtypes = [ctypes.POINTER(FwdArgs), _ptr]
        lib.selscan_b200_bwd.restype = ctypes.c_int
        lib.selscan_b200_bwd.argtypes = [ctypes.POINTER(BwdArgs), _ptr]
        for fn in (lib.selscan_b200_cross_scan, lib.selscan_b200_cross_merge):
            fn.restype = ctypes.c_int
            fn.argtypes = [_ptr, _ptr, _i32, _i32, _i32, _i32, _i64, _ptr]
        _f32 = ctypes.c_float
        lib.selscan_b200_ss2d_in_fwd.restype = ctypes.c_int
        lib.selscan_b200_ss2d_in_fwd.argtypes = [_ptr, _i64, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _i64, _i32, _ptr]
        lib.selscan_b200_ss2d_in_bwd.restype = ctypes.c_int
        lib.selscan_b200_ss2d_in_bwd.argtypes = [_ptr, _ptr, _i64, _ptr, _ptr, _ptr, _i64, _ptr, _i32, _i32, _i32, _i32, _i64, _i32, _ptr]
        lib.selscan_b200_ss2d_out_partial_elems.restype = ctypes.c_int64
        lib.selscan_b200_ss2d_out_partial_elems.argtypes = [_i32] * 4
        lib.selscan_b200_ss2d_out_fwd.restype = ctypes.c_int
        lib.selscan_b200_ss2d_out_fwd.argtypes = [_ptr, _i64, _ptr, _i64, _ptr, _ptr, _f32, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32,
                                                  _i32, _ptr]
        lib.selscan_b200_ss2d_out_bwd.restype = ctypes.c_int
        lib.selscan_b200_ss2d_out_bwd.argtypes = [_ptr, _ptr, _i64, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr, _i64, _ptr, _i32, _i32,
                                                  _i32, _i32, _i32, _ptr]
        lib.selscan_b200_layernorm_supported.restype = ctypes.c_int
        lib.selscan_b200_layernorm_supported.argtypes = [_i32]
        lib.selscan_b200_layernorm_partial_elems.restype = ctypes.c_int64
        lib.selscan_b200_layernorm_partial_elems.argtypes = [_i64, _i32]
        lib.selscan_b200_layernorm_fwd.restype = ctypes.c_int
        lib.selscan_b200_layernorm_fwd.argtypes = [_ptr, _ptr, _ptr, _f32, _ptr, _ptr, _ptr, _i64, _i32, _ptr]
        lib.selscan_b200_layernorm_bwd.restype = ctypes.c_int
        lib.selscan_b200_layernorm_bwd.argtypes = [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _ptr]
        lib.selscan_b200_gemm_3xtf32.restype = ctypes.c_int
        lib.selscan_b200_gemm_3xtf32.argtypes = [_ptr, _i64, _i32, _ptr, _i64, _i32, _ptr, _i64, _i32, _i32, _i32, _i32, _i64, _i64,
                                                 _i64, _i32, _i32, _i32, _i32, _ptr]
        if lib.selscan_b200_abi_version() != ABI_VERSION:
            raise RuntimeError("libselscan_b200.so ABI version mismatch: rebuild with mamba-unet_b200/build.py")
        _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = load().selscan_b200_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
