"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, the ctypes structs match the header field-for-field, and argument validation / error reporting
work without a GPU (no kernel is launched here)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "selscan_b200.h")


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, os.path.join(ROOT, "mamba-unet_b200"))
    import build as _build  # mamba-unet_b200/build.py
    _build.build()
    from selscan_b200 import _lib
    return _lib


def _header_text():
    with open(HEADER) as f:
        return re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)


def _struct_fields(text, name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), text, flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        m = re.match(r"(const float\*|float\*|int32_t|int64_t) (.*)", decl)
        assert m, decl
        kind = {"const float*": "ptr", "float*": "ptr", "int32_t": "i32", "int64_t": "i64"}[m.group(1)]
        for nm in m.group(2).split(","):
            fields.append((nm.strip(), kind))
    return fields


def test_header_symbols_are_exported(lib):
    text = _header_text()
    declared = set(re.findall(r"\b(selscan_b200_\w+)\s*\(", text))
    assert {"selscan_b200_fwd", "selscan_b200_bwd", "selscan_b200_ckpt_elems", "selscan_b200_last_error",
            "selscan_b200_abi_version"} <= declared
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib.LIB_PATH], text=True)
    exported = set(re.findall(r" T (selscan_b200_\w+)", out))
    assert declared <= exported, declared - exported
    # nothing but the C ABI leaks out of the library
    leaked = [s for s in re.findall(r" [TW] (\S+)", out) if not s.startswith("selscan_b200_")]
    assert not leaked, leaked[:5]


@pytest.mark.parametrize("cname,pyname", [("selscan_fwd_args", "FwdArgs"), ("selscan_bwd_args", "BwdArgs")])
def test_ctypes_structs_match_header(lib, cname, pyname):
    want = _struct_fields(_header_text(), cname)
    kinds = {ctypes.c_void_p: "ptr", ctypes.c_int32: "i32", ctypes.c_int64: "i64"}
    got = [(n, kinds[t]) for n, t in getattr(lib, pyname)._fields_]
    assert got == want


def test_constants_match_header(lib):
    text = _header_text()
    for macro, val in (("SELSCAN_B200_ABI_VERSION", lib.ABI_VERSION), ("SELSCAN_B200_CKPT_INTERVAL", lib.CKPT_INTERVAL),
                       ("SELSCAN_B200_STATE_PAD", lib.STATE_PAD)):
        assert int(re.search(r"#define %s (\d+)" % macro, text).group(1)) == val
    assert lib.load().selscan_b200_abi_version() == lib.ABI_VERSION


def test_ckpt_elems(lib):
    L = lib.load()
    assert L.selscan_b200_ckpt_elems(1, 4, 8, 16) == 0
    assert L.selscan_b200_ckpt_elems(1, 4, 9, 16) == 4 * 16
    assert L.selscan_b200_ckpt_elems(24, 768, 3136, 16) == 24 * 768 * 391 * 16
    assert L.selscan_b200_ckpt_elems(0, 4, 100, 16) == 0
    assert L.selscan_b200_ckpt_elems(2, 4, 17, 40) == 3 * 2 * 4 * 2 * 16   # 3 state blocks of 16


def test_argument_validation_reports_errors(lib):
    L = lib.load()
    a = lib.FwdArgs(batch=1, dim=4, seqlen=8, dstate=16, ngroups=1)
    assert L.selscan_b200_fwd(a, None) == -1  # NULL pointers
    assert b"must not be NULL" in L.selscan_b200_last_error()
    a = lib.FwdArgs(batch=1, dim=4, seqlen=8, dstate=300, ngroups=1)
    assert L.selscan_b200_fwd(a, None) == -1
    assert b"state dimension" in L.selscan_b200_last_error()
    a = lib.FwdArgs(batch=1, dim=6, seqlen=8, dstate=16, ngroups=4)
    assert L.selscan_b200_fwd(a, None) == -1
    assert b"divisible" in L.selscan_b200_last_error()
    assert L.selscan_b200_fwd(None, None) == -1
    b = lib.BwdArgs(batch=1, dim=4, seqlen=8, dstate=16, ngroups=1)
    assert L.selscan_b200_bwd(b, None) == -1


def test_op_refuses_cpu_tensors_and_ref_runs_on_cpu(lib):
    import torch
    from mamba_ssm.ops.selective_scan_interface import selective_scan_fn, selective_scan_ref

    u = torch.randn(1, 4, 8)
    with pytest.raises(RuntimeError, match="CUDA"):
        selective_scan_fn(u, u, torch.zeros(4, 16), torch.zeros(1, 1, 16, 8), torch.zeros(1, 1, 16, 8))
    out, last = selective_scan_ref(u, u.abs(), -torch.rand(4, 16), torch.randn(1, 2, 16, 8), torch.randn(1, 2, 16, 8),
                                   torch.ones(4), None, torch.zeros(4), True, True)
    assert out.shape == (1, 4, 8) and last.shape == (1, 4, 16)


def test_shipped_ref_matches_oracle(lib, oracle):
    """The selective_scan_ref we export (API mirror) agrees with the pinned oracle, forward and autograd backward."""
    import numpy as np
    import torch
    from mamba_ssm.ops.selective_scan_interface import selective_scan_ref

    inp = oracle.make_inputs(2, 8, 33, 16, 2, dist="T", seed=9, has_z=True)
    t = {k: (torch.from_numpy(v).requires_grad_(k != "dout") if v is not None else None) for k, v in inp.items()}
    out, last = selective_scan_ref(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], t["z"], t["delta_bias"], True, True)
    out.backward(t["dout"])
    ref_out, ref_last = oracle.oracle_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                                          inp["delta_bias"], True, return_last_state=True)
    ref_g = oracle.oracle_bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                              inp["delta_bias"], inp["dout"], True)
    np.testing.assert_allclose(out.detach().numpy(), ref_out, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(last.detach().numpy(), ref_last, rtol=1e-4, atol=1e-4)
    for k, name in (("u", "du"), ("delta", "ddelta"), ("A", "dA"), ("B", "dB"), ("C", "dC"), ("D", "dD"), ("z", "dz"),
                    ("delta_bias", "ddelta_bias")):
        np.testing.assert_allclose(t[k].grad.numpy(), ref_g[name], rtol=1e-3, atol=1e-3, err_msg=name)
