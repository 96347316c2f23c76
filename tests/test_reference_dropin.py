"""The drop-in boundary against the UNCHANGED reference model code (build container only: needs /root/reference; CPU).

  * `code/networks/mamba_sys.py` (imported by file, unchanged) resolves `selective_scan_fn` / `selective_scan_ref` from this repo's
    `mamba_ssm.ops.selective_scan_interface` (mamba_sys.py:17-20), with the reference's exact call signature;
  * `ss2d.patch_ss2d` installs the B200 core on the real `SS2D` class before model construction (`self.forward_core =
    self.forward_corev0` is bound in __init__, :332) and the patched model keeps the reference's state-dict (names, shapes);
  * the call the reference makes into the op (argument convention, shapes, dtypes, strides) still is what the committed trace
    fixtures say (tests/golden/ss2d_call_*.npz, replayed on the GPU by tests/test_reference_call_trace_gpu.py).
"""
import importlib.util
import inspect
import os
import sys
import types

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


@pytest.fixture()
def ref_mamba_sys(monkeypatch):
    """mamba_sys.py imported unchanged, with stand-ins ONLY for timm / fvcore (absent from this image); `mamba_ssm` is this repo's."""
    keep = {k: sys.modules.get(k) for k in ("timm", "timm.models", "timm.models.layers", "fvcore", "fvcore.nn")}

    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            return x

    _stub("timm"); _stub("timm.models")
    _stub("timm.models.layers", DropPath=DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    _stub("fvcore")
    _stub("fvcore.nn", FlopCountAnalysis=None, flop_count_str=None, flop_count=None, parameter_count=None)
    for k in [k for k in sys.modules if k == "mamba_ssm" or k.startswith("mamba_ssm.")]:
        del sys.modules[k]                    # whatever an earlier test bound there: resolve from sys.path (this repo's package)
    spec = importlib.util.spec_from_file_location("_ref_mamba_sys_dropin", os.path.join(REF, "code", "networks", "mamba_sys.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    yield mod
    for k, v in keep.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


def test_reference_model_imports_our_op(ref_mamba_sys):
    from selscan_b200 import ops

    assert ref_mamba_sys.selective_scan_fn is ops.selective_scan_fn
    assert ref_mamba_sys.selective_scan_ref is ops.selective_scan_ref


def test_signature_matches_reference_op():
    """Parameter names, order and defaults of selective_scan_fn / selective_scan_ref (selective_scan_interface.py:77-78, :86-87)."""
    for n in ("causal_conv1d", "causal_conv1d_cuda", "selective_scan_cuda"):
        if n not in sys.modules:
            _stub(n, causal_conv1d_fn=None)
    spec = importlib.util.spec_from_file_location("_ref_op_sig", os.path.join(REF, "mamba", "mamba_ssm", "ops", "selective_scan_interface.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from selscan_b200 import ops

    for name in ("selective_scan_fn", "selective_scan_ref"):
        a, b = inspect.signature(getattr(ref, name)), inspect.signature(getattr(ops, name))
        assert [(p.name, p.default, p.kind) for p in a.parameters.values()] == [(p.name, p.default, p.kind) for p in b.parameters.values()], name
    assert issubclass(ops.SelectiveScanFn, torch.autograd.Function)


def test_patch_ss2d_on_the_real_class(ref_mamba_sys):
    from selscan_b200 import ss2d, vssm

    cls = ref_mamba_sys.SS2D
    orig_core, orig_fwd = cls.forward_corev0, cls.forward
    try:
        ss2d.patch_ss2d(cls)
        assert cls.forward_corev0 is ss2d.forward_core_b200 and cls.forward is ss2d.forward_b200
        model = ref_mamba_sys.VSSM(patch_size=4, in_chans=3, num_classes=4, depths=[1, 1, 1, 1], dims=[16, 32, 64, 128], drop_path_rate=0.0)
        blocks = [m for m in model.modules() if isinstance(m, cls)]
        assert len(blocks) == 7
        for m in blocks:   # bound at construction from the patched class attribute (mamba_sys.py:332)
            assert m.forward_core.__func__ is ss2d.forward_core_b200
            assert ss2d.fused_supported(m, 16, 16)      # the attributes forward_b200 reads exist on the real module
            for attr in ("in_proj", "conv2d", "x_proj_weight", "dt_projs_weight", "dt_projs_bias", "A_logs", "Ds", "out_norm", "out_proj",
                         "dt_rank", "d_state", "d_inner"):
                assert hasattr(m, attr), attr
        ours = vssm.VSSM(depths=(1, 1, 1, 1), dims=(16, 32, 64, 128), drop_path_rate=0.0)
        a, b = model.state_dict(), ours.state_dict()
        assert list(a.keys()) == list(b.keys())
        assert all(a[k].shape == b[k].shape for k in a)
    finally:
        cls.forward_corev0, cls.forward = orig_core, orig_fwd


@pytest.mark.parametrize("name", ["ss2d_call_d32_7x7", "ss2d_call_d64_14x14"])
def test_call_trace_fixture_is_current(name):
    """Re-trace the reference's call into the op and compare its layout with the committed fixture."""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    try:
        import make_golden_call_trace as mk
        ref = mk.load_reference_vssm()
        d_model, H, W = mk.CASES[name]
        now = mk.layout_signature(mk.trace_case(ref, d_model, H, W))
        with np.load(os.path.join(HERE, "golden", name + ".npz"), allow_pickle=False) as f:
            was = mk.layout_signature({k: (f[k] if f[k].dtype.kind not in "US" else (str(f[k]) if f[k].ndim == 0 else f[k]))
                                       for k in f.files})
        assert now == was
        # the one call Mamba-UNet makes: fp32, variable B / C with 4 groups, z = None, softplus on, no last state (:420-426)
        assert now["n_positional"] == 6 and now["kw_names"] == ["delta_bias", "delta_softplus", "return_last_state", "z"]
        assert now["kw.z"] == -1 and now["kw.delta_softplus"] == 1 and now["kw.return_last_state"] == 0
    finally:
        sys.path.remove(os.path.join(HERE, "golden"))
        for k in [k for k in sys.modules if k == "mamba_ssm" or k.startswith("mamba_ssm.")]:
            del sys.modules[k]               # the tracer bound the reference's selective_scan_ref there: do not leak it
