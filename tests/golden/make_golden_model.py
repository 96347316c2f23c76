"""Golden fixture for the caller context: forward + backward of the REFERENCE's own VSSM (code/networks/mamba_sys.py:694-829, with its
SS2D.forward_corev0 :396-436 running the reference selective_scan_ref on CPU), on a reduced-width configuration.

    python tests/golden/make_golden_model.py         (build container only: needs /root/reference)

The reference file is imported unchanged with stand-ins for the packages this image lacks (timm's DropPath / trunc_normal_, fvcore's
flop counters -- SURVEY.md appendix C) and with `mamba_ssm.ops.selective_scan_interface` bound to the reference's own
selective_scan_ref (selective_scan_fn needs its CUDA extension).  Stored: the state dict, the input, the logits, and the gradients of
a seeded linear loss w.r.t. the input and a sample of parameters.  tests/test_vssm_gpu.py loads the state dict into
selscan_b200.vssm.VSSM and must reproduce all of it on the GPU kernels.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_CODE = "/root/reference/code"
REF_OP = "/root/reference/mamba/mamba_ssm/ops/selective_scan_interface.py"


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference_vssm():
    class DropPath(torch.nn.Module):
        def __init__(self, p=0.0):
            super().__init__()
            self.p = p

        def forward(self, x):
            assert not self.training or self.p == 0.0
            return x

    _stub("timm")
    _stub("timm.models")
    _stub("timm.models.layers", DropPath=DropPath, trunc_normal_=torch.nn.init.trunc_normal_)
    _stub("fvcore")
    _stub("fvcore.nn", FlopCountAnalysis=None, flop_count_str=None, flop_count=None, parameter_count=None)
    for n in ("causal_conv1d", "causal_conv1d_cuda", "selective_scan_cuda"):
        _stub(n, causal_conv1d_fn=None)
    spec = importlib.util.spec_from_file_location("_ref_op", REF_OP)
    op = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(op)
    _stub("mamba_ssm")
    _stub("mamba_ssm.ops")
    _stub("mamba_ssm.ops.selective_scan_interface", selective_scan_fn=op.selective_scan_ref, selective_scan_ref=op.selective_scan_ref)
    spec = importlib.util.spec_from_file_location("_ref_mamba_sys", os.path.join(REF_CODE, "networks", "mamba_sys.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference_vssm()
    torch.manual_seed(1337)
    cfg = dict(patch_size=4, in_chans=3, num_classes=4, depths=[1, 1, 1, 1], dims=[16, 32, 64, 128], drop_path_rate=0.0)
    model = ref.VSSM(**cfg).eval()
    with torch.no_grad():  # move the special parameters away from their symmetric initial values
        for n, p in model.named_parameters():
            if n.endswith("A_logs") or n.endswith("Ds") or n.endswith("dt_projs_bias"):
                p.add_(0.1 * torch.randn_like(p))
    x = torch.rand(2, 3, 64, 64, requires_grad=True)
    out = model(x)
    dout = torch.randn_like(out)
    (out * dout).sum().backward()
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    grads = {}
    for n, p in model.named_parameters():
        if ("layers.0.blocks.0" in n or "layers.3.blocks.0.self_attention" in n or "layers_up.3.blocks.0.self_attention" in n
                or n.startswith("patch_embed") or n.startswith("output")):
            grads["grad." + n] = p.grad.numpy()
    path = os.path.join(HERE, "model_vssm_small.npz")
    np.savez_compressed(path, x=x.detach().numpy(), out=out.detach().numpy(), dout=dout.numpy(), dx=x.grad.numpy(),
                        **{"sd." + k: v for k, v in sd.items()}, **grads)
    print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB; params", sum(p.numel() for p in model.parameters()),
          "out abs max", float(out.abs().max()))


if __name__ == "__main__":
    main()
