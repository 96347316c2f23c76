"""Golden fixture of the CALL the unchanged reference model makes into the op: the exact positional / keyword arguments, shapes,
dtypes, strides and storage offsets that `SS2D.forward_corev0` (/root/reference/code/networks/mamba_sys.py:396-436) hands to
`selective_scan_fn` (:420-426), with the values, the op's result (the reference's own selective_scan_ref on CPU) and the whole
`SS2D.forward` (:527-540) input / output / gradients around it.

    python tests/golden/make_golden_call_trace.py          (build container only: needs /root/reference)

The reference model sources cannot travel to the GPU box; this trace can.  tests/test_reference_dropin.py re-traces the call in
the build container and fails if the committed fixture and the reference diverge; tests/test_reference_call_trace_gpu.py replays
the call -- same strides, same keywords -- through this repo's `mamba_ssm.ops.selective_scan_interface.selective_scan_fn` on the
GPU, and runs the patched SS2D block (ss2d.patch_ss2d) on the same parameters.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_model import load_reference_vssm  # noqa: E402

# (d_model, H, W): d_inner = 2 * d_model = 64 / 128 channels per direction -> the tiled kernels' shapes; L = 49 and 196
CASES = {"ss2d_call_d32_7x7": (32, 7, 7), "ss2d_call_d64_14x14": (64, 14, 14)}


def trace_case(ref, d_model, H, W, seed=7):
    """Run the reference SS2D once on CPU, recording the selective_scan_fn call.  Returns (record dict, module, x, y)."""
    calls = []
    real = ref.selective_scan_fn                          # = the reference's selective_scan_ref (make_golden_model's binding)

    def recording(*args, **kwargs):
        out = real(*args, **kwargs)
        calls.append((args, kwargs, out))
        return out

    ref.selective_scan_fn = recording
    try:
        torch.manual_seed(seed)
        m = ref.SS2D(d_model=d_model, d_state=16, dropout=0.0).eval()
        with torch.no_grad():
            for n, p in m.named_parameters():
                if n in ("A_logs", "Ds", "dt_projs_bias"):
                    p.add_(0.1 * torch.randn_like(p))
        x = torch.randn(2, H, W, d_model, requires_grad=True)
        y = m(x)
        dy = torch.randn_like(y)
        (y * dy).sum().backward()
    finally:
        ref.selective_scan_fn = real
    assert len(calls) == 1
    args, kwargs, out = calls[0]
    names = ["u", "delta", "A", "B", "C", "D"]
    rec = {"n_positional": len(args), "kw_names": sorted(kwargs)}
    tensors = dict(zip(names, args))
    tensors.update({k: v for k, v in kwargs.items() if torch.is_tensor(v)})
    for k, t in tensors.items():
        rec[f"arg.{k}.value"] = t.detach().numpy().copy()
        rec[f"arg.{k}.stride"] = np.array(t.stride(), dtype=np.int64)
        rec[f"arg.{k}.offset"] = np.int64(t.storage_offset())
        rec[f"arg.{k}.storage_numel"] = np.int64(t.untyped_storage().nbytes() // t.element_size())
        rec[f"arg.{k}.dtype"] = str(t.dtype)
    for k, v in kwargs.items():
        if not torch.is_tensor(v):
            rec[f"kw.{k}"] = np.array(-1 if v is None else int(v))
    rec["op_out"] = out.detach().numpy()
    rec["x"], rec["y"], rec["dy"], rec["dx"] = x.detach().numpy(), y.detach().numpy(), dy.numpy(), x.grad.numpy()
    for k, v in m.state_dict().items():
        rec["sd." + k] = v.numpy()
    for n, p in m.named_parameters():
        rec["grad." + n] = p.grad.numpy()
    return rec


def layout_signature(rec):
    """What must not drift: argument passing convention, shapes, dtypes, strides, offsets."""
    sig = {"n_positional": int(rec["n_positional"]), "kw_names": [str(k) for k in rec["kw_names"]]}
    for k in rec:
        if k.startswith("arg.") and not k.endswith(".value"):
            v = rec[k]
            sig[k] = v.tolist() if isinstance(v, np.ndarray) else (str(v) if isinstance(v, str) else int(v))
        elif k.startswith("arg.") and k.endswith(".value"):
            sig[k[:-6] + ".shape"] = list(rec[k].shape)
        elif k.startswith("kw."):
            sig[k] = int(rec[k])
    return sig


def main():
    ref = load_reference_vssm()
    for name, (d_model, H, W) in CASES.items():
        rec = trace_case(ref, d_model, H, W)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **rec)
        print("wrote", path, f"{os.path.getsize(path) / 1e6:.2f} MB", {k: v for k, v in layout_signature(rec).items() if "stride" in k})


if __name__ == "__main__":
    main()
