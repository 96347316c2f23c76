"""LayerNorm on the caller side of the SS2D blocks (VSSBlock.ln_1 and the patch merging / expanding norms,
/root/reference/code/networks/mamba_sys.py:205,242,552,559,756-757) through the short-row kernels of libselscan_b200.

`LayerNorm` subclasses torch.nn.LayerNorm (same parameters, same state-dict keys); `patch_layernorms(model)` retargets the
nn.LayerNorm modules of an already-built reference model.  fp32 CUDA inputs with dim <= 1536 take the kernels; anything else
goes to torch's own CUDA / CPU implementation of the same formula."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        lib = _lib.load()
        D = x.shape[-1]
        x2 = x.contiguous().view(-1, D)
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        save = any(ctx.needs_input_grad)
        mean = x2.new_empty((rows,)) if save else None
        rstd = x2.new_empty((rows,)) if save else None
        weight, bias = weight.contiguous(), bias.contiguous()
        with torch.cuda.device(x.device):
            _lib.check(lib.selscan_b200_layernorm_fwd(x2.data_ptr(), weight.data_ptr(), bias.data_ptr(), float(eps), y.data_ptr(),
                                                      mean.data_ptr() if save else None, rstd.data_ptr() if save else None, rows, D,
                                                      torch.cuda.current_stream(x.device).cuda_stream), "selscan_b200_layernorm_fwd")
        if save:
            ctx.save_for_backward(x2, mean, rstd, weight)
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, weight = ctx.saved_tensors
        lib = _lib.load()
        rows, D = x2.shape
        dy2 = dy.contiguous().view(rows, D)
        dx = torch.empty_like(x2)
        part = x2.new_empty((int(lib.selscan_b200_layernorm_partial_elems(rows, D)),))
        with torch.cuda.device(x2.device):
            _lib.check(lib.selscan_b200_layernorm_bwd(dy2.data_ptr(), x2.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                      weight.data_ptr(), dx.data_ptr(), part.data_ptr(), rows, D,
                                                      torch.cuda.current_stream(x2.device).cuda_stream), "selscan_b200_layernorm_bwd")
        part = part.view(-1, 2, D).sum(0)
        return dx.view(dy.shape), part[0], part[1], None


def layer_norm(x, weight, bias, eps=1e-5):
    """F.layer_norm(x, (dim,), weight, bias, eps) over the last dimension."""
    if (x.is_cuda and x.dtype == torch.float32 and weight is not None and bias is not None and weight.dtype == torch.float32
            and x.numel() > 0 and _lib.load().selscan_b200_layernorm_supported(x.shape[-1])):
        return LayerNormFn.apply(x, weight, bias, eps)
    return F.layer_norm(x, (x.shape[-1],), weight, bias, eps)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        if len(self.normalized_shape) != 1:
            return super().forward(x)
        return layer_norm(x, self.weight, self.bias, self.eps)


def patch_layernorms(model):
    """Route every single-axis nn.LayerNorm of `model` through the kernels (parameters and state dict untouched)."""
    n = 0
    for m in model.modules():
        if type(m) is nn.LayerNorm and len(m.normalized_shape) == 1 and m.elementwise_affine:
            m.__class__ = LayerNorm
            n += 1
    return n
