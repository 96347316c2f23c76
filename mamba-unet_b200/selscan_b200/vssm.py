"""Caller context of the hot path: a from-scratch MambaUnet / VSSM whose SS2D blocks run on the sm_100a selective scan.

Why it exists: the metric names "MambaUNet 224^2 train img/s", the reference model code cannot travel to the GPU box, and the
scan must be measured inside its real caller.  The architecture and every parameter name follow the reference so that its
checkpoints load with `load_state_dict` (tests/test_vssm_gpu.py checks outputs against a golden forward of the reference model):
    VSSM          /root/reference/code/networks/mamba_sys.py:694-829   (encoder 4 stages, decoder 3 stages + skip concat, x4 head)
    VSSBlock      :543-560     x + DropPath(SS2D(LayerNorm(x)))
    SS2D          :267-338, :527-540   in_proj -> depthwise 3x3 conv + SiLU -> forward_core -> * silu(z) -> out_proj
    PatchEmbed2D  :165-188,  PatchMerging2D :191-230,  PatchExpand :232-246,  FinalPatchExpand_X4 :248-263
    MambaUnet     /root/reference/code/networks/vision_mamba.py:23-46  (1 -> 3 channel repeat, attribute `mamba_unet`)
Linear layers / GEMMs are stock torch (cuBLAS); everything between in_proj and out_proj of an SS2D block runs on this
repo's kernels (`ss2d.forward_b200`: conv + SiLU + CrossScan, scan, CrossMerge + LayerNorm + gate).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .layernorm import LayerNorm
from .ss2d import forward_b200, forward_core_b200


class DropPath(nn.Module):
    """Per-sample stochastic depth (what the reference takes from timm, mamba_sys.py:12,556)."""

    def __init__(self, p=0.0):
        super().__init__()
        self.p = float(p)

    def forward(self, x):
        if self.p == 0.0 or not self.training:
            return x
        keep = 1.0 - self.p
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


class SS2D(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=3, expand=2, dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4):
        super().__init__()
        self.d_model, self.d_state = d_model, d_state
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16)
        D, R, N, K = self.d_inner, self.dt_rank, d_state, 4
        self.in_proj = nn.Linear(d_model, 2 * D, bias=False)
        self.conv2d = nn.Conv2d(D, D, kernel_size=d_conv, padding=(d_conv - 1) // 2, groups=D, bias=True)
        self.x_proj_weight = nn.Parameter(torch.empty(K, R + 2 * N, D))
        self.dt_projs_weight = nn.Parameter(torch.empty(K, D, R))
        self.dt_projs_bias = nn.Parameter(torch.empty(K, D))
        self.A_logs = nn.Parameter(torch.empty(K * D, N))
        self.Ds = nn.Parameter(torch.ones(K * D))
        self.out_norm = LayerNorm(D)
        self.out_proj = nn.Linear(D, d_model, bias=False)
        self.A_logs._no_weight_decay = True
        self.Ds._no_weight_decay = True
        with torch.no_grad():   # the reference's initialisation (mamba_sys.py:341-394)
            bound = 1.0 / math.sqrt(D)
            self.x_proj_weight.uniform_(-bound, bound)                       # nn.Linear default for the four x_proj
            self.dt_projs_weight.uniform_(-R ** -0.5, R ** -0.5)
            dt = torch.exp(torch.rand(K, D) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min)).clamp(min=dt_init_floor)
            self.dt_projs_bias.copy_(dt + torch.log(-torch.expm1(-dt)))      # softplus^-1(dt)
            self.A_logs.copy_(torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(K * D, 1))

    forward_core = forward_core_b200
    fused = True        # conv / CrossScan / CrossMerge / LayerNorm / gate through the SS2D edge kernels (ss2d.forward_b200)

    def forward(self, x):                                   # (B, H, W, C), mamba_sys.py:527-540
        if self.fused:
            return forward_b200(self, x)
        xz = self.in_proj(x)
        x, z = xz.chunk(2, dim=-1)
        x = F.silu(self.conv2d(x.permute(0, 3, 1, 2).contiguous()))
        y = self.forward_core(x)
        return self.out_proj(y * F.silu(z))


class VSSBlock(nn.Module):
    def __init__(self, hidden_dim, drop_path=0.0, d_state=16):
        super().__init__()
        self.ln_1 = LayerNorm(hidden_dim)
        self.self_attention = SS2D(hidden_dim, d_state=d_state)
        self.drop_path = DropPath(drop_path)

    def forward(self, x):
        return x + self.drop_path(self.self_attention(self.ln_1(x)))


class PatchMerging2D(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = LayerNorm(4 * dim)

    def forward(self, x):                                   # (B, H, W, C) -> (B, H/2, W/2, 2C)
        x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], dim=-1)
        return self.reduction(self.norm(x))


class PatchExpand(nn.Module):
    """(B, H, W, C) -> (B, s*H, s*W, C_out): linear expand, pixel-shuffle in channels-last, LayerNorm."""

    def __init__(self, dim, scale, out_dim):
        super().__init__()
        self.scale, self.out_dim = scale, out_dim
        self.expand = nn.Linear(dim, scale * scale * out_dim, bias=False)
        self.norm = LayerNorm(out_dim)

    def forward(self, x):
        B, H, W, _ = x.shape
        s, c = self.scale, self.out_dim
        x = self.expand(x).view(B, H, W, s, s, c).permute(0, 1, 3, 2, 4, 5).reshape(B, H * s, W * s, c)
        return self.norm(x)


class VSSLayer(nn.Module):
    def __init__(self, dim, depth, drop_path, d_state, downsample=False, upsample=False):
        super().__init__()
        self.blocks = nn.ModuleList([VSSBlock(dim, drop_path[i], d_state) for i in range(depth)])
        self.downsample = PatchMerging2D(dim) if downsample else None
        self.upsample = PatchExpand(dim, 2, dim // 2) if upsample else None

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x)
        if self.downsample is not None:
            x = self.downsample(x)
        if self.upsample is not None:
            x = self.upsample(x)
        return x


class VSSM(nn.Module):
    def __init__(self, patch_size=4, in_chans=3, num_classes=4, depths=(2, 2, 2, 2), dims=(96, 192, 384, 768), d_state=16,
                 drop_path_rate=0.2):
        super().__init__()
        n = len(depths)
        self.num_layers = n
        d0 = dims[0]
        self.patch_embed = nn.Module()
        self.patch_embed.proj = nn.Conv2d(in_chans, d0, kernel_size=patch_size, stride=patch_size)
        self.patch_embed.norm = LayerNorm(d0)
        dpr = torch.linspace(0, drop_path_rate, sum(depths)).tolist()
        sl = lambda i: dpr[sum(depths[:i]):sum(depths[:i + 1])]
        self.layers = nn.ModuleList([VSSLayer(d0 * 2 ** i, depths[i], sl(i), d_state, downsample=(i < n - 1)) for i in range(n)])
        self.layers_up = nn.ModuleList()
        self.concat_back_dim = nn.ModuleList()
        for i in range(n):
            dim = d0 * 2 ** (n - 1 - i)
            if i == 0:
                self.layers_up.append(PatchExpand(dim, 2, dim // 2))
                self.concat_back_dim.append(nn.Identity())
            else:
                self.layers_up.append(VSSLayer(dim, depths[n - 1 - i], sl(n - 1 - i), d_state, upsample=(i < n - 1)))
                self.concat_back_dim.append(nn.Linear(2 * dim, dim))
        self.norm = LayerNorm(d0 * 2 ** (n - 1))
        self.norm_up = LayerNorm(d0)
        self.up = PatchExpand(d0, 4, d0)
        self.output = nn.Conv2d(d0, num_classes, kernel_size=1, bias=False)
        self.apply(self._init)

    @staticmethod
    def _init(m):                                            # mamba_sys.py:769-784
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)

    def forward(self, x):                                   # (B, 3, H, W) -> (B, classes, H, W)
        x = self.patch_embed.norm(self.patch_embed.proj(x).permute(0, 2, 3, 1))
        skips = []
        for layer in self.layers:
            skips.append(x)
            x = layer(x)
        x = self.norm(x)
        for i, up in enumerate(self.layers_up):
            if i > 0:
                x = self.concat_back_dim[i](torch.cat([x, skips[self.num_layers - 1 - i]], dim=-1))
            x = up(x)
        x = self.up(self.norm_up(x))                        # (B, H, W, d0)
        # the 1x1 `output` convolution (mamba_sys.py:761,823) as the per-pixel linear map it is, on the channels-last tensor: same
        # parameter (classes, d0, 1, 1), no NHWC -> NCHW copy of the widest activation, and a contiguous weight gradient (cuDNN's
        # came back in channels-last strides, which DDP's bucket views reject with a warning and a copy on every step)
        w = self.output.weight
        return F.linear(x, w.view(w.shape[0], -1)).permute(0, 3, 1, 2).contiguous()


class MambaUnet(nn.Module):
    """vision_mamba.py:23-46 with configs/vmamba_tiny.yaml (EMBED_DIM 96, DEPTHS [2,2,2,2], DROP_PATH_RATE 0.2)."""

    def __init__(self, num_classes=4, depths=(2, 2, 2, 2), dims=(96, 192, 384, 768), drop_path_rate=0.2, d_state=16):
        super().__init__()
        self.num_classes = num_classes
        self.mamba_unet = VSSM(in_chans=3, num_classes=num_classes, depths=depths, dims=dims, d_state=d_state,
                               drop_path_rate=drop_path_rate)

    def forward(self, x):
        if x.shape[1] == 1:
            x = x.repeat(1, 3, 1, 1)
        return self.mamba_unet(x)


class DiceLoss(nn.Module):
    """Soft Dice over one-hot targets, the loss of train_fully_supervised_2D_VIM.py:137,156 (code/utils/losses.py:332-368),
    without the per-class `.item()` host syncs."""

    def __init__(self, n_classes):
        super().__init__()
        self.n_classes = n_classes

    def forward(self, probs, target):                       # probs (B, C, H, W), target (B, 1, H, W)
        onehot = torch.zeros_like(probs).scatter_(1, target.long(), 1.0)
        dims = (0, 2, 3)
        inter = (probs * onehot).sum(dims)
        denom = (probs * probs).sum(dims) + (onehot * onehot).sum(dims)
        return (1.0 - (2 * inter + 1e-5) / (denom + 1e-5)).mean()
