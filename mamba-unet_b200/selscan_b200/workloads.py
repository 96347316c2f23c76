"""Step bodies of the two training workloads BASELINE.json names, on synthetic tensors (the scripts themselves need ACDC data,
tensorboardX, medpy ... which neither this image nor the GPU box has -- SURVEY.md section 7, hard part 8).

    supervised_step   /root/reference/code/train_fully_supervised_2D_VIM.py:152-160   (bs 24, 0.5*(CE + Dice), SGD)
    semi_step         /root/reference/code/train_Semi_Mamba_UNet.py:210-250           (bs 16 = 8 labeled + 8 unlabeled, two networks,
                      cross pseudo supervision + ConstraLoss (code/utils/losses.py:169-181), one backward, two SGD steps)
"""
import math

import torch
import torch.nn.functional as F

from .vssm import DiceLoss


def consistency_weight(iter_num, consistency=0.1, rampup=200.0):
    """train_Semi_Mamba_UNet.py:126-128 with code/utils/ramps.py:20-28 (sigmoid ramp-up, evaluated every 150 iterations)."""
    cur = min(max(iter_num // 150, 0.0), rampup)
    phase = 1.0 - cur / rampup
    return consistency * math.exp(-5.0 * phase * phase)


def constra_loss(a, b):
    """MSE between the L2-normalised global-average-pooled logits of the two networks (losses.py:169-181)."""
    pa = F.normalize(a.mean(dim=(2, 3)), p=2, dim=1)
    pb = F.normalize(b.mean(dim=(2, 3)), p=2, dim=1)
    return ((pa - pb) ** 2).mean()


def supervised_step(model, opt, dice, x, y):
    out = model(x)
    loss = 0.5 * (dice(torch.softmax(out, dim=1), y.unsqueeze(1)) + F.cross_entropy(out, y.long()))
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


def semi_step(model1, model2, opt1, opt2, dice, x, y, labeled_bs, cw, side=None):
    """side: a second CUDA stream.  The two networks are independent until the losses, and at 16 images per network most of their
    kernels leave a B200 under-filled (the scan at stage 1: 192 CTAs on 296 slots); with `side` the second network's forward runs
    there, concurrently with the first one's, and autograd runs each network's backward on the stream of its forward -- the
    arithmetic is unchanged.  Also capturable (GraphedStep): the capture forks and joins the two streams."""
    if side is None:
        o1, o2 = model1(x), model2(x)
    else:
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)                 # x, y and both networks' parameters (previous optimizer step) are ready
        with torch.cuda.stream(side):
            o2 = model2(x)
        o1 = model1(x)
        cur.wait_stream(side)                 # join before the losses
        o2.record_stream(cur)
    s1, s2 = torch.softmax(o1, dim=1), torch.softmax(o2, dim=1)
    lb = labeled_bs
    loss1 = 0.5 * (F.cross_entropy(o1[:lb], y[:lb].long()) + dice(s1[:lb], y[:lb].unsqueeze(1)))
    loss2 = 0.5 * (F.cross_entropy(o2[:lb], y[:lb].long()) + dice(s2[:lb], y[:lb].unsqueeze(1)))
    p1 = torch.argmax(s1[lb:].detach(), dim=1)
    p2 = torch.argmax(s2[lb:].detach(), dim=1)
    ps1 = dice(s1[lb:], p2.unsqueeze(1))
    ps2 = dice(s2[lb:], p1.unsqueeze(1))
    con = constra_loss(o1, o2)
    loss = (loss1 + cw * ps1 + 0.5 * con) + (loss2 + cw * ps2 + 0.5 * con)
    opt1.zero_grad(set_to_none=True)
    opt2.zero_grad(set_to_none=True)
    loss.backward()
    opt1.step()
    opt2.step()
    return loss


class GraphedStep:
    """A whole static-shape step (forward, loss, backward, optimizer) captured once as a CUDA graph and replayed: at batch 24 a
    MambaUnet step is ~2600 kernel launches, and the host can no longer issue them as fast as a B200 retires them (SURVEY.md
    section 8f row 4).  `fn(*tensors)` must be sync-free; inputs are copied into the captured buffers before every replay."""

    def __init__(self, fn, *inputs, warmup=3, stream=None):
        self.inputs = [t.clone() for t in inputs]
        side = stream if stream is not None else torch.cuda.Stream()   # DDP: the stream the wrapper was constructed on
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # eager warm-up on a side stream: lazy optimizer state, kernel attributes, autotuning
            for _ in range(warmup):
                fn(*self.inputs)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=stream):
            self.out = fn(*self.inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out


def make_sgd(model, lr=0.01):
    return torch.optim.SGD(model.parameters(), lr=lr, momentum=0.9, weight_decay=1e-4)


__all__ = ["supervised_step", "semi_step", "make_sgd", "GraphedStep", "consistency_weight", "constra_loss", "DiceLoss"]
