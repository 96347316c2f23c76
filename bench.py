#!/usr/bin/env python
"""bench.py -- selective-scan fwd+bwd throughput on B200 (BASELINE.json metric).

Workload (config.workload): the 14 SS2D selective scans of ONE MambaUnet (vmamba_tiny, 224x224) supervised
training step at batch 24 -- stage shapes traced in SURVEY.md section 3.4: S1 x4 (KD=768, L=3136), S2 x4 (1536, 784),
S3 x4 (3072, 196), S4 x2 (6144, 49); d_state 16, K = G = 4, fp32, delta_softplus, D and delta_bias present, z=None
(exactly what code/networks/mamba_sys.py:420-426 passes).  One "step" = forward + backward of all 14.

  value     algorithmic GB/s (SURVEY.md section 8d byte formula) with inputs resident in HBM, kernels launched through
            the C ABI with caller-allocated outputs; CUDA events; max over ranks.
  e2e       the same metric through the public op (mamba_ssm.ops.selective_scan_interface.selective_scan_fn +
            autograd backward) starting from pinned HOST buffers, H2D and D2H copies inside the timed region.
  roofline  dominant kernel (the backward kernel at stage S1): algorithmic bytes / mean event-timed duration
            inside the timed steps, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  oracle/ref_torch.py (PyTorch port of the reference's selective_scan_ref) fwd + autograd bwd on
            the host cores: one scan direction of config 1 (B=1, 192 of its 768 channel rows, L=3136; --cpu-full: all of it,
            ~100 s), once, plus the stage-3 / stage-4 shapes at B=1 (rank 0, N=1 only).
  reference_cuda  (rank 0, N=1, when oracle/_ref was built) the reference's OWN CUDA kernels rebuilt for sm_100a on the same
            resident tensors, per stage and for the whole 14-scan step, and MambaUnet bs24 inference / training img/s with
            those kernels + the ATen chain of forward_corev0 (oracle/ref_model.py) -- BASELINE.md rows B2 / B4 / B5 / B6.

N > 1 (torchrun): every rank runs the same batch-24 workload on its own GPU (weak scaling: the scan shards over the
image batch only); the parameter gradients of the scan (dA, dD, ddelta_bias) are all-reduced over NCCL per call, overlapped
with the following scans and waited for at the end of the step, as DDP would.  `--impl reference` times the reference's CPU path (the port) instead; see the tier contract.
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "mamba-unet_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "selective_scan_fwd_bwd_hbm_gbps"
UNIT = "GB/s"
N_STATE, K_DIR = 16, 4
# (name, d_inner, L, SS2D calls per model forward)
STAGES = [("S1", 192, 3136, 4), ("S2", 384, 784, 4), ("S3", 768, 196, 4), ("S4", 1536, 49, 2)]
DT_RANK = {"S1": 6, "S2": 12, "S3": 24, "S4": 48}   # ceil(d_model / 16), d_model = d_inner / 2 (mamba_sys.py:271-296)
BATCH = 24


def bytes_fwd(b, kd, L, g=K_DIR, n=N_STATE):
    return 4 * (3 * b * kd * L + 2 * b * g * n * L + kd * (n + 2))


def bytes_bwd(b, kd, L, g=K_DIR, n=N_STATE):
    return 4 * (5 * b * kd * L + 4 * b * g * n * L + 2 * kd * (n + 2))


def workload_bytes(batch):
    return sum(c * (bytes_fwd(batch, K_DIR * d, L) + bytes_bwd(batch, K_DIR * d, L)) for _, d, L, c in STAGES)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  Started BEFORE the warm-up (nvidia-smi needs a few hundred ms to come
    up, more than a short timed region lasts); summary() keeps the samples that arrived between begin() and end()."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0, self.t1 = index, None, [], None, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lines = list(self.lines)
        if self.t0 is not None and self.t1 is not None:   # samples taken under the timed load (a line arrives up to one period late)
            inside = [ln for ts, ln in lines if self.t0 <= ts <= self.t1 + 0.06]
            lines = inside if inside else [ln for ts, ln in lines if self.t0 - 0.25 <= ts <= self.t1 + 0.25]
        else:
            lines = [ln for _, ln in lines]
        for ln in lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [s for s in sm if s >= 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(threads=None):
    """oracle/ref_torch.py (port of the reference selective_scan_ref, CPU) fwd + autograd bwd.  Returns one(dim, L, batch) ->
    seconds for a (batch, dim, L) call with K = 4 groups and d_state 16."""
    import torch
    from oracle.ref_torch import selective_scan_ref_torch

    if threads:
        torch.set_num_threads(threads)

    def one(dim, L=3136, batch=1, seed=0):
        g = torch.Generator().manual_seed(seed)
        u = torch.randn(batch, dim, L, generator=g).requires_grad_()
        dt = (0.5 * torch.rand(batch, dim, L, generator=g)).requires_grad_()
        A = (-0.5 * torch.rand(dim, N_STATE, generator=g)).requires_grad_()
        Bm = torch.randn(batch, K_DIR, N_STATE, L, generator=g).requires_grad_()
        Cm = torch.randn(batch, K_DIR, N_STATE, L, generator=g).requires_grad_()
        D = torch.randn(dim, generator=g).requires_grad_()
        bias = (0.5 * torch.rand(dim, generator=g)).requires_grad_()
        dout = torch.randn(batch, dim, L, generator=g)
        t0 = time.perf_counter()
        out = selective_scan_ref_torch(u, dt, A, Bm, Cm, D, None, bias, True)
        out.backward(dout)
        return time.perf_counter() - t0

    one(4)  # first call pays thread-pool / allocator warm-up
    return one


REF_SAMPLE_DIM = 64   # --impl reference: rows of config 1 per step (the port's cost per row is flat above ~32 rows)


def run_reference_impl(args):
    """--impl reference: the reference's CPU path (port) on the host cores; rank 0 only."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    one = cpu_reference_leg()
    dim, L = REF_SAMPLE_DIM, 3136
    for _ in range(args.warmup):
        one(dim)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one(dim)
    dt = (time.perf_counter() - t0) / args.steps
    nbytes = bytes_fwd(1, dim, L) + bytes_bwd(1, dim, L)
    val = nbytes / dt / 1e9
    sample = f"first {dim} of 768 channel rows of config 1 (B=1, K=4, d_state=16, L={L}), fwd + autograd bwd per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "MambaUnet-tiny 224x224 bs24 training step: 14 SS2D selective scans fwd+bwd "
                               "(reference arm: bounded CPU sample of config 1)", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
def reference_cuda_scan_leg(bufs, stages, batch, peak):
    """The reference's own selective_scan_cuda (oracle/_ref: unmodified sources rebuilt for sm_100a) on the SAME resident tensors,
    called as its autograd wrapper calls it (selective_scan_interface.py:37, :62-65; it allocates and zeroes its own outputs).
    BASELINE.md rows B2 / B3.  Baseline only: nothing here is on the product path."""
    import torch
    from oracle import ref_cuda

    if not ref_cuda.available():
        return {"unavailable": "oracle/_ref/selective_scan_cuda.so is not built (python oracle/build_ref.py in the build container)"}

    def timeit(fn, warm=2, iters=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    res = {"kernels": "selective_scan_fwd_kernel / selective_scan_bwd_kernel (mamba_ssm csrc, -O3 --use_fast_math, sm_100a)", "per_stage": {}}
    step_ms, step_bytes = 0.0, 0
    for name, d_inner, L, calls in stages:
        t = bufs[name]
        kd = K_DIR * d_inner
        args = (t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"])
        x = ref_cuda.ref_fwd(*args, True)[1]
        f = timeit(lambda: ref_cuda.ref_fwd(*args, True))
        w = timeit(lambda: ref_cuda.ref_bwd(*args, t["dout"], x, None, True))
        bf, bb = bytes_fwd(batch, kd, L), bytes_bwd(batch, kd, L)
        res["per_stage"][name] = {"fwd_ms": round(f, 4), "bwd_ms": round(w, 4), "fwdbwd_gbps": round((bf + bb) / (f + w) / 1e6, 1),
                                  "fwdbwd_frac": round((bf + bb) / (f + w) / 1e6 / peak, 4)}
        step_ms += calls * (f + w)
        step_bytes += calls * (bf + bb)
        del x
    res["ms_per_step"] = round(step_ms, 3)
    res["value"] = round(step_bytes / step_ms / 1e6, 2)
    res["unit"] = UNIT
    res["frac_of_hbm_peak"] = round(step_bytes / step_ms / 1e6 / peak, 4)
    return res


# ----------------------------------------------------------------------------------------------------------------
def model_leg(world, rank, dev, steps):
    """MambaUnet (vmamba_tiny, 19.1 M parameters, random init) on synthetic 1x224x224 slices, 4 classes -- BASELINE configs 2-4.
    The model is the from-scratch caller context (selscan_b200/vssm.py); every SS2D block runs the sm_100a scan.

    Arms (keys of the result):
      *                      default path: x_proj / dt_proj on the tcgen05 tensor cores (3xTF32, ss2d.TC_PROJ), nn.Linear on cuBLAS fp32
      *_cublas_proj          x_proj / dt_proj on cuBLAS fp32 as well (round-1 default)
      *_tc3xtf32             additionally every nn.Linear on the 3xTF32 kernel (layers OUTSIDE the hot path: separate key)
      *_cudagraph            the whole step replayed as one CUDA graph
    N > 1: DistributedDataParallel (static graph, gradients as bucket views, bucket size from SELSCAN_DDP_BUCKET_MB, default 25)."""
    import torch
    import torch.distributed as dist
    from selscan_b200 import ss2d, tcgemm
    from selscan_b200 import workloads as wl
    from selscan_b200.vssm import DiceLoss, MambaUnet

    torch.manual_seed(1337 + rank)
    res = {}
    dice = DiceLoss(4)
    bucket_mb = float(os.environ.get("SELSCAN_DDP_BUCKET_MB", "25"))

    def timed(fn, n, warm=3):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def ddp(m):
        if world == 1:
            return m
        return torch.nn.parallel.DistributedDataParallel(m, device_ids=[dev.index], gradient_as_bucket_view=True, static_graph=True,
                                                         bucket_cap_mb=bucket_mb)

    def entry(ms, imgs):
        return {"ms_per_step": round(ms, 3), "img_per_s": round(world * imgs / ms * 1e3, 1)}

    n = max(2, min(steps, 5))
    x24 = torch.rand(24, 1, 224, 224, device=dev)
    y24 = torch.randint(0, 4, (24, 224, 224), device=dev)
    x16, y16 = x24[:16], y24[:16]
    cw = wl.consistency_weight(3000)

    def graphed(key, fn, *inputs, imgs):
        try:
            g = wl.GraphedStep(fn, *inputs, warmup=11 if world > 1 else 3)   # DDP needs its first iterations outside the capture
            res[key] = entry(timed(lambda: g(*inputs), n), imgs)
            del g
        except Exception as e:  # noqa: BLE001 -- reported, not hidden
            res[key] = {"unavailable": repr(e)[:200]}
            torch.cuda.synchronize()

    tc_default = ss2d.TC_PROJ
    # DDP is constructed (and later warmed up / captured) on ONE side stream: whole-step CUDA-graph capture of a DDP model needs
    # its reducer and AccumulateGrad nodes created on the capturing stream.  Every eager leg runs first; the graph legs run last so
    # that a failed capture cannot disturb an eager number.
    cap_stream = torch.cuda.Stream(dev) if world > 1 else None

    def ddp_on_stream(m, stream=None):
        """DDP wrapper constructed on `stream` (default: the capture stream): its reducer and AccumulateGrad nodes belong to the
        stream the model's forwards will run on."""
        if world == 1:
            return m
        st = cap_stream if stream is None else stream
        st.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(st):
            w = ddp(m)
        torch.cuda.current_stream(dev).wait_stream(st)
        return w

    try:
        # ---- default arm: inference, supervised training ----
        model = MambaUnet(num_classes=4).to(dev)
        model.eval()
        with torch.no_grad():
            res["infer_bs24"] = entry(timed(lambda: model(x24), n), 24)
            if world == 1:
                graphed("infer_bs24_cudagraph", lambda x: model(x), x24, imgs=24)
        model.train()
        net = ddp(model)
        opt = wl.make_sgd(net)
        res["train_supervised_bs24"] = entry(timed(lambda: wl.supervised_step(net, opt, dice, x24, y24), n), 24)
        # ---- x_proj / dt_proj back on cuBLAS fp32 (same model, same optimizer state) ----
        ss2d.TC_PROJ = False
        res["train_supervised_bs24_cublas_proj"] = entry(timed(lambda: wl.supervised_step(net, opt, dice, x24, y24), n), 24)
        ss2d.TC_PROJ = tc_default
        del net, opt, model
        # ---- every nn.Linear on the 3xTF32 tensor-core GEMM too (outside the hot path: own keys) ----
        try:
            model = MambaUnet(num_classes=4).to(dev).train()
            tcgemm.patch_linears(model)
            ss2d.TC_PROJ = True
            net = ddp(model)
            opt = wl.make_sgd(net)
            res["train_supervised_bs24_tc3xtf32"] = entry(timed(lambda: wl.supervised_step(net, opt, dice, x24, y24), n), 24)
            del net, opt, model
        except Exception as e:  # noqa: BLE001
            res["train_supervised_bs24_tc3xtf32"] = {"unavailable": repr(e)[:200]}
        finally:
            ss2d.TC_PROJ = tc_default
        # ---- BASELINE config 4: dual-network semi-supervised step, 16 images per GPU (8 labeled) ----
        m1, m2 = ddp(MambaUnet(num_classes=4).to(dev).train()), ddp(MambaUnet(num_classes=4).to(dev).train())
        o1, o2 = wl.make_sgd(m1), wl.make_sgd(m2)
        res["train_semi_dual_bs16"] = entry(timed(lambda: wl.semi_step(m1, m2, o1, o2, dice, x16, y16, 8, cw), n), 16)
        # the same step with the second network on a second stream: at 16 images per network most kernels under-fill a B200 (stage-1
        # scan: 192 CTAs on 296 slots), the two networks are independent until the losses
        del m1, m2, o1, o2
        torch.cuda.empty_cache()
        try:
            side = torch.cuda.Stream(dev)
            m1 = ddp(MambaUnet(num_classes=4).to(dev).train())
            m2 = ddp_on_stream(MambaUnet(num_classes=4).to(dev).train(), side)      # (the second network lives on the second stream)
            o1, o2 = wl.make_sgd(m1), wl.make_sgd(m2)
            res["train_semi_dual_bs16_2streams"] = entry(timed(lambda: wl.semi_step(m1, m2, o1, o2, dice, x16, y16, 8, cw, side), n,
                                                               warm=8 if world > 1 else 3), 16)    # (DDP's static graph settles in its first iterations)
            del m1, m2, o1, o2
        except Exception as e:  # noqa: BLE001
            res["train_semi_dual_bs16_2streams"] = {"unavailable": repr(e)[:200]}
        torch.cuda.empty_cache()

        # ---- the same steps replayed as ONE CUDA graph each (fresh models: under DDP the wrapper must be built on the capture stream) ----
        def sup_factory(tc_linears):
            def make():
                m = MambaUnet(num_classes=4).to(dev).train()
                if tc_linears:
                    tcgemm.patch_linears(m)
                w = ddp_on_stream(m)
                o = wl.make_sgd(w)
                return lambda x, y: wl.supervised_step(w, o, dice, x, y)
            return make

        def semi_factory(two_streams=False):
            def make():
                side = torch.cuda.Stream(dev) if two_streams else None
                a, b = ddp_on_stream(MambaUnet(num_classes=4).to(dev).train()), ddp_on_stream(MambaUnet(num_classes=4).to(dev).train(), side)
                oa, ob = wl.make_sgd(a), wl.make_sgd(b)
                return lambda x, y: wl.semi_step(a, b, oa, ob, dice, x, y, 8, cw, side)
            return make

        for key, make, inputs, imgs, tc in (("train_supervised_bs24_cudagraph", sup_factory(False), (x24, y24), 24, tc_default),
                                            ("train_supervised_bs24_tc3xtf32_cudagraph", sup_factory(True), (x24, y24), 24, True),
                                            ("train_semi_dual_bs16_cudagraph", semi_factory(), (x16, y16), 16, tc_default),
                                            ("train_semi_dual_bs16_2streams_cudagraph", semi_factory(True), (x16, y16), 16, tc_default)):
            ss2d.TC_PROJ = tc
            try:
                fn = make()
                g = wl.GraphedStep(fn, *inputs, warmup=11 if world > 1 else 3, stream=cap_stream)
                res[key] = entry(timed(lambda: g(*inputs), n), imgs)
                del g, fn
            except Exception as e:  # noqa: BLE001 -- reported, not hidden
                res[key] = {"unavailable": repr(e)[:200]}
                try:
                    torch.cuda.synchronize()
                except Exception:  # noqa: BLE001
                    pass
                if world > 1:
                    break               # a failed capture under DDP leaves the collectives of the ranks out of step: stop here
            torch.cuda.empty_cache()
    finally:
        ss2d.TC_PROJ = tc_default
    res["note"] = ("img/s is the whole-job aggregate over %d GPU(s); per-GPU batch fixed (weak scaling); DDP (static_graph, "
                   "gradient_as_bucket_view, bucket_cap_mb=%g) gradient all-reduce over NCCL when n_gpus > 1 (19.1 M fp32 gradients per "
                   "model); default arm: ss2d.TC_PROJ=%s; *_2streams: the second network of the dual-network step runs on a second CUDA "
                   "stream (workloads.semi_step(side=...)), same arithmetic" % (world, bucket_mb, tc_default))
    return res


def reference_cuda_model_leg(dev, steps):
    """MambaUnet bs24 with the reference's kernels + the ATen chain of forward_corev0 / SS2D.forward + torch LayerNorm + the
    reference DiceLoss (oracle/ref_model.py) -- BASELINE.md rows B4 / B5 / B6 on this GPU.  Rank 0, N = 1 only."""
    import torch
    from oracle import ref_cuda, ref_model
    from selscan_b200 import workloads as wl
    from selscan_b200.vssm import MambaUnet

    if not ref_cuda.available():
        return {"unavailable": "oracle/_ref/selective_scan_cuda.so is not built"}
    torch.manual_seed(1337)
    n = max(2, min(steps, 3))

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    res = {}
    model = ref_model.to_reference_path(MambaUnet(num_classes=4).to(dev))
    x24 = torch.rand(24, 1, 224, 224, device=dev)
    y24 = torch.randint(0, 4, (24, 224, 224), device=dev)
    model.eval()
    with torch.no_grad():
        ms = timed(lambda: model(x24))
    res["infer_bs24"] = {"ms_per_step": round(ms, 3), "img_per_s": round(24 / ms * 1e3, 1)}
    model.train()
    opt = wl.make_sgd(model)
    dice = ref_model.RefDiceLoss(4)
    ms = timed(lambda: wl.supervised_step(model, opt, dice, x24, y24))
    res["train_supervised_bs24"] = {"ms_per_step": round(ms, 3), "img_per_s": round(24 / ms * 1e3, 1)}
    del model, opt
    m1 = ref_model.to_reference_path(MambaUnet(num_classes=4).to(dev).train())
    m2 = ref_model.to_reference_path(MambaUnet(num_classes=4).to(dev).train())
    o1, o2 = wl.make_sgd(m1), wl.make_sgd(m2)
    ms = timed(lambda: wl.semi_step(m1, m2, o1, o2, dice, x24[:16], y24[:16], 8, wl.consistency_weight(3000)))
    res["train_semi_dual_bs16"] = {"ms_per_step": round(ms, 3), "img_per_s": round(16 / ms * 1e3, 1)}
    res["note"] = "same architecture / initialisation as the product model legs; reference scan kernels, ATen forward_corev0 chain, torch LayerNorm, DiceLoss with .item()"
    return res


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="images per GPU (BASELINE config: 24)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-full", action="store_true", help="CPU baseline on config 1 in full (~100 s) instead of one direction of it")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-model", action="store_true", help="skip the MambaUnet img/s legs")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the reference-CUDA-kernel baseline block")
    ap.add_argument("--stages", default="", help="comma list to restrict (profiling only), e.g. S1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference_impl(args)
        return

    import torch
    import torch.distributed as dist
    from selscan_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    batch = args.batch
    stages = [s for s in STAGES if not args.stages or s[0] in args.stages.split(",")]

    clk = ClockSampler(local)
    clk.__enter__()     # nvidia-smi comes up while the inputs are generated; begin() / end() bracket the timed region

    # ---- resident inputs / preallocated outputs per stage (model-like distribution M) ----
    gen = torch.Generator(device=dev).manual_seed(1337 + rank)
    bufs = {}
    for name, d_inner, L, calls in stages:
        kd = K_DIR * d_inner
        t = {}
        def rows(fill=None):
            # (batch, kd, L) with 16-byte aligned rows -- the layout forward_core_b200 / SelectiveScanFn allocate (L = 49 -> pitch 52)
            r = ops.empty_rows(batch, kd, L, dev)
            if fill is not None:
                r.copy_(fill)
            return r

        t["u"] = rows(torch.randn(batch, kd, L, device=dev, generator=gen))
        t["delta"] = rows(0.5 * torch.randn(batch, kd, L, device=dev, generator=gen))
        t["A"] = -torch.arange(1, N_STATE + 1, device=dev, dtype=torch.float32).repeat(kd, 1).contiguous()
        # B / C exactly as SS2D hands them over: strided views of x_dbl = x_proj(xs), (batch, K, R + 2N, L) (ss2d.py: the batched
        # GEMM writes that layout; group stride (R + 2N) * L, unit position stride) -- no (N, L)-contiguous copy exists in the model
        R = DT_RANK[name]
        t["x_dbl"] = torch.randn(batch, K_DIR, R + 2 * N_STATE, L, device=dev, generator=gen)
        t["B"] = t["x_dbl"][:, :, R:R + N_STATE]
        t["C"] = t["x_dbl"][:, :, R + N_STATE:]
        t["D"] = torch.ones(kd, device=dev)
        dtv = torch.exp(torch.rand(kd, device=dev, generator=gen) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
        t["bias"] = dtv + torch.log(-torch.expm1(-dtv))
        t["dout"] = rows(torch.randn(batch, kd, L, device=dev, generator=gen))
        t["out"] = rows()
        t["ckpt"] = torch.empty(max(ops.ckpt_elems(batch, kd, L, N_STATE), 4), device=dev)
        t["du"] = rows()
        t["ddelta"] = rows()
        nbc = batch * K_DIR * N_STATE * L
        t["flat"] = torch.zeros(2 * nbc + kd * N_STATE + 2 * kd, device=dev)   # dB | dC | dA | dD | dbias: one memset per call
        t["dB"] = t["flat"][:nbc].view(batch, K_DIR, N_STATE, L)
        t["dC"] = t["flat"][nbc:2 * nbc].view(batch, K_DIR, N_STATE, L)
        t["dA"] = t["flat"][2 * nbc:2 * nbc + kd * N_STATE].view(kd, N_STATE)
        t["dD"] = t["flat"][2 * nbc + kd * N_STATE:2 * nbc + kd * N_STATE + kd]
        t["dbias"] = t["flat"][2 * nbc + kd * N_STATE + kd:]
        bufs[name] = t

    ev_cache = []
    pending = []
    if world > 1:
        for name, d_inner, L, calls in stages:
            t = bufs[name]
            n_par = t["flat"].numel() - t["dB"].numel() - t["dC"].numel()
            t["red"] = [torch.empty(n_par, device=dev) for _ in range(calls)]

    def step(record=None):
        """One pass over the workload through the C ABI.  record: list to append (stage, ev0, ev1, ev2) to."""
        for name, d_inner, L, calls in stages:
            t = bufs[name]
            for ci in range(calls):
                t["flat"].zero_()
                if record is not None:
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record()
                ops.launch_fwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], True, t["out"],
                               None, None, t["ckpt"])
                if record is not None:
                    e1.record()
                ops.launch_bwd(t["u"], t["delta"], t["A"], t["B"], t["C"], t["D"], None, t["bias"], t["dout"], None,
                               t["ckpt"], True, t["du"], t["ddelta"], t["dA"], t["dB"], t["dC"], t["dD"], None, t["dbias"])
                if record is not None:
                    e2.record()
                    record.append((name, e0, e1, e2))
                if world > 1:
                    # what DDP would reduce for this op's parameters: dA | dD | dbias are contiguous -> one bucket per call,
                    # all-reduced on NCCL's stream while the next scans run (DDP overlaps its buckets with the rest of the
                    # backward in the same way); the step ends by waiting for every bucket
                    red = t["red"][ci]
                    red.copy_(t["flat"][t["dB"].numel() + t["dC"].numel():])
                    pending.append(dist.all_reduce(red, async_op=True))
        for w in pending:
            w.wait()
        pending.clear()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rec = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        for _ in range(args.warmup):
            step()
        barrier()
        clk.begin()
        t_wall0 = time.perf_counter()
        start.record()
        for _ in range(args.steps):
            step(rec)
        stop.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clk.end()
    finally:
        clk.__exit__(None, None, None)
    ms_total = start.elapsed_time(stop)
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_step = ms_total / args.steps
    total_bytes = world * sum(c * (bytes_fwd(batch, K_DIR * d, L) + bytes_bwd(batch, K_DIR * d, L)) for _, d, L, c in stages)
    value = total_bytes / (ms_step * 1e-3) / 1e9
    clocks = clk.summary()

    # ---- per-kernel table from the events recorded inside the timed steps -------------------------------------
    peak, peak_src = measured_peak()
    kern = {}
    for name, e0, e1, e2 in rec:
        kern.setdefault(name, {"fwd": [], "bwd": []})
        kern[name]["fwd"].append(e0.elapsed_time(e1))
        kern[name]["bwd"].append(e1.elapsed_time(e2))
    table = {}
    for name, d_inner, L, calls in stages:
        kd = K_DIR * d_inner
        f_ms, b_ms = statistics.mean(kern[name]["fwd"]), statistics.mean(kern[name]["bwd"])
        bf, bb = bytes_fwd(batch, kd, L), bytes_bwd(batch, kd, L)
        table[name] = {"fwd_ms": round(f_ms, 4), "bwd_ms": round(b_ms, 4),
                       "fwd_gbps": round(bf / f_ms / 1e6, 1), "bwd_gbps": round(bb / b_ms / 1e6, 1),
                       "fwdbwd_gbps": round((bf + bb) / (f_ms + b_ms) / 1e6, 1),
                       "fwdbwd_frac": round((bf + bb) / (f_ms + b_ms) / 1e6 / peak, 4)}
    dom = max(table, key=lambda n: table[n]["bwd_ms"] * dict((s[0], s[3]) for s in stages)[n])
    dom_stage = [s for s in stages if s[0] == dom][0]
    dom_bytes = bytes_bwd(batch, K_DIR * dom_stage[1], dom_stage[2])
    achieved = dom_bytes / table[dom]["bwd_ms"] / 1e6
    traffic, traffic_src = None, None
    try:  # DRAM bytes per launch of the same kernel from the committed ncu capture (profiles/), batch 24 only
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as f:
            if batch == BATCH:
                traffic = json.load(f)["stages"][dom]["bwd"]["dram_bytes"]
                traffic_src = "profiles/r02_ncu_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, per launch)"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": f"{ops.bwd_kernel_name()} @ {dom} (B={batch}, KD={K_DIR * dom_stage[1]}, L={dom_stage[2]})",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": table[dom]["bwd_ms"],
                "share_of_step": round(dict((s[0], s[3]) for s in stages)[dom] * table[dom]["bwd_ms"] /
                                       sum(s[3] * (table[s[0]]["fwd_ms"] + table[s[0]]["bwd_ms"]) for s in stages), 4),
                "note": "not HBM-bound: both warp roles of this kernel run dependent instruction chains ~85-90% of the time "
                        "(shared-memory pipe 71% busy, issue slots 54%, MUFU 51% at once: profiles/r02_ncu_scan_summary.txt; what each "
                        "cost is worth: profiles/r02_bwd_whatif.json); see DESIGN.md section 4"}

    # ---- e2e: public op from pinned host buffers ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        from mamba_ssm.ops.selective_scan_interface import selective_scan_fn

        e_steps = max(1, min(args.steps, 3))
        host = {}
        h2d = d2h = 0
        for name, d_inner, L, calls in stages:
            t = bufs[name]
            hin = {k: torch.empty(tuple(t[k].shape), dtype=torch.float32, pin_memory=True).copy_(t[k]) for k in ("u", "delta", "B", "C", "dout")}   # B / C travel as dense (N, L) blocks
            hout = {k: torch.empty(tuple(t[k].shape), dtype=torch.float32, pin_memory=True) for k in ("out", "du", "ddelta", "dB", "dC", "dA", "dD", "dbias")}
            host[name] = (hin, hout)
            h2d += calls * sum(v.numel() * 4 for v in hin.values())
            d2h += calls * sum(v.numel() * 4 for v in hout.values())

        # Host-side pipeline: the next call's inputs go up on a copy stream while the current call computes, and the previous
        # call's results come down on a second copy stream (PCIe is full duplex).  All three are inside the timed region.
        main = torch.cuda.current_stream(dev)
        h2d_s, d2h_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        call_list = [name for name, _, _, calls in stages for _ in range(calls)]

        # device-side input staging: two reusable buffer sets per stage shape (no allocator traffic inside the timed region)
        ring = {name: [{k: torch.empty_like(bufs[name][k], memory_format=torch.contiguous_format) if k in ("B", "C") else torch.empty_like(bufs[name][k])
                        for k in ("u", "delta", "B", "C", "dout")} for _ in range(2)]
                for name, _, _, _ in stages}
        ring_done = {name: [None, None] for name, _, _, _ in stages}
        ring_next = {name: 0 for name, _, _, _ in stages}

        def upload(name):
            slot = ring_next[name]
            ring_next[name] = slot ^ 1
            with torch.cuda.stream(h2d_s):
                if ring_done[name][slot] is not None:
                    h2d_s.wait_event(ring_done[name][slot])      # the call that last computed from this set has finished
                dv = ring[name][slot]
                for k, v in host[name][0].items():
                    dv[k].copy_(v, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(h2d_s)
            return dv, ev, slot

        def e2e_step():
            h2d_s.wait_stream(main)
            nxt = upload(call_list[0])
            for i, name in enumerate(call_list):
                t = bufs[name]
                hout = host[name][1]
                dv, ev, slot = nxt
                if i + 1 < len(call_list):
                    nxt = upload(call_list[i + 1])
                main.wait_event(ev)
                u, dl, Bm, Cm = (dv[k].detach().requires_grad_() for k in ("u", "delta", "B", "C"))
                A, Dp, bias = (t[k].detach().clone().requires_grad_() for k in ("A", "D", "bias"))
                out = selective_scan_fn(u, dl, A, Bm, Cm, Dp, z=None, delta_bias=bias, delta_softplus=True,
                                        return_last_state=False)
                out.backward(dv["dout"])
                if world > 1:
                    for p in (A, Dp, bias):
                        dist.all_reduce(p.grad)
                done = torch.cuda.Event()
                done.record(main)
                ring_done[name][slot] = done
                d2h_s.wait_event(done)
                with torch.cuda.stream(d2h_s):
                    for k, src in (("out", out.detach()), ("du", u.grad), ("ddelta", dl.grad), ("dB", Bm.grad), ("dC", Cm.grad),
                                   ("dA", A.grad), ("dD", Dp.grad), ("dbias", bias.grad)):
                        hout[k].copy_(src, non_blocking=True)
                        src.record_stream(d2h_s)
            main.wait_stream(d2h_s)   # the step ends when its last result is on the host

        e2e_step()
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        s2, p2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(e_steps):
            e2e_step()
        p2.record()
        barrier()
        e_ms = max(s2.elapsed_time(p2), (time.perf_counter() - t0) * 1e3) / e_steps  # host-side copies count
        if world > 1:
            tt = torch.tensor([e_ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_ms = float(tt.item())
        e2e = {"value": round(total_bytes / (e_ms * 1e-3) / 1e9, 2), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": round(e_ms, 3), "steps": e_steps,
               "api": "mamba_ssm.ops.selective_scan_interface.selective_scan_fn + autograd backward, pinned host buffers; "
                      "H2D / compute / D2H pipelined on three streams"}
        del host

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        one = cpu_reference_leg()
        # BASELINE config 1 is the stage-1 SS2D of ONE 224x224 slice (B=1, K=4, 768 channel rows, L=3136).  In full the port needs
        # ~100 s on the box's 16 cores (profiles/r02_bench_n1_cpufull.json: 100.7 s, its autograd backward thrashes memory), so the
        # default run times ONE of its four directions (192 rows, same L) and --cpu-full runs all of it once.
        dim, L = (768, 3136) if args.cpu_full else (192, 3136)
        dt = one(dim, L)
        nb = bytes_fwd(1, dim, L) + bytes_bwd(1, dim, L)
        what = "config 1 in full (B=1, K=4, all 768 channel rows" if args.cpu_full else "one of the four scan directions of config 1 (B=1, 192 of its 768 channel rows"
        cpu = {"value": round(nb / dt / 1e9, 6), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "host_cpus": os.cpu_count(), "seconds": round(dt, 2),
               "sample": f"oracle/ref_torch.py (PyTorch port of selective_scan_ref) fwd + autograd bwd, {what}, d_state=16, L={L}), one pass",
               "full_config1_measured": {"seconds": 100.67, "gbps": 0.000815, "cores": 16, "source": "profiles/r02_bench_n1_cpufull.json"},
               "other_shapes": {}}
        for nm, d_inner, Ls, _ in STAGES[2:]:   # the short stages at B=1 (seconds each): a multi-point comparison
            ts = one(K_DIR * d_inner, Ls)
            nbs = bytes_fwd(1, K_DIR * d_inner, Ls) + bytes_bwd(1, K_DIR * d_inner, Ls)
            cpu["other_shapes"][nm] = {"dim": K_DIR * d_inner, "L": Ls, "seconds": round(ts, 3), "gbps": round(nbs / ts / 1e9, 6)}

    # ---- the reference's own CUDA kernels on the same tensors (rank 0, N = 1 only) -------------------------------
    ref_cuda_block = None
    if world == 1 and rank == 0 and not args.no_reference_cuda:
        try:
            ref_cuda_block = {"scan": reference_cuda_scan_leg(bufs, stages, batch, peak)}
        except Exception as e:  # noqa: BLE001 -- a baseline leg must not take the product's line down
            ref_cuda_block = {"scan": {"unavailable": repr(e)[:300]}}
            torch.cuda.synchronize()

    model = None
    if not args.no_model and not args.stages:
        for t in bufs.values():
            t.clear()
        torch.cuda.empty_cache()
        model = model_leg(world, rank, dev, args.steps)
        if ref_cuda_block is not None:
            torch.cuda.empty_cache()
            try:
                ref_cuda_block["mambaunet"] = reference_cuda_model_leg(dev, args.steps)
            except Exception as e:  # noqa: BLE001
                ref_cuda_block["mambaunet"] = {"unavailable": repr(e)[:300]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MambaUnet-tiny 224x224 bs%d training step: 14 SS2D selective scans fwd+bwd "
                                   "(S1 4x[KD768,L3136], S2 4x[1536,784], S3 4x[3072,196], S4 2x[6144,49]; N=16, G=4)" % batch,
                       "batch_per_gpu": batch, "stages": [s[0] for s in stages],
                       "l2": "inputs larger than L2: every call streams >= 234 MB (no explicit flush)",
                       "bc_layout": "B / C are strided views of x_dbl (batch, K, R+2N, L) as SS2D passes them; dB / dC dense",
                       "algorithmic_bytes_per_step": total_bytes, "parallelism": f"dp{world} (batch-sharded replicas)"},
            "frac_of_hbm_peak": round(value / world / peak, 4),
            "roofline": roofline, "per_stage": table, "mambaunet": model, "reference_cuda": ref_cuda_block, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": args.steps * sum(2 * s[3] for s in stages), "wall_s_timed_region": round(t_wall, 3),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
