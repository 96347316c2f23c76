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
            the host cores, on a bounded sample of config 1 (rank 0, N=1 only).

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
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

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
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
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
    """Time oracle/ref_torch.py (port of the reference selective_scan_ref, CPU) fwd + autograd bwd on a bounded
    sample of config 1 (B=1, K=4, d_state=16, L=3136): the first 64 of its 768 channel rows (a few seconds a pass)."""
    import torch
    from oracle.ref_torch import selective_scan_ref_torch

    if threads:
        torch.set_num_threads(threads)
    L = 3136

    def one(dim, seed=0):
        g = torch.Generator().manual_seed(seed)
        u = torch.randn(1, dim, L, generator=g).requires_grad_()
        dt = (0.5 * torch.rand(1, dim, L, generator=g)).requires_grad_()
        A = (-0.5 * torch.rand(dim, N_STATE, generator=g)).requires_grad_()
        Bm = torch.randn(1, K_DIR, N_STATE, L, generator=g).requires_grad_()
        Cm = torch.randn(1, K_DIR, N_STATE, L, generator=g).requires_grad_()
        D = torch.randn(dim, generator=g).requires_grad_()
        bias = (0.5 * torch.rand(dim, generator=g)).requires_grad_()
        dout = torch.randn(1, dim, L, generator=g)
        t0 = time.perf_counter()
        out = selective_scan_ref_torch(u, dt, A, Bm, Cm, D, None, bias, True)
        out.backward(dout)
        return time.perf_counter() - t0

    one(4)  # first call pays thread-pool / allocator warm-up
    dim = 64  # fixed sample: ~5 s per pass on 8 cores (the port's cost per row is flat above ~32 rows)
    return dim, L, one


def run_reference_impl(args):
    """--impl reference: the reference's CPU path (port) on the host cores; rank 0 only."""
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dim, L, one = cpu_reference_leg()
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
def model_leg(world, rank, dev, steps):
    """MambaUnet (vmamba_tiny, 19.1 M parameters, random init) on synthetic 1x224x224 slices, 4 classes -- BASELINE configs 2-4.
    The model is the from-scratch caller context (selscan_b200/vssm.py); every SS2D block runs the sm_100a scan."""
    import torch
    import torch.distributed as dist
    from selscan_b200 import workloads as wl
    from selscan_b200.vssm import DiceLoss, MambaUnet

    torch.manual_seed(1337 + rank)
    res = {}
    dice = DiceLoss(4)

    def timed(fn, n):
        for _ in range(3):
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

    n = max(2, min(steps, 5))
    model = MambaUnet(num_classes=4).to(dev)
    x24 = torch.rand(24, 1, 224, 224, device=dev)
    y24 = torch.randint(0, 4, (24, 224, 224), device=dev)
    model.eval()
    with torch.no_grad():
        ms = timed(lambda: model(x24), n)
    res["infer_bs24"] = {"ms_per_step": round(ms, 3), "img_per_s": round(world * 24 / ms * 1e3, 1)}
    if world == 1:   # the same forward replayed as one CUDA graph (host launch overhead removed)
        try:
            with torch.no_grad():
                g = wl.GraphedStep(lambda x: model(x), x24)
                ms = timed(lambda: g(x24), n)
            res["infer_bs24_cudagraph"] = {"ms_per_step": round(ms, 3), "img_per_s": round(24 / ms * 1e3, 1)}
            del g
        except Exception as e:  # noqa: BLE001 -- reported, not hidden
            res["infer_bs24_cudagraph"] = {"unavailable": repr(e)[:200]}
    model.train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], gradient_as_bucket_view=True) if world > 1 else model
    opt = wl.make_sgd(net)
    ms = timed(lambda: wl.supervised_step(net, opt, dice, x24, y24), n)
    res["train_supervised_bs24"] = {"ms_per_step": round(ms, 3), "img_per_s": round(world * 24 / ms * 1e3, 1)}
    if world == 1:
        try:
            g = wl.GraphedStep(lambda x, y: wl.supervised_step(net, opt, dice, x, y), x24, y24)
            ms = timed(lambda: g(x24, y24), n)
            res["train_supervised_bs24_cudagraph"] = {"ms_per_step": round(ms, 3), "img_per_s": round(24 / ms * 1e3, 1)}
            del g
        except Exception as e:  # noqa: BLE001
            res["train_supervised_bs24_cudagraph"] = {"unavailable": repr(e)[:200]}
    from selscan_b200 import ss2d
    if world == 1:   # opt-in legs: Linear layers, x_proj, dt_proj and d(xs) on the tcgen05 tensor cores with the 3xTF32 split
        # (fp32-level accuracy, ~2x the rounding error of cuBLAS fp32; NOT the reference's cuBLAS arithmetic, hence separate keys)
        try:
            from selscan_b200 import ss2d, tcgemm
            tcgemm.patch_linears(model)
            ss2d.TC_PROJ = True                    # x_proj / dt_proj too
            ms = timed(lambda: wl.supervised_step(net, opt, dice, x24, y24), n)
            res["train_supervised_bs24_tc3xtf32"] = {"ms_per_step": round(ms, 3), "img_per_s": round(24 / ms * 1e3, 1)}
            g = wl.GraphedStep(lambda x, y: wl.supervised_step(net, opt, dice, x, y), x24, y24)
            ms = timed(lambda: g(x24, y24), n)
            res["train_supervised_bs24_tc3xtf32_cudagraph"] = {"ms_per_step": round(ms, 3), "img_per_s": round(24 / ms * 1e3, 1)}
            del g
        except Exception as e:  # noqa: BLE001
            res["train_supervised_bs24_tc3xtf32"] = {"unavailable": repr(e)[:200]}
        finally:
            ss2d.TC_PROJ = False
    del net, opt, model
    m1, m2 = MambaUnet(num_classes=4).to(dev).train(), MambaUnet(num_classes=4).to(dev).train()
    if world > 1:
        m1 = torch.nn.parallel.DistributedDataParallel(m1, device_ids=[dev.index], gradient_as_bucket_view=True)
        m2 = torch.nn.parallel.DistributedDataParallel(m2, device_ids=[dev.index], gradient_as_bucket_view=True)
    o1, o2 = wl.make_sgd(m1), wl.make_sgd(m2)
    x16, y16 = x24[:16], y24[:16]
    cw = wl.consistency_weight(3000)
    ms = timed(lambda: wl.semi_step(m1, m2, o1, o2, dice, x16, y16, 8, cw), n)
    res["train_semi_dual_bs16"] = {"ms_per_step": round(ms, 3), "img_per_s": round(world * 16 / ms * 1e3, 1)}
    res["note"] = ("img/s is the whole-job aggregate over %d GPU(s); per-GPU batch fixed (weak scaling); DDP gradient all-reduce over "
                   "NCCL when n_gpus > 1 (19.1 M fp32 gradients per model)" % world)
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
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-model", action="store_true", help="skip the MambaUnet img/s legs")
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
        t["B"] = torch.randn(batch, K_DIR, N_STATE, L, device=dev, generator=gen)
        t["C"] = torch.randn(batch, K_DIR, N_STATE, L, device=dev, generator=gen)
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

    for _ in range(args.warmup):
        step()
    barrier()
    rec = []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        t_wall0 = time.perf_counter()
        start.record()
        for _ in range(args.steps):
            step(rec)
        stop.record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
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
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
            if batch == BATCH:
                traffic = json.load(f)["stages"][dom]["bwd"]["dram_bytes"]
                traffic_src = "profiles/r01_ncu_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, per launch)"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": f"{ops.bwd_kernel_name()} @ {dom} (B={batch}, KD={K_DIR * dom_stage[1]}, L={dom_stage[2]})",
                "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": table[dom]["bwd_ms"],
                "share_of_step": round(dict((s[0], s[3]) for s in stages)[dom] * table[dom]["bwd_ms"] /
                                       sum(s[3] * (table[s[0]]["fwd_ms"] + table[s[0]]["bwd_ms"]) for s in stages), 4),
                "note": "not HBM-bound: SM-side limits bound this kernel (shared-memory/shuffle pipe ~79% busy, issue slots 54%, MUFU 48% "
                        "at once, profiles/r01_ncu_final_summary.txt); see DESIGN.md section 4"}

    # ---- e2e: public op from pinned host buffers ---------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        from mamba_ssm.ops.selective_scan_interface import selective_scan_fn

        e_steps = max(1, min(args.steps, 3))
        host = {}
        h2d = d2h = 0
        for name, d_inner, L, calls in stages:
            t = bufs[name]
            hin = {k: torch.empty(tuple(t[k].shape), dtype=torch.float32, pin_memory=True).copy_(t[k]) for k in ("u", "delta", "B", "C", "dout")}
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
        ring = {name: [{k: torch.empty_like(bufs[name][k]) for k in ("u", "delta", "B", "C", "dout")} for _ in range(2)]
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
        dim, L, one = cpu_reference_leg()
        dt = one(dim)
        nb = bytes_fwd(1, dim, L) + bytes_bwd(1, dim, L)
        cpu = {"value": round(nb / dt / 1e9, 6), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "host_cpus": os.cpu_count(), "seconds": round(dt, 2),
               "sample": f"oracle/ref_torch.py (PyTorch port of selective_scan_ref) fwd + autograd bwd, first {dim} of 768 "
                         f"channel rows of config 1 (B=1, K=4, d_state=16, L={L}), one pass"}

    model = None
    if not args.no_model and not args.stages:
        for t in bufs.values():
            t.clear()
        torch.cuda.empty_cache()
        model = model_leg(world, rank, dev, args.steps)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "MambaUnet-tiny 224x224 bs%d training step: 14 SS2D selective scans fwd+bwd "
                                   "(S1 4x[KD768,L3136], S2 4x[1536,784], S3 4x[3072,196], S4 2x[6144,49]; N=16, G=4)" % batch,
                       "batch_per_gpu": batch, "stages": [s[0] for s in stages],
                       "l2": "inputs larger than L2: every call streams >= 234 MB (no explicit flush)",
                       "algorithmic_bytes_per_step": total_bytes, "parallelism": f"dp{world} (batch-sharded replicas)"},
            "frac_of_hbm_peak": round(value / world / peak, 4),
            "roofline": roofline, "per_stage": table, "mambaunet": model, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": args.steps * sum(2 * s[3] for s in stages), "wall_s_timed_region": round(t_wall, 3),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
