"""N > 1 host logic on CPU (gloo, world_size 2): the path shards over the image batch only, so parameter gradients of the scan
(dA, dD, ddelta_bias) are the SUM over ranks of per-shard gradients, and everything else (out, du, ddelta, dB, dC) is rank-local.
Checked with the C oracle standing in for the kernels (no GPU here): shard -> per-rank backward -> all_reduce == full-batch backward."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import oracle as orc

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = 4
        inp = orc.make_inputs(batch, 8, 40, 16, 4, dist="M", seed=5)        # identical on every rank (seeded)
        lo, hi = rank * batch // world, (rank + 1) * batch // world        # this rank's images
        sh = {k: (v[lo:hi] if v is not None and v.ndim >= 3 else v) for k, v in inp.items()}
        out = orc.oracle_fwd(sh["u"], sh["delta"], sh["A"], sh["B"], sh["C"], sh["D"], None, sh["delta_bias"], True)
        g = orc.oracle_bwd(sh["u"], sh["delta"], sh["A"], sh["B"], sh["C"], sh["D"], None, sh["delta_bias"], sh["dout"], True)
        red = {}
        for k in ("dA", "dD", "ddelta_bias"):                                # what DDP all-reduces for this op
            t = torch.from_numpy(g[k].astype(np.float64))
            dist.all_reduce(t)
            red[k] = t.numpy()
        gathered = [None] * world
        dist.all_gather_object(gathered, {"out": out, "du": g["du"], "dB": g["dB"]})
        if rank == 0:
            full_out = orc.oracle_fwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None, inp["delta_bias"], True)
            full = orc.oracle_bwd(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], None, inp["delta_bias"],
                                  inp["dout"], True)
            for k in red:
                np.testing.assert_allclose(red[k], full[k], rtol=1e-5, atol=1e-5)
            np.testing.assert_array_equal(np.concatenate([x["out"] for x in gathered]), full_out)
            np.testing.assert_array_equal(np.concatenate([x["du"] for x in gathered]), full["du"])
            np.testing.assert_array_equal(np.concatenate([x["dB"] for x in gathered]), full["dB"])
            ret["ok"] = True
    finally:
        dist.destroy_process_group()


def test_batch_sharding_allreduce_gloo():
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert ret.get("ok") is True


def test_bench_byte_model_matches_survey():
    """The algorithmic-byte formula bench.py reports against (SURVEY.md section 8d worked numbers)."""
    sys.path.insert(0, ROOT)
    import bench

    assert bench.bytes_fwd(1, 768, 3136) + bench.bytes_bwd(1, 768, 3136) == pytest.approx(82.05e6, rel=2e-3)
    assert bench.workload_bytes(24) == pytest.approx(14.03e9, rel=2e-3)
