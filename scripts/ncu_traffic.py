"""profiles/rNN_ncu_launches_bench.csv (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none on `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-model`)  ->
profiles/rNN_ncu_traffic.json: per-launch DRAM bytes and time of the scan kernels by stage, and each kernel's share of the step."""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_ncu_launches_bench.csv")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = next(r for r in rows if r[0] == "ID")
ix = {h: i for i, h in enumerate(hdr)}
launches = {}
for r in rows:
    if r[0] == "ID" or not r[0].isdigit():
        continue
    d = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    if r[ix["Metric Name"]] == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)          # -> us
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)   # -> bytes
    d[r[ix["Metric Name"]]] = v
stage_of_grid = {288: "S1", 576: "S2", 1152: "S3", 2304: "S4"}
calls = {"S1": 4, "S2": 4, "S3": 4, "S4": 2}
acc = {}
# the forward is persistent (grid = min(work items, CTA slots)): its stage is the one of the backward launch that follows it
ordered = [launches[k] for k in sorted(launches)]
pending_fwd = []
for d in ordered:
    kind = "fwd" if "selscan_fwd_tma" in d["name"] else ("bwd" if "selscan_bwd_ws" in d["name"] else None)
    if kind is None:
        continue
    if kind == "fwd":
        pending_fwd.append(d)
        continue
    g = int(d["grid"].strip("()").split(",")[0])
    if g not in stage_of_grid:
        pending_fwd.clear()
        continue
    for item, k in [(f, "fwd") for f in pending_fwd] + [(d, "bwd")]:
        a = acc.setdefault((stage_of_grid[g], k), {"n": 0, "t": 0.0, "r": 0.0, "w": 0.0})
        a["n"] += 1
        a["t"] += item["gpu__time_duration.sum"]
        a["r"] += item["dram__bytes_read.sum"]
        a["w"] += item["dram__bytes_write.sum"]
    pending_fwd.clear()
step_us = sum(calls[s] * a["t"] / a["n"] for (s, _), a in acc.items())
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none on `python "
                 "bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-model --no-reference-cuda` (batch 24), " + os.path.relpath(src, ROOT) + "; "
                 "per-launch means; times are cold-cache/serialised (compare shares, not absolutes)", "stages": {}}
for (s, k), a in sorted(acc.items()):
    out["stages"].setdefault(s, {})[k] = {
        "dram_read_bytes": int(a["r"] / a["n"]), "dram_write_bytes": int(a["w"] / a["n"]), "dram_bytes": int((a["r"] + a["w"]) / a["n"]),
        "ncu_time_us": round(a["t"] / a["n"], 1), "launches": a["n"], "share_of_step": round(calls[s] * a["t"] / a["n"] / step_us, 4)}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["stages"]["S1"], indent=1))
