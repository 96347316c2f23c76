#!/bin/bash
# A/B timing of library variants on one B200: per-stage table of the scan bench for each .so given (default build first).
# usage: scripts/gpu_ab.sh tag1=path1.so tag2=path2.so ...   (results: gpurun_out/ab_<tag>.json)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline > gpurun_out/ab_default.json 2> gpurun_out/ab_default.err
for kv in "$@"; do
  tag="${kv%%=*}"; lib="${kv#*=}"
  SELSCAN_B200_LIB="$PWD/$lib" python bench.py --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline > "gpurun_out/ab_$tag.json" 2> "gpurun_out/ab_$tag.err"
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/ab_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], {k: (v["fwd_ms"], v["bwd_ms"]) for k, v in d["per_stage"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
