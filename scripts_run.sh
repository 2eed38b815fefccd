#!/usr/bin/env bash
# gpurun helper: GPU parity tests, then one full bench line (tight timeouts, nothing can hang the box)
set -u
mkdir -p gpurun_out
echo "== tests"; timeout -s KILL 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== bench"; timeout -s KILL 200 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -c 400 gpurun_out/bench_latest.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_latest.json").read())
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "stages", {k: round(v,3) for k,v in d["stage_ms"].items()})
print("e2e", d["e2e"]); print("roofline", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), "launches", d["gpu_launches"], "clocks", d["clocks"])
print("cpu", d.get("cpu_baseline"))
PY
