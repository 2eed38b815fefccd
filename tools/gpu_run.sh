#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench"; timeout -s KILL 240 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -c 600 gpurun_out/bench_latest.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_latest.json").read())
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "serial", round(d.get("stage_ms_serial",0),3), "stages", {k: round(v,3) for k,v in d["stage_ms"].items()})
print("e2e", {k:v for k,v in d["e2e"].items() if k!="note"}); print("e2e_pw", {k:v for k,v in d.get("e2e_per_window",{}).items() if k!="note"})
print("roofline", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), "launches", d["gpu_launches"], "clocks", d["clocks"])
print("cpu", d.get("cpu_baseline"))
PY
