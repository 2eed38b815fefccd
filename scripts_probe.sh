#!/usr/bin/env bash
# diagnostics: time the Gram kernel with parts switched off (GB_GRAM_PROBE bitmask: 1 no TMA, 2 no MMA, 4 no fold)
set -u
mkdir -p gpurun_out
echo "== tests"; timeout -s KILL 240 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for pr in "$@"; do
  GB_GRAM_PROBE=$pr timeout -s KILL 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/probe_$pr.json 2> gpurun_out/probe_$pr.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/probe_$pr.json").read())
    print("probe $pr", "gram_ms", round(d["stage_ms"]["gram"],3), "TOPS", round(d["roofline"]["achieved"],1), "step", round(d["ms_per_step"],3), d["clocks"]["sm_mhz"])
except Exception as e:
    print("probe $pr FAILED", e); print(open("gpurun_out/probe_$pr.err").read()[-600:])
PY
done
