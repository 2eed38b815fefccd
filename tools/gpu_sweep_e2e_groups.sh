#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for g in "$@"; do
  GB_E2E_GROUPS=$g timeout -s KILL 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e2e_$g.json 2> gpurun_out/e2e_$g.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/e2e_$g.json").read())
    print("groups $g", "e2e", round(d["e2e"]["value"]), "ms", round(d["e2e"]["ms_per_step"],2), "| value", round(d["value"]), "ms/step", round(d["ms_per_step"],3))
except Exception as e:
    print("groups $g FAILED", e); print(open("gpurun_out/e2e_$g.err").read()[-800:])
PY
done
