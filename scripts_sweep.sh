#!/usr/bin/env bash
# tuning helper: parity tests + bench under several Gram configurations (run under gpurun)
set -u
mkdir -p gpurun_out
echo "== tests (default cluster)"; timeout -s KILL 150 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for cfg in "$@"; do
  fmt=${cfg%%:*}; shape=${cfg##*:}
  echo "== bench $fmt $shape"
  GB_PANEL_FORMAT=$fmt GB_GRAM_CLUSTER=$shape timeout -s KILL 90 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/sweep_${fmt}_$shape.json 2> gpurun_out/sweep_${fmt}_$shape.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sweep_${fmt}_$shape.json").read())
    print("$fmt $shape", "value", round(d["value"]), "gram_ms", round(d["stage_ms"]["gram"],3), "TOPS", round(d["roofline"]["achieved"],1), "ok", d["windows_ok"], "pack_ms", round(d["pack"]["ms"],2), d["clocks"]["sm_mhz"], d["clocks"].get("power_w_max"))
except Exception as e:
    print("$fmt $shape FAILED", e); print(open("gpurun_out/sweep_${fmt}_$shape.err").read()[-600:])
PY
done
