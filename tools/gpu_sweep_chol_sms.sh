#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== tests"; timeout -s KILL 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for r in "$@"; do
  GB_CHOL_SMS=$r timeout -s KILL 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/chol_$r.json 2> gpurun_out/chol_$r.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/chol_$r.json").read())
    print("chol_sms $r", "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), "serial", round(d["stage_ms_serial"],3), {k: round(v,3) for k,v in d["stage_ms"].items()}, "ok", d["windows_ok"])
except Exception as e:
    print("chol_sms $r FAILED", e); print(open("gpurun_out/chol_$r.err").read()[-600:])
PY
done
