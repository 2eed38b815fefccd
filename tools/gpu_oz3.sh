#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu suite"; timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for sms in 64 32 48 80 96; do
  GB_CHOL_SMS=$sms GB_OZ_TRACE=1 timeout -s KILL 200 python bench.py --workload chr22 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_oz_$sms.json 2> gpurun_out/bench_oz_$sms.err
  grep "oz trace" gpurun_out/bench_oz_$sms.err | tail -1
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_oz_$sms.json").read().strip().split("\n")[-1])
    print("chol_sms $sms value %.3f M  ms %.3f" % (d["value"]/1e6, d["ms_per_step"]), "stage", {k: round(v,3) for k,v in d["stage_ms"].items()}, round(d["stage_ms_serial"],3))
except Exception as e: print("no json", e)
PY
done
echo "== genome full"; timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | tail -4
