#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu tests"; timeout -s KILL 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
echo "== bench default"; time (timeout -s KILL 900 python bench.py --steps 5 --warmup 2 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err); tail -5 gpurun_out/bench_r2d.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2d.json").read().strip().split("\n")[-1])
print("value", d["value"], "ms", d["ms_per_step"], "wall", d.get("wall_ms_per_step"))
print("e2e", {k:v for k,v in d.get("e2e",{}).items() if k!="note"})
g=d["genome"]; print("genome fill", g["fill_synthetic_s"], "plan", g["plan_s"], g["result_checksum_u64"], g["gram_tops"], g["solve_tflops"])
for k in ("roofline","roofline_gram","roofline_chol","roofline_expand5"):
    print(k, {a:b for a,b in d[k].items() if a in ("achieved","peak","frac","ms")})
print("stage", d["stage_ms"], d["stage_ms_serial"])
print("chr22", d["chr22"]["value"], {k:v.get("value") for k,v in d["chr22"].items() if isinstance(v,dict) and "value" in v})
print("int8", d.get("int8",{}).get("value"))
PY
