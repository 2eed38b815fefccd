#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo skip tests
echo skip
echo "== bench default"; time timeout -s KILL 900 python bench.py --steps 3 --warmup 1 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 1500 gpurun_out/bench_r2b.err | grep -v "^\s" | tail -15; 
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2b.json").read().strip().split("\n")[-1])
print("value", d["value"], "ms", d["ms_per_step"], "wall", d.get("wall_ms_per_step"))
print("e2e", {k:v for k,v in d.get("e2e",{}).items() if k!="note"})
print("genome", {k:v for k,v in d["genome"].items() if k not in ("note","host_gather","shards")})
for k in ("roofline","roofline_gram","roofline_chol","roofline_expand5"):
    print(k, {a:b for a,b in d[k].items() if a not in ("note",)})
print("stage", d["stage_ms"], d["stage_ms_serial"])
print("chr22", {k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a!="note"}) for k,v in d["chr22"].items() if k!="config"})
print("int8", {k:v for k,v in d.get("int8",{}).items() if k not in ("roofline_gram","note")}, d.get("int8",{}).get("roofline_gram",{}).get("achieved"), d.get("int8",{}).get("roofline_gram",{}).get("frac"))
print("probes", {k:v for k,v in d["probes"].items() if k in ("i8","mxf4","fp64","copy")})
print("cpu", d.get("cpu_baseline"))
PY
