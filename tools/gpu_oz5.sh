#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== smoke (v2)"; timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== parity (v2)"; timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -5
for k in 2 1; do
GB_OZ_KERNEL=$k GB_OZ_TRACE=1 timeout -s KILL 200 python bench.py --workload chr22 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_oz_k$k.json 2> gpurun_out/bench_oz_k$k.err
grep "oz trace" gpurun_out/bench_oz_k$k.err | tail -1
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_oz_k$k.json").read().strip().split("\n")[-1])
    print("kernel v$k value %.3f M  ms %.3f" % (d["value"]/1e6, d["ms_per_step"]), {kk: round(v,3) for kk,v in d["stage_ms"].items()}, round(d["stage_ms_serial"],3))
except Exception as e: print("no json", e)
PY
done
