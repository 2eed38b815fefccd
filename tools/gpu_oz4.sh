#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== gpu parity"; timeout -s KILL 1500 python -m pytest tests/test_gpu_parity.py tests/test_genome.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -4
GB_OZ_TRACE=1 timeout -s KILL 200 python bench.py --workload chr22 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_oz.json 2> gpurun_out/bench_oz.err
grep "oz trace" gpurun_out/bench_oz.err | tail -1
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_oz.json").read().strip().split("\n")[-1])
    print("value %.3f M  ms %.3f" % (d["value"]/1e6, d["ms_per_step"]), "stage", {k: round(v,3) for k,v in d["stage_ms"].items()}, round(d["stage_ms_serial"],3))
except Exception as e: print("no json", e)
PY
echo "== genome full"; timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | tail -4
