#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== default solver (int8-split): full gpu suite"; timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
echo "== fp64 solver: parity subset"; GB_SOLVE=fp64 timeout -s KILL 400 python -m pytest tests/test_gpu_parity.py tests/test_genome.py -m gpu -x -q 2>&1 | tail -4
echo "== chr22 stage times (int8-split)"; GB_OZ_TRACE=1 timeout -s KILL 200 python bench.py --workload chr22 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_oz.json 2> gpurun_out/bench_oz.err; echo rc=$?; grep "oz trace" gpurun_out/bench_oz.err | tail -2
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_oz.json").read().strip().split("\n")[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "stage", d["stage_ms"], d["stage_ms_serial"])
except Exception as e: print("no json", e)
PY
echo "== genome full"; timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | tail -5
