#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== parity"; timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_genome.py tests/test_qcat.py -m gpu -x -q 2>&1 | tail -5
GB_CHOL_TRACE=1 timeout -s KILL 200 python bench.py --workload chr22 --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/bench_linv.json 2> gpurun_out/bench_linv.err
grep "chol trace" gpurun_out/bench_linv.err | tail -1 | cut -c1-400
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_linv.json").read().strip().split("\n")[-1])
    print("value %.3f M  ms %.3f" % (d["value"]/1e6, d["ms_per_step"]), {kk: round(v,3) for kk,v in d["stage_ms"].items()}, round(d["stage_ms_serial"],3))
except Exception as e: print("no json", e)
PY
for c in 32 24 16; do echo "== genome CHAIN_SMS=$c"; GB_GENOME_CHAIN_SMS=$c timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | grep -E "step [12]|non-ok" | tail -3; done
