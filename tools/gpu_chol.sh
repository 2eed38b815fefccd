#!/usr/bin/env bash
set -u
echo "== parity"; timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_genome.py tests/test_qcat.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do timeout -s KILL 200 python bench.py --workload chr22 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split(\"\n\")[-1])
print(\"value %.3f M  ms %.3f\" % (d[\"value\"]/1e6, d[\"ms_per_step\"]), {k: round(v,3) for k,v in d[\"stage_ms\"].items()}, {k: round(v,3) for k,v in d[\"stage_ms_alone\"].items() if k!=\"note\"})
"; done
timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | grep -E "step [12]" | tail -2
