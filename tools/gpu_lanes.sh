#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== genome tests"; timeout -s KILL 900 python -m pytest tests/test_genome.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for c in 0 32 24 40 48; do
  echo "== GB_GENOME_CHAIN_SMS=$c"; GB_GENOME_CHAIN_SMS=$c timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | grep -E "step [12]|non-ok" | tail -3
done
