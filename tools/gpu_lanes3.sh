#!/usr/bin/env bash
set -u
echo "== genome tests"; timeout -s KILL 900 python -m pytest tests/test_genome.py -m gpu -x -q 2>&1 | tail -3
for d in 2 3 1; do for c in 16 24 32; do
  echo "== DEPTH=$d CHAIN_SMS=$c"; GB_GENOME_LANE_DEPTH=$d GB_GENOME_CHAIN_SMS=$c timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | grep -E "step 2|non-ok|expanded" | cut -c1-200 | tail -2
done; done
