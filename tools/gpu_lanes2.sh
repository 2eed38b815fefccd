#!/usr/bin/env bash
set -u
for c in 32 48; do for skip in none chain heavy; do
  echo "== CHAIN_SMS=$c SKIP=$skip"; GB_GENOME_SKIP=$skip GB_GENOME_CHAIN_SMS=$c timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | grep -E "step [12]" | tail -2
done; done
