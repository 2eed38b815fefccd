#!/usr/bin/env bash
# diagnostics: per-role stall cycles of the Gram kernel (GB_GRAM_TRACE), optionally with probe masks
set -u
mkdir -p gpurun_out
for pr in "$@"; do
  echo "== probe $pr"
  GB_GRAM_TRACE=1 GB_GRAM_PROBE=$pr timeout -s KILL 90 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 >/dev/null | grep "gram trace" | tail -1
done
