#!/usr/bin/env bash
# diagnostics: where the Gram kernel's MMA thread waits (GB_GRAM_TRACE)
set -u
GB_GRAM_TRACE=1 timeout -s KILL 90 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 >/dev/null | grep "gram trace" | tail -2
