#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== genome tests"; timeout -s KILL 600 python -m pytest tests/test_genome.py -m gpu -x -q 2>&1 | tail -5
echo "== genome full default"; timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | tail -6
echo "== genome full batch 96"; GB_GENOME_BATCH_WINDOWS=96 timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | tail -4
echo "== genome full batch 24"; GB_GENOME_BATCH_WINDOWS=24 timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | tail -4
echo "== genome full pack5 only"; GB_GENOME_RESIDENT=pack5 timeout -s KILL 300 python tools/genome_try.py --steps 3 2>&1 | tail -3
nvidia-smi --query-gpu=memory.used --format=csv
