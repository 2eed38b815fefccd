#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used --format=csv | head -3
head -3 /proc/meminfo; nproc
echo "== genome tests"; timeout -s KILL 600 python -m pytest tests/test_genome.py -m gpu -x -q 2>&1 | tail -15
echo "== all gpu tests"; timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== genome small"; timeout -s KILL 300 python tools/genome_try.py --mb 48,51,59 --steps 3 2>&1 | tail -12
echo "== genome small 1 stream"; GB_GENOME_STREAMS=1 timeout -s KILL 300 python tools/genome_try.py --mb 48,51,59 --steps 3 2>&1 | tail -4
echo "== genome small batch 24"; GB_GENOME_BATCH_WINDOWS=24 timeout -s KILL 300 python tools/genome_try.py --mb 48,51,59 --steps 3 2>&1 | tail -4
echo "== genome small batch 96"; GB_GENOME_BATCH_WINDOWS=96 timeout -s KILL 300 python tools/genome_try.py --mb 48,51,59 --steps 3 2>&1 | tail -4
echo "== genome small e2m1"; GB_GENOME_RESIDENT=e2m1 timeout -s KILL 300 python tools/genome_try.py --mb 48,51,59 --steps 3 2>&1 | tail -4
echo "== genome full"; timeout -s KILL 600 python tools/genome_try.py --steps 3 2>&1 | tail -12
