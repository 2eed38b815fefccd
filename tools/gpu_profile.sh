#!/usr/bin/env bash
# round-1 final profile pass: plain run must exit 0 first, then the launch list, then one --set full capture
set -u
mkdir -p gpurun_out
ARGS="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout -s KILL 120 python bench.py $ARGS > gpurun_out/plain_${TAG:-r1}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG:-r1}.log; exit 1; }
echo "plain ok"
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gram_seg|gram_fin|chol_|trsm_|row_prep|pack_rows|expand2|pd_bound|copy_shift|gather_rows" -c 400 --csv --log-file gpurun_out/launches_${TAG:-r1}.csv python bench.py $ARGS > gpurun_out/ncu_launch_${TAG:-r1}.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"gram_seg|gram_finalize|trsm_finalize" -s 6 -c 4 -o gpurun_out/prof_${TAG:-r1} python bench.py $ARGS > gpurun_out/ncu_full_${TAG:-r1}.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/prof_${TAG:-r1}.ncu-rep
