#!/usr/bin/env bash
# round-2 profile pass: plain run must exit 0 first, then the launch list, then one --set full capture
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02}
ARGS="--workload chr22 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
echo "== tests"; timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout -s KILL 200 python bench.py $ARGS > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
echo "plain ok"
K='regex:gram_seg|gram_fin|chol_|trsm_|row_prep|pack_rows|expand5|pd_bound|copy_shift|gather_rows|synth_pack5|probe'
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $ARGS > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 700 ncu --set full --clock-control none --import-source on -k 'regex:gram_seg|gram_finalize|trsm_finalize|expand5_rows|chol_update' -s 3 -c 9 -o gpurun_out/prof_${TAG} -f python bench.py $ARGS > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/prof_${TAG}.ncu-rep
