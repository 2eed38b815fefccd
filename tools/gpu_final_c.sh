#!/usr/bin/env bash
# final N=1 evidence: full GPU suite (both solvers), smoke, launch list, ncu capture, default bench, reference arm
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02e}
echo "== tests (default solver)"; timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== tests (GB_SOLVE=fp64)"; GB_SOLVE=fp64 timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_qcat.py tests/test_gpu_configs.py -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
ARGS="--workload chr22 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout -s KILL 200 python bench.py $ARGS > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
echo "plain ok"
K='regex:gram_seg|gram_fin|chol_|trsm_|linv_|row_prep|pack_rows|expand5|pd_bound|copy_shift|gather_rows|synth_pack5|probe|ozaki|oz_'
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 800 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $ARGS > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_seg|gram_finalize|ozaki_solve|oz_slice_x|linv_row' -s 8 -c 14 -o gpurun_out/prof_${TAG} -f python bench.py $ARGS > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"
echo "== bench N=1 default"; ( time timeout -s KILL 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; tail -c 300 gpurun_out/bench_n1.err
echo "== reference arm"; ( time timeout -s KILL 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -c 200 gpurun_out/bench_ref.json
