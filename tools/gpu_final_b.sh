#!/usr/bin/env bash
# final N=1 evidence: ncu capture of the step's kernels, smoke, default bench, reference arm
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02d}
ARGS="--workload chr22 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout -s KILL 200 python bench.py $ARGS > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
echo "plain ok"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_seg|gram_finalize|trsm_finalize|ozaki_solve|oz_slice_x' -s 5 -c 8 -o gpurun_out/prof_${TAG} -f python bench.py $ARGS > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"
echo "== smoke"; timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== bench N=1 default"; ( time timeout -s KILL 1500 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2>&1 | grep real; tail -c 300 gpurun_out/bench_n1.err
echo "== reference arm"; ( time timeout -s KILL 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real; tail -c 600 gpurun_out/bench_ref.json
