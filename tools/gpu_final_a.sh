#!/usr/bin/env bash
# full GPU suite + profile pass (plain run first, then launch list, then one --set full capture)
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02c}
echo "== tests"; timeout -s KILL 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ARGS="--workload chr22 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
timeout -s KILL 200 python bench.py $ARGS > gpurun_out/plain_${TAG}.json 2> gpurun_out/plain_${TAG}.err || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.err; exit 1; }
echo "plain ok"
python - <<PY
import json
d=json.loads(open("gpurun_out/plain_${TAG}.json").read().strip().split("\n")[-1])
print("value %.3f M ms %.3f" % (d["value"]/1e6, d["ms_per_step"]), d["stage_ms"], d["stage_ms_serial"])
for k in ("roofline","roofline_solve","roofline_linv","roofline_finish","roofline_chol","solve"):
    r=d.get(k); print(k, {kk: (round(v,4) if isinstance(v,float) else v) for kk,v in (r or {}).items() if kk!="note"})
PY
K='regex:gram_seg|gram_fin|chol_|trsm_|row_prep|pack_rows|expand5|pd_bound|copy_shift|gather_rows|synth_pack5|probe|ozaki|oz_'
timeout -s KILL 500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py $ARGS > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k 'regex:gram_seg|gram_finalize|trsm_finalize|ozaki_solve|oz_slice_x|chol_update' -s 6 -c 16 -o gpurun_out/prof_${TAG} -f python bench.py $ARGS > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/prof_${TAG}.ncu-rep
