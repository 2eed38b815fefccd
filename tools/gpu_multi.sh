#!/usr/bin/env bash
# multi-GPU validation: N=${NG:-2} GPUs of one box
set -u
NG=${NG:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; head -2 /proc/meminfo; nproc
echo "== multi-GPU tests"; timeout -s KILL 600 python -m pytest tests/test_genome.py -m gpu -x -q 2>&1 | tail -3
echo "== single-process bench, $NG GPUs"
timeout -s KILL 400 python bench.py --gpus $NG --single-process --steps 3 --warmup 1 --quick --no-cpu-baseline > gpurun_out/bench_sp${NG}.json 2> gpurun_out/bench_sp${NG}.err; echo rc=$?; tail -3 gpurun_out/bench_sp${NG}.err
echo "== torchrun bench, $NG ranks"
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 1 --quick > gpurun_out/bench_tr${NG}.json 2> gpurun_out/bench_tr${NG}.err; echo rc=$?; tail -3 gpurun_out/bench_tr${NG}.err
python - <<PY
import json
for f in ("gpurun_out/bench_sp${NG}.json","gpurun_out/bench_tr${NG}.json"):
    try:
        d=json.loads(open(f).read().strip().split("\n")[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", round(d["value"]), "ms", round(d["ms_per_step"],1), "wall", round(d.get("wall_ms_per_step",0),1), "n_gpus", d["n_gpus"])
    print("  e2e", {k:v for k,v in d.get("e2e",{}).items() if k!="note"})
    g=d["genome"]; print("  genome", g["device_ms_per_gpu"], g["result_checksum_u64"], "fill", g["fill_synthetic_s"], "plan", g["plan_s"], g["rank0_shard"])
PY
