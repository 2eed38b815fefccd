#!/usr/bin/env bash
set -u
NG=${NG:-8}
mkdir -p gpurun_out
head -2 /proc/meminfo; nproc
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 5 --warmup 2 > gpurun_out/bench_tr${NG}.json 2> gpurun_out/bench_tr${NG}.err; echo rc=$?; tail -3 gpurun_out/bench_tr${NG}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_tr${NG}.json").read().strip().split("\n")[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"],1), "wall", round(d.get("wall_ms_per_step",0),1), "n_gpus", d["n_gpus"])
print("e2e", {k:v for k,v in d.get("e2e",{}).items() if k!="note"})
g=d["genome"]; print("genome", g["device_ms_per_gpu"], g["result_checksum_u64"], "fill", g["fill_synthetic_s"], "plan", g["plan_s"], g["rank0_shard"])
print("clocks", d["clocks"])
PY
