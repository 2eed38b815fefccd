#!/usr/bin/env python
"""ncu report -> the handful of per-launch figures the bench and DESIGN.md quote.
usage: python tools/ncu_rows.py gpurun_out/prof_r02c.ncu-rep profiles/r02c_ncu_rows.json"""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = {
    "kernel": "Kernel Name", "grid": "Grid Size", "dur_us": "gpu__time_duration.sum",
    "dram_rd": "dram__bytes_read.sum", "dram_wr": "dram__bytes_write.sum",
    "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "imma_pct": "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "dmma_inst_pct": "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "fp64_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct", "regs": "launch__registers_per_thread",
    "smem_dyn_kb": "launch__shared_mem_per_block_dynamic",
}
res = []
for d in data:
    r = {}
    for k, name in want.items():
        if name in hdr:
            i = hdr.index(name)
            v = d[i]
            if k == "kernel":
                v = v.split("(")[0].split("::")[-1]
            r[k] = v
            if units[i] and k not in ("kernel", "grid"):
                r[k + "_unit"] = units[i]
    res.append(r)
json.dump(res, open(out, "w"), indent=0)
for r in res:
    print({k: v for k, v in r.items() if not k.endswith("_unit")})
