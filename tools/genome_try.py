#!/usr/bin/env python
"""Quick genome-driver run (tuning helper, not the bench): python tools/genome_try.py [--mb 48,51] [--gpus N] [--steps K]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gauss_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--mb", default="all")
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--upload", action="store_true")
a = ap.parse_args()
mb = synth.HG19_MB if a.mb == "all" else [int(x) for x in a.mb.split(",")]
_, sizes, w = synth.flagged_33kg_pgc2()
t0 = time.time(); chroms = synth.genome_layout(mb); t_lay = time.time() - t0
g = api.Genome(a.gpus, sizes, w)
for c in chroms:
    g.add_chromosome(c["n_rows"], c["t_off"], c["rows_t"], c["u_off"], c["rows_u"], c["z_t"], sites=c["sites"])
t0 = time.time(); g.plan(); t_plan = time.time() - t0
t0 = time.time(); g.fill_synthetic(20260101); t_fill = time.time() - t0
infos = [g.shard_info(i) for i in range(a.gpus)]
n_imp = sum(x["n_imputed"] for x in infos)
print(f"layout {t_lay:.1f}s plan {t_plan:.2f}s fill {t_fill:.2f}s windows {sum(x['n_windows'] for x in infos)} imputed {n_imp}")
for x in infos: print(x)
z = info = st = None
for k in range(a.steps):
    t0 = time.time(); z, info, st, ms = g.run(z, info, st); wall = time.time() - t0
    print(f"step {k}: wall {wall*1e3:.1f} ms, gpu ms {np.round(ms,1)}, {n_imp/wall/1e6:.2f} M SNPs/s (wall) {n_imp/(ms.max()/1e3)/1e6:.2f} (device)")
bad = sum(int((s != 0).sum()) for s in st)
print("non-ok windows", bad, "nan z", sum(int(np.isnan(x).sum()) for x in z), "launches", g.launch_count)
print(open("/proc/meminfo").read().split("\n")[0:3])
g.close()
