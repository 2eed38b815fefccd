import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gauss_b200 as gb
from gauss_b200 import synth
sizes = np.array([61, 103, 40, 25, 2, 330, 97, 128, 31, 33], np.int32)
g = synth.make_genotypes(330, sizes, seed=21)
ctx = gb.Context(0)
for fmt in ("int8", "e2m1"):
    panel = gb.Panel(ctx, sizes, len(g), fmt)
    panel.append_host(g.astype(np.int8), is_ascii=False)
    sxy, sx, sxx = panel.gram_counts(np.arange(0, 200), np.arange(193, 330))
    offs = np.concatenate([[0], np.cumsum(sizes)])
    bad = 0
    for p in range(len(sizes)):
        a = g[0:200, offs[p]:offs[p+1]].astype(np.int32); b = g[193:330, offs[p]:offs[p+1]].astype(np.int32)
        bad += int((sxy[p] != a @ b.T).sum()) + int((sx[p] != a.sum(1)).sum())
    print(fmt, "mismatches", bad, flush=True)
