import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gauss_b200 as gb
from gauss_b200 import synth
_, sizes, w = synth.flagged_33kg_pgc2()
g = synth.make_genotypes(200, sizes, seed=9).astype(np.int8)
g[0] = 2; g[1] = 2; g[2] = 1; g[1, ::5] = 1
ctx = gb.Context(0)
offs = np.concatenate([[0], np.cumsum(sizes)])
res = {}
for fmt in ("int8", "e2m1"):
    panel = gb.Panel(ctx, sizes, len(g), fmt)
    panel.append_host(g, is_ascii=False)
    sxy, sx, sxx = panel.gram_counts(np.arange(0, 130), np.arange(0, 200))
    bad = 0
    for p in range(len(sizes)):
        a = g[0:130, offs[p]:offs[p+1]].astype(np.int32); b = g[0:200, offs[p]:offs[p+1]].astype(np.int32)
        bad += int((sxy[p] != a @ b.T).sum())
    B11, B21 = panel.window_cor(np.arange(0, 60), np.arange(60, 200), None)
    res[fmt] = (B11, B21)
    print(fmt, os.environ.get("GB_GRAM_KIND", "default"), "count mismatches", bad, "max count", int(sxy.max()), flush=True)
print("pooled B11/B21 identical:", (res["int8"][0] == res["e2m1"][0]).all(), (res["int8"][1] == res["e2m1"][1]).all())
