"""Diagnostic: does a window's result depend on memory the batch never wrote?  (GB_POISON, run on the GPU box)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import gauss_b200 as gb  # noqa: E402
from helpers import small_case  # noqa: E402

c = small_case(seed=33, n_snps=900, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(0, 900))
g, t = c["g"].astype(np.int8), c["type"]
names = ["x", "pa", "pb", "y", "scr", "ut", "tt", "dinv", "zu", "info"]
for lo, hi in [(0, 300), (0, 900)]:
    idx = np.arange(lo, hi)
    rt, ru = idx[t[lo:hi] == 1], idx[t[lo:hi] == 0]
    os.environ.pop("GB_POISON", None)
    os.environ["GB_SOLVE"] = "fp64"
    ctx = gb.Context(0)
    panel = gb.Panel(ctx, c["pop_sizes"], len(g))
    panel.append_host(g, is_ascii=False)
    z0, i0, _ = panel.window_distmix(rt, ru, c["z"][rt], c["w"])
    ctx.close()
    for solver in ("int8", "fp64"):
        os.environ["GB_SOLVE"] = solver
        ctx = gb.Context(0)
        panel = gb.Panel(ctx, c["pop_sizes"], len(g))
        panel.append_host(g, is_ascii=False)
        for byte in (0x55, 0xFF, 0x00):
            for bit in range(10):
                os.environ["GB_POISON"] = "%d:%d" % (1 << bit, byte)
                z, i, _ = panel.window_distmix(rt, ru, c["z"][rt], c["w"])
                d = np.abs(z - z0)
                bad = np.isnan(d).any() or d.max() > 1e-9
                if bad:
                    print((lo, hi), solver, "poison", names[bit], hex(byte), "max|dz| %.3e" % np.nanmax(d), "nan", int(np.isnan(z).sum()),
                          "n_bad", int((~(d <= 1e-9)).sum()), "of", len(z), flush=True)
        os.environ.pop("GB_POISON", None)
        ctx.close()
print("done")
