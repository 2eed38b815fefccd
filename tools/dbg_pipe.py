"""Diagnostic: the lagging-wait flow of test_pipe_matches_single_window_calls under both solvers (run on the GPU box)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import gauss_b200 as gb  # noqa: E402
from helpers import small_case  # noqa: E402

c = small_case(seed=33, n_snps=900, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(0, 900))
g, t = c["g"].astype(np.int8), c["type"]
variants = {
    "test": [(0, 300), (150, 520), (400, 900), (880, 900), (300, 700), (100, 400), (0, 900)],
    "no_reject": [(0, 300), (150, 520), (400, 900), (300, 700), (100, 400), (0, 900)],
    "pre": [(0, 300), (150, 520), (400, 900), (300, 700)],
}
res = {}
order = sys.argv[1].split(',') if len(sys.argv) > 1 else list(variants)
for vname in order:
  wins = variants[vname]
  if True:
    for solver in ((sys.argv[2],) if len(sys.argv) > 2 else ("int8", "fp64")):
        os.environ["GB_SOLVE"] = solver
        ctx = gb.Context(0)
        panel = gb.Panel(ctx, c["pop_sizes"], len(g))
        panel.append_host(g, is_ascii=False)
        pre = {}
        if vname == "pre":
            for k, (lo, hi) in enumerate(wins):
                idx = np.arange(lo, hi)
                rt, ru = idx[t[lo:hi] == 1], idx[t[lo:hi] == 0]
                pre[k] = panel.window_distmix(rt, ru, c["z"][rt], c["w"])[0].copy()
        pipe = gb.Pipe(ctx, c["pop_sizes"], 900, depth=2)
        tickets, outs, sts = [], [], {}
        for lo, hi in wins:
            idx = np.arange(lo, hi)
            rt, ru = idx[t[lo:hi] == 1], idx[t[lo:hi] == 0]
            tk, z, info = pipe.submit(g[rt], g[ru], c["z"][rt], c["w"])
            tickets.append(tk)
            outs.append((rt, ru, z, info))
            if len(tickets) >= 2:
                k = len(tickets) - 2
                sts[k] = pipe.wait(tickets[k])
        sts[len(tickets) - 1] = pipe.wait(tickets[-1])
        for k, (rt, ru, z, info) in enumerate(outs):
            if len(ru) < 10 or len(rt) < 10:
                continue
            z1, i1, _ = panel.window_distmix(rt, ru, c["z"][rt], c["w"])
            res[(vname, solver, k)] = (z.copy(), z1.copy())
            print(vname, solver, "win", k, wins[k], "n_t", len(rt), "n_u", len(ru), "status", sts[k],
                  "pipe-vs-call z %.3e info %.3e" % (np.abs(z - z1).max(), np.abs(info - i1).max()),
                  "nan", int(np.isnan(z).sum()), "pre-vs-call %.3e" % (np.abs(pre[k] - z1).max() if pre else -1), flush=True)
        pipe.close()
        ctx.close()
for key in sorted(k for k in res if k[1] == "int8" and (k[0], "fp64", k[2]) in res):
    a, b = res[key], res[(key[0], "fp64", key[2])]
    print(key, "int8-vs-fp64: pipe z %.3e  call z %.3e" % (np.abs(a[0] - b[0]).max(), np.abs(a[1] - b[1]).max()))
