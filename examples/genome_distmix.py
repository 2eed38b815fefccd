#!/usr/bin/env python
"""Window-sharded distmix over one or more GPUs -- the shape of BASELINE config 4 at a size that runs in seconds.

    python examples/genome_distmix.py                                  # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29600 examples/genome_distmix.py                 # 2 GPUs, one process each

Every rank builds the same synthetic site table and window list, takes one contiguous cost-balanced run of windows
(gauss_b200.shard), packs ONLY the panel rows its windows touch into ternary host rows, runs the chromosome driver
(gb_chrom_run_pack5: host rows in, host z / info out) and the per-window results are gathered on the host -- no
collective touches genotype or correlation data (windows are independent, reference dist.cpp:63-75).
tests/test_example_genome.py runs `run()` and checks windows against the CPU oracle."""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import gauss_b200 as gb                      # noqa: E402
from gauss_b200 import api, shard, synth     # noqa: E402


def run(n_windows: int = 12, per_mb: float = 600.0):
    """-> dict(windows, results {window id: (z, info, status)}, bp, type, z_site, sizes, w, seed) on every rank."""
    args = argparse.Namespace(windows=n_windows, per_mb=per_mb)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")      # results only: ~16 bytes per imputed SNP

    # one synthetic chromosome: measured SNPs every ~3.5 kb, unmeasured sites between them, 1 Mb windows with 0.5 Mb wings
    rng = np.random.default_rng(4)
    span = args.windows * 1_000_000
    bp_m = np.sort(rng.choice(np.arange(1, span), size=int(span / 3500), replace=False))
    bp, type_, windows = synth.chr22_windows(bp_m, unmeasured_per_mb=args.per_mb)
    _, sizes, w = synth.flagged_33kg_pgc2()
    sizes = (sizes // 8).clip(min=5).astype(np.int32)     # 33KG's 21 flagged populations at an eighth of their size
    z_site = rng.standard_normal(len(bp)) * 1.34

    cost = shard.window_cost([len(x["measured"]) for x in windows], [len(x["unmeasured"]) for x in windows], int(sizes.sum()))
    lo, hi = shard.partition_contiguous(cost, world)[rank]
    first, last = shard.rows_needed(windows, lo, hi)
    n_rows = max(0, last - first + 1)

    # this rank's panel slice: generated here, on a real panel read from the packed file (gauss_b200.packfile)
    g = synth.make_genotypes(len(bp), sizes, seed=11)[first:last + 1] if n_rows else np.zeros((0, int(sizes.sum())), np.int8)
    rows5 = api.pack5_rows_host(sizes, g)

    t_off, u_off, rows_t, rows_u, z_t = [0], [0], [], [], []
    for x in windows[lo:hi]:
        rows_t.append(x["measured"] - first)
        rows_u.append(x["unmeasured"] - first)
        z_t.append(z_site[x["measured"]])
        t_off.append(t_off[-1] + len(x["measured"]))
        u_off.append(u_off[-1] + len(x["unmeasured"]))
    res = {}
    t0 = time.perf_counter()
    if hi > lo:
        ctx = gb.Context(local)
        panel = gb.Panel(ctx, sizes, max(n_rows, 1), "e2m1")
        z, info, status = panel.chrom_run_pack5(rows5.ctypes.data, n_rows, rows5.strides[0], t_off, np.concatenate(rows_t),
                                                u_off, np.concatenate(rows_u), np.concatenate(z_t), w, n_groups=2)
        for k in range(hi - lo):
            res[lo + k] = (z[u_off[k]:u_off[k + 1]], info[u_off[k]:u_off[k + 1]], int(status[k]))
    dt = time.perf_counter() - t0
    allres = shard.gather_window_results(res, dist)
    if rank == 0:
        n_imp = sum(len(v[0]) for v in allres.values() if v[2] == 0)
        print(f"{world} rank(s): {len(allres)} windows, {n_imp} imputed SNPs, rank 0 ran windows [{lo}, {hi}) on rows "
              f"[{first}, {last}] in {dt * 1e3:.1f} ms; statuses {sorted(set(v[2] for v in allres.values()))}")
        assert sorted(allres) == list(range(len(windows)))
    if dist is not None:
        dist.destroy_process_group()
    return dict(windows=windows, results=allres, bp=bp, type=type_, z_site=z_site, sizes=sizes, w=w, seed=11)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=12)
    ap.add_argument("--per-mb", type=float, default=600.0, help="synthetic unmeasured sites per Mb")
    a = ap.parse_args()
    run(a.windows, a.per_mb)
