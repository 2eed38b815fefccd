"""Window sharding across GPUs.  Host logic only -- no compute.

Every dist()/distmix() window is a closed computation (reference dist.cpp:63-75 rebuilds all state per
call), so a genome-wide run shards by window with NO data-path collective: each rank (one process per
GPU) takes one contiguous run of the bp-sorted window list, keeps the panel rows those windows touch
resident in its own HBM, and the per-window (z, info) arrays are gathered on the host at the end.
Runs are contiguous so that a rank's panel slice is one row range plus one wing of halo per boundary,
and cost-balanced because a window's work varies ~60x along a chromosome (n_t from 156 to 1,213 on chr22).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def window_cost(n_t, n_u, n_samples: int):
    """Relative device time of a window: Gram multiply-adds N*(n_u*n_t + n_t^2/2) on the tensor cores
    plus the fp64 solve n_t^2*n_u + n_t^3/3, the latter weighted by the measured ratio of the two
    rates (about 100 Gram MACs per fp64 FMA on B200).  Windows the reference refuses
    (<= 10 measured or unmeasured SNPs, dist.cpp:146) cost nothing."""
    n_t = np.asarray(n_t, np.float64)
    n_u = np.asarray(n_u, np.float64)
    gram = n_samples * (n_u * n_t + 0.5 * n_t * n_t)
    solve = n_t * n_t * n_u + n_t ** 3 / 3.0
    cost = gram + 100.0 * solve
    return np.where((n_t > 10) & (n_u > 10), cost, 0.0)


def partition_contiguous(costs: Sequence[float], n_parts: int) -> list[tuple[int, int]]:
    """Cut range(len(costs)) into n_parts contiguous [lo, hi) runs minimising the largest run cost
    (binary search on the bottleneck + greedy fill).  Runs may be empty when n_parts > len(costs)."""
    costs = np.asarray(costs, np.float64)
    n = len(costs)
    if n_parts < 1:
        raise ValueError("n_parts must be >= 1")
    if n == 0:
        return [(0, 0)] * n_parts

    def cuts_for(limit: float):
        cuts, acc, parts = [0], 0.0, 1
        for i, c in enumerate(costs):
            if acc + c > limit and acc > 0.0:
                cuts.append(i)
                acc = 0.0
                parts += 1
            acc += c
        return cuts, parts

    lo, hi = float(costs.max()), float(costs.sum())
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        _, parts = cuts_for(mid)
        if parts <= n_parts:
            hi = mid
        else:
            lo = mid
    cuts, parts = cuts_for(hi * (1.0 + 1e-12))   # the bisection's upper end is feasible up to rounding
    cuts = cuts[:n_parts] + [n] * (n_parts + 1 - min(len(cuts), n_parts))
    return [(cuts[i], cuts[i + 1]) for i in range(n_parts)]


def rows_needed(windows: Sequence[dict], lo: int, hi: int) -> tuple[int, int]:
    """Smallest [first, last] site-index range covering the measured (incl. wings) and unmeasured SNPs of
    windows lo..hi-1: the panel slice a rank must hold.  Returns (0, -1) for an empty run."""
    first, last = None, -1
    for w in windows[lo:hi]:
        for key in ("measured", "unmeasured"):
            idx = np.asarray(w[key])
            if len(idx):
                first = int(idx.min()) if first is None else min(first, int(idx.min()))
                last = max(last, int(idx.max()))
    return (0, -1) if first is None else (first, last)


def gather_window_results(local: dict, dist=None) -> dict:
    """Host-side gather: `local` maps window id -> (z, info, status) computed by this rank; returns the
    union on every rank (torch.distributed object all-gather over whatever backend is initialised --
    gloo in the tests; results are ~16 bytes per imputed SNP, this is not a bandwidth path)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    out: dict = {}
    for p in parts:
        dup = set(out) & set(p)
        if dup:
            raise RuntimeError(f"windows computed by two ranks: {sorted(dup)[:5]}")
        out.update(p)
    return out
