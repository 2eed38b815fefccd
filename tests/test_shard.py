"""Host-side multi-GPU logic on CPU: cost-balanced contiguous window shards, panel-slice ranges, and the
host gather at world size 2 over gloo (windows are independent, no data-path collective)."""
import os
import socket

import numpy as np
import pytest

from gauss_b200 import shard, synth

SITES = os.path.join(os.path.dirname(__file__), "golden", "pgc2_chr22_sites.npz")


def chr22_windows():
    d = np.load(SITES)
    bp_m = np.unique(d["bp"].astype(np.int64))
    return synth.chr22_windows(bp_m)


def test_partition_is_contiguous_complete_and_balanced():
    bp, type_, windows = chr22_windows()
    n_t = np.array([len(w["measured"]) for w in windows])
    n_u = np.array([len(w["unmeasured"]) for w in windows])
    cost = shard.window_cost(n_t, n_u, 32147)
    assert (cost[(n_t <= 10) | (n_u <= 10)] == 0).all() and cost.max() / cost[cost > 0].min() > 10
    for parts in (1, 2, 4, 8):
        runs = shard.partition_contiguous(cost, parts)
        assert len(runs) == parts and runs[0][0] == 0 and runs[-1][1] == len(windows)
        assert all(a[1] == b[0] for a, b in zip(runs, runs[1:]))
        loads = np.array([cost[lo:hi].sum() for lo, hi in runs])
        # optimal bottleneck is at least the mean and at least the largest single window
        assert loads.max() <= max(cost.sum() / parts, cost.max()) * 1.35
    # more parts than windows: trailing runs are empty, nothing is lost
    runs = shard.partition_contiguous(cost[:3], 8)
    assert sum(hi - lo for lo, hi in runs) == 3 and len(runs) == 8
    assert shard.partition_contiguous([], 4) == [(0, 0)] * 4


def test_panel_slice_of_a_shard_covers_its_wings():
    bp, type_, windows = chr22_windows()
    runs = shard.partition_contiguous(np.ones(len(windows)), 4)
    for lo, hi in runs:
        first, last = shard.rows_needed(windows, lo, hi)
        for w in windows[lo:hi]:
            for key in ("measured", "unmeasured"):
                if len(w[key]):
                    assert first <= w[key].min() and w[key].max() <= last
    # neighbouring shards overlap by (at most) the wings, never by a core window
    (f0, l0), (f1, l1) = shard.rows_needed(windows, *runs[0]), shard.rows_needed(windows, *runs[1])
    assert f1 <= l0 and l0 - f1 < (l0 - f0) // 2


def _worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    bp, type_, windows = chr22_windows()
    n_t = np.array([len(w["measured"]) for w in windows])
    n_u = np.array([len(w["unmeasured"]) for w in windows])
    runs = shard.partition_contiguous(shard.window_cost(n_t, n_u, 32147), world)
    lo, hi = runs[rank]
    # stand-in for the device result of window w (placement-independent by construction)
    local = {w: (np.full(int(n_u[w]), float(w)), np.full(int(n_u[w]), 0.5), 0) for w in range(lo, hi)}
    merged = shard.gather_window_results(local, dist)
    ok = sorted(merged) == list(range(len(windows))) and all(
        merged[w][0].shape == (n_u[w],) and (merged[w][0] == w).all() for w in merged)
    q.put((rank, ok, hi - lo))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_at_world_size_2_over_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and sum(n for _, _, n in res) == 36
    # a window computed twice is an error, not a silent overwrite
    with pytest.raises(RuntimeError):
        class FakeDist:
            @staticmethod
            def is_initialized(): return True
            @staticmethod
            def get_world_size(): return 2
            @staticmethod
            def all_gather_object(parts, obj):
                parts[0], parts[1] = obj, obj
        shard.gather_window_results({1: (None, None, 0)}, FakeDist)
