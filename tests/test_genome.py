"""Genome driver (gb_genome_*): one process, 1..8 GPUs, contiguous cost-balanced shards, resident
rows, host gather.  Parity: window results do not depend on placement -- 1 part vs 3 parts vs N GPUs byte-identical
(SURVEY.md section 4 v) -- and sampled windows equal the oracle on the very rows the device generator produced."""
import os

import numpy as np
import pytest

from gauss_b200 import api, synth

POPS = np.array([61, 103, 40, 25, 2, 330, 97], np.int32)
SEED = 20260101


def tiny_genome():
    return synth.genome_layout(chrom_mb=[5, 7, 4], measured_per_mb=55.0, unmeasured_per_mb=180.0, seed=11)


def weights():
    return np.random.default_rng(3).dirichlet(np.ones(len(POPS))) * 1.061


# ---- host logic (no GPU) ---------------------------------------------------------------------------------------------
def test_partition_is_contiguous_balanced_and_matches_python():
    from gauss_b200 import shard
    rng = np.random.default_rng(0)
    nt = rng.integers(5, 1300, 700)
    nu = rng.integers(0, 4000, 700)
    for parts in (1, 2, 3, 8, 64):
        cuts, cost = api.partition_windows(nt, nu, 32147, parts)
        assert cuts[0] == 0 and cuts[-1] == 700 and (np.diff(cuts) >= 0).all()
        loads = np.array([cost[cuts[i]:cuts[i + 1]].sum() for i in range(parts)])
        # bottleneck within one window of the ideal share
        assert loads.max() <= cost.sum() / parts + cost.max() + 1e-6
        ref = shard.partition_contiguous(cost, parts)
        assert max(cost[a:b].sum() for a, b in ref) == pytest.approx(loads.max(), rel=1e-9)
    # refused windows (<= 10 SNPs, dist.cpp:146) cost (almost) nothing
    _, cost = api.partition_windows([10, 11], [500, 500], 1000, 1)
    assert cost[0] == 1.0 and cost[1] > 1e6


def test_unpack5_is_the_inverse_of_the_host_packer():
    g = synth.make_genotypes(9, POPS, seed=2)
    assert np.array_equal(api.unpack5_rows(api.pack5_rows_host(POPS, g), POPS), g)


def test_layout_windows_follow_the_reference_rule():
    ch = tiny_genome()[1]
    n_m = ch["n_measured"]
    for w in (0, 3, len(ch["start_bp"]) - 1):
        s = ch["start_bp"][w]
        rt = ch["rows_t"][ch["t_off"][w]:ch["t_off"][w + 1]]
        ru = ch["rows_u"][ch["u_off"][w]:ch["u_off"][w + 1]] - n_m
        want_t = np.where((ch["bp_m"] >= s - 500_000) & (ch["bp_m"] <= s + 999_999 + 500_000))[0]   # dist.cpp:136-141
        want_u = np.where((ch["bp_u"] >= s) & (ch["bp_u"] <= s + 999_999))[0]                       # dist.cpp:132-135
        assert np.array_equal(rt, want_t) and np.array_equal(ru, want_u)
    assert len(set(ch["sites"].tolist())) == ch["n_rows"]


# ---- GPU -------------------------------------------------------------------------------------------------------------
def run_genome(chroms, n_gpus=1, n_parts=None, first_part=0, host_rows=None, env=None, devices=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        g = api.Genome(n_gpus, POPS, weights(), devices=devices)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    for i, c in enumerate(chroms):
        kw = {}
        if host_rows is not None and i != 1:
            kw = dict(rows5_ptr=host_rows[i].ctypes.data, row_stride=host_rows[i].strides[0])
        g.add_chromosome(c["n_rows"], c["t_off"], c["rows_t"], c["u_off"], c["rows_u"], c["z_t"], sites=c["sites"], **kw)
    g.plan(n_parts or n_gpus, first_part)
    if host_rows is not None:
        # chromosome 1: handed over piece by piece, only the rows the shard keeps resident (what a feeder would read)
        keep = []
        for gi in range(n_gpus):
            for lo, hi in g.resident_ranges(gi, 1):
                piece = np.ascontiguousarray(host_rows[1][lo:hi])
                keep.append(piece)
                g.set_host_rows(1, lo, hi - lo, piece.ctypes.data, piece.strides[0])
        g.upload(wait=False)     # asynchronous: the batches wait for their own rows
    else:
        g.fill_synthetic(SEED)
    z, info, status, ms = g.run()
    infos = [g.shard_info(i) for i in range(n_gpus)]
    launches = g.launch_count
    g.close()
    return z, info, status, ms, infos, launches


@pytest.mark.gpu
def test_genome_equals_oracle_and_is_placement_independent(gpu_ctx, oracle):
    chroms = tiny_genome()
    z1, i1, s1, ms, infos, launches = run_genome(chroms)
    assert launches > 0 and ms[0] > 0
    nw = sum(len(c["t_off"]) - 1 for c in chroms)
    assert infos[0]["n_windows"] == nw and infos[0]["n_batches"] >= 3
    for c, st in zip(chroms, s1):
        nt, nu = np.diff(c["t_off"]), np.diff(c["u_off"])
        assert np.array_equal(st == 0, (nt > 10) & (nu > 10))
    # (a) oracle on sampled windows, fed with the rows the device generator made
    w = weights()
    for ci, wi in ((0, 1), (1, 4), (2, 0)):
        c = chroms[ci]
        rows5 = api.synth_pack5_rows(gpu_ctx, SEED, ci, POPS, c["n_rows"], sites=c["sites"])
        g = api.unpack5_rows(rows5, POPS)
        rt = c["rows_t"][c["t_off"][wi]:c["t_off"][wi + 1]]
        ru = c["rows_u"][c["u_off"][wi]:c["u_off"][wi + 1]]
        if len(rt) <= 10 or len(ru) <= 10:
            continue
        sel = np.concatenate([rt, ru])
        t = np.concatenate([np.ones(len(rt), np.int32), np.zeros(len(ru), np.int32)])
        zz = np.concatenate([c["z_t"][c["t_off"][wi]:c["t_off"][wi + 1]], np.zeros(len(ru))])
        r = oracle.run_window(t, np.arange(len(sel), dtype=np.int64), zz, g[sel], POPS, w, 0, 10 ** 12)
        assert r["rc"] == 0
        got_z = z1[ci][c["u_off"][wi]:c["u_off"][wi + 1]]
        got_i = i1[ci][c["u_off"][wi]:c["u_off"][wi + 1]]
        assert np.abs(got_z - r["z"][len(rt):]).max() <= 1e-6        # north-star bar
        assert np.abs(got_i - r["info"][len(rt):]).max() <= 1e-6
        assert np.abs(got_z - r["z"][len(rt):]).max() <= 1e-9        # achieved
    # (b) uploaded host rows (asynchronous upload) == generated rows, bit for bit
    host = [api.synth_pack5_rows(gpu_ctx, SEED, ci, POPS, c["n_rows"], sites=c["sites"]) for ci, c in enumerate(chroms)]
    z2, i2, s2, *_ = run_genome(chroms, host_rows=host)
    # (c) three parts computed by three separate genomes (what three ranks would do) == one part
    z3 = [np.full(len(a), np.nan) for a in z1]
    i3 = [np.full(len(a), np.nan) for a in z1]
    seen = 0
    for part in range(3):
        zp, ip, sp, _, inf, _ = run_genome(chroms, n_parts=3, first_part=part)
        seen += inf[0]["n_windows"]
        # a part only writes its own windows: merge what it produced
        g0 = inf[0]["first_window"]
        off = 0
        for ci, c in enumerate(chroms):
            n_w = len(c["t_off"]) - 1
            lo, hi = max(g0 - off, 0), min(g0 + inf[0]["n_windows"] - off, n_w)
            if hi > lo:
                a, b = c["u_off"][lo], c["u_off"][hi]
                z3[ci][a:b], i3[ci][a:b] = zp[ci][a:b], ip[ci][a:b]
            off += n_w
    assert seen == nw
    # (d) both residency modes and a single compute stream
    z4, i4, *_ = run_genome(chroms, env={"GB_GENOME_RESIDENT": "pack5", "GB_GENOME_STREAMS": "1", "GB_GENOME_BATCH_WINDOWS": "3"})
    z5, i5, _, _, inf5, _ = run_genome(chroms, env={"GB_GENOME_RESIDENT": "e2m1"})
    assert inf5[0]["e2m1_resident"]
    # hybrid residency (what a whole genome on one GPU gets): a budget that keeps only some batches expanded, with
    # generated rows and with uploaded rows
    hyb = {"GB_GENOME_EXPANDED_GB": "0.0008", "GB_GENOME_BATCH_WINDOWS": "2"}
    z6, i6, _, _, inf6, _ = run_genome(chroms, env=hyb)
    assert 0 < inf6[0]["expanded_pct"] < 100
    z7, i7, *_ = run_genome(chroms, host_rows=host, env=hyb)
    # (e) both schedules: the two-lane software pipeline (default) against per-batch forks on two alternating streams,
    # and the fp64 triangular solve within its distance of the int8-split one
    z8, i8, *_ = run_genome(chroms, env={"GB_GENOME_CHAIN_SMS": "0", "GB_GENOME_BATCH_WINDOWS": "2"})
    z9, i9, *_ = run_genome(chroms, env={"GB_GENOME_CHAIN_SMS": "48", "GB_GENOME_BATCH_WINDOWS": "2"})
    zf, jf, *_ = run_genome(chroms, env={"GB_SOLVE": "fp64"})
    for ci in range(len(chroms)):
        ok = ~np.isnan(z1[ci])
        assert np.array_equal(np.isnan(zf[ci]), ~ok)
        assert np.abs(zf[ci][ok] - z1[ci][ok]).max() <= 1e-10 and np.abs(jf[ci][ok] - i1[ci][ok]).max() <= 1e-10
    for ci in range(len(chroms)):
        for other_z, other_i in ((z2, i2), (z3, i3), (z4, i4), (z5, i5), (z6, i6), (z7, i7), (z8, i8), (z9, i9)):
            assert np.array_equal(z1[ci], other_z[ci], equal_nan=True)
            assert np.array_equal(i1[ci], other_i[ci], equal_nan=True)


@pytest.mark.gpu
def test_genome_on_several_gpus_is_byte_identical(gpu_ctx):
    import torch
    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs 2+ GPUs")
    chroms = tiny_genome()
    z1, i1, s1, *_ = run_genome(chroms)
    for n in [k for k in (2, 4, 8) if k <= n_dev]:
        zn, in_, sn, ms, infos, _ = run_genome(chroms, n_gpus=n)
        assert sum(x["n_windows"] for x in infos) == sum(len(c["t_off"]) - 1 for c in chroms)
        assert all(m > 0 for m in ms)
        for ci in range(len(chroms)):
            assert np.array_equal(z1[ci], zn[ci], equal_nan=True) and np.array_equal(i1[ci], in_[ci], equal_nan=True)
            assert np.array_equal(s1[ci], sn[ci])
