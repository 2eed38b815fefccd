"""Parity tests proper: the CUDA path, called through the C-ABI, against the CPU oracle and the
golden vectors minted from the reference's own code.  Bars (BASELINE.json north_star):
Gram counts bit-exact; correlations, imputed Z and info within 1e-6 absolute (fp64)."""
import glob
import os

import numpy as np
import pytest

import gauss_b200 as gb
from gauss_b200 import synth
from helpers import small_case, split_rows

pytestmark = pytest.mark.gpu

TOL = 1e-6        # the tolerance north_star states
TIGHT = 1e-10     # what fp64 Cholesky-vs-LU actually achieves on these well-conditioned windows
GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("pgc2_"))


FORMATS = ["e2m1", "int8"]   # enum gb_panel_format: both feed the same kernel and must agree bit for bit


def make_panel(ctx, g, pop_sizes, mode="i8", fmt=None):
    p = gb.Panel(ctx, pop_sizes, len(g), fmt)
    if mode == "i8":
        p.append_host(g.astype(np.int8), is_ascii=False)
    elif mode == "ascii":
        p.append_host((g.astype(np.int16) + 48).astype(np.uint8), is_ascii=True)
    assert p.n_rows == len(g)
    return p


# ------------------------------------------------------------------ integer surface: bit-exact
@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("mode", ["i8", "ascii"])
def test_gram_counts_bit_exact(gpu_ctx, oracle, mode, fmt):
    c = small_case(seed=21, n_snps=330, pop_sizes=(61, 103, 40, 25, 2, 330, 97, 128, 31, 33))
    panel = make_panel(gpu_ctx, c["g"], c["pop_sizes"], mode, fmt)
    assert panel.format == fmt
    rows_a = np.arange(0, 200)
    rows_b = np.arange(193, 330)           # ragged: 200 x 137, overlapping ranges
    sxy, sx, sxx = panel.gram_counts(rows_a, rows_b)
    o_sxy, o_sx, o_sxx = oracle.gram_counts(c["g"][rows_a], c["g"][rows_b], c["pop_sizes"])
    np.testing.assert_array_equal(sxy, o_sxy)
    np.testing.assert_array_equal(sx, o_sx)
    np.testing.assert_array_equal(sxx, o_sxx)


def test_gram_counts_gathered_rows(gpu_ctx, oracle):
    """Non-contiguous row lists go through the gather-to-scratch path."""
    c = small_case(seed=22, n_snps=300)
    panel = make_panel(gpu_ctx, c["g"], c["pop_sizes"])
    rng = np.random.default_rng(0)
    rows_a = np.sort(rng.choice(300, 150, replace=False))
    rows_b = rng.permutation(300)[:70]
    sxy, _, _ = panel.gram_counts(rows_a, rows_b)
    o_sxy, _, _ = oracle.gram_counts(c["g"][rows_a], c["g"][rows_b], c["pop_sizes"])
    np.testing.assert_array_equal(sxy, o_sxy)


@pytest.mark.parametrize("fmt", FORMATS)
def test_gram_counts_33kg_shape(gpu_ctx, fmt):
    """Full K extent (32,147 individuals, 21 populations): exactness via numpy int64 matmul.  Rows of
    all-2 / all-1 dosages drive the per-population counts to their maximum (4 x 6,360 = 25,440)."""
    _, sizes, _ = synth.flagged_33kg_pgc2()
    g = synth.make_genotypes(160, sizes, seed=3)
    g[31], g[32], g[33] = 2, 1, 2
    panel = make_panel(gpu_ctx, g, sizes, fmt=fmt)
    sxy, sx, sxx = panel.gram_counts(np.arange(0, 130), np.arange(30, 160))
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for p in range(len(sizes)):
        a = g[0:130, offs[p]:offs[p + 1]].astype(np.int32)
        b = g[30:160, offs[p]:offs[p + 1]].astype(np.int32)
        np.testing.assert_array_equal(sxy[p], a @ b.T)
        np.testing.assert_array_equal(sx[p], a.sum(1))
        np.testing.assert_array_equal(sxx[p], (a * a).sum(1))


# ------------------------------------------------------------------ correlations
@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("mix", [True, False])
def test_correlation_blocks_match_oracle(gpu_ctx, oracle, mix, fmt):
    c = small_case(seed=23, n_snps=420, pop_sizes=(33, 129, 500, 64, 7))
    meas, unme = split_rows(c)
    w = c["w"] if mix else None
    panel = make_panel(gpu_ctx, c["g"], c["pop_sizes"], fmt=fmt)
    B11, B21 = panel.window_cor(meas, unme, w)
    r = oracle.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"], dump=True)
    assert np.abs(B11 - r["B11"]).max() <= 1e-12
    assert np.abs(B21 - r["B21"]).max() <= 1e-12
    if fmt == "int8" or not mix:
        # int8 panels (and the pooled r of dist() on any panel) keep the reference's fp64 operation
        # order term by term: entries are bit-identical to CalWgtCov / CalCor
        assert np.abs(B11 - r["B11"]).max() <= 1e-13 and np.abs(B21 - r["B21"]).max() <= 1e-13
        assert (B21 == r["B21"]).mean() > 0.999
        assert (B11 == r["B11"]).mean() > 0.999
    np.testing.assert_array_equal(np.diag(B11), np.full(len(meas), 1.1))


# ------------------------------------------------------------------ imputation
@pytest.mark.parametrize("mix", [True, False])
def test_window_imputation_matches_oracle(gpu_ctx, oracle, mix):
    c = small_case(seed=24, n_snps=500, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.4, core=(30, 470))
    meas, unme = split_rows(c)
    w = c["w"] if mix else None
    panel = make_panel(gpu_ctx, c["g"], c["pop_sizes"])
    if mix:
        z, info, rc = panel.window_distmix(meas, unme, c["z"][meas], w)
    else:
        z, info, rc = panel.window_dist(meas, unme, c["z"][meas])
    assert rc == gb.GB_OK
    r = oracle.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    assert r["rc"] == 0 and r["n_t"] == len(meas) and r["n_u"] == len(unme)
    assert np.abs(z - r["z"][unme]).max() <= TOL
    assert np.abs(info - r["info"][unme]).max() <= TOL
    assert np.abs(z - r["z"][unme]).max() <= TIGHT
    assert np.abs(info - r["info"][unme]).max() <= TIGHT


@pytest.mark.parametrize("fmt", FORMATS)
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_against_golden_reference_vectors(gpu_ctx, path, fmt):
    d = np.load(path)
    panel = make_panel(gpu_ctx, d["g"], d["pop_sizes"], fmt=fmt)
    meas, unme = d["meas"], d["unme"]
    z, info, _ = panel.window_distmix(meas, unme, d["z"][meas], d["w"])
    assert np.abs(z - d["mix_z"][unme]).max() <= TOL and np.abs(info - d["mix_info"][unme]).max() <= TOL
    z, info, _ = panel.window_dist(meas, unme, d["z"][meas])
    assert np.abs(z - d["dist_z"][unme]).max() <= TOL and np.abs(info - d["dist_info"][unme]).max() <= TOL
    ld, _ = panel.window_ld(meas, d["w"])
    assert np.abs(ld - d["ld"]).max() <= TOL
    np.testing.assert_array_equal(np.diag(ld), np.ones(len(meas)))
    np.testing.assert_array_equal(ld, ld.T)


def test_formats_agree_bitwise_and_pooled_counts_reach_17_bits(gpu_ctx, oracle):
    """dist() pools all 32,147 individuals into ONE accumulation (counts up to 4 x 32,147 = 128,588):
    the fp32 accumulator of the E2M1 path must still hold exact integers, i.e. give the very same
    doubles as the int8 path, and both must match CalCor."""
    _, sizes, w = synth.flagged_33kg_pgc2()
    g = synth.make_genotypes(200, sizes, seed=31)
    g[0], g[1] = 2, 2
    g[1, ::7] = 1
    g[0, 5] = 1
    meas, unme = np.arange(0, 60), np.arange(60, 200)
    out = {}
    for fmt in FORMATS:
        panel = make_panel(gpu_ctx, g, sizes, fmt=fmt)
        out[fmt] = panel.window_cor(meas, unme, None) + panel.window_cor(meas, unme, w)
    for a, b in zip(out["e2m1"][:2], out["int8"][:2]):      # pooled r: same integers -> same doubles
        np.testing.assert_array_equal(a, b)
    # mixture: regrouped vs literal term order.  Rows 0 and 1 are almost monomorphic at dosage 2 (allele
    # frequency 0.99998, far outside the reference's 0.01..0.99 AF filter): sums of ~1e8 cancel down to a
    # variance of ~1e-4, the worst case for the regrouped fold -- still 5 orders inside the 1e-6 bar.
    for a, b in zip(out["e2m1"][2:], out["int8"][2:]):
        assert np.abs(a - b).max() <= 1e-10
        assert np.abs(a[2:, 2:] - b[2:, 2:]).max() <= 1e-12 if a.shape[0] == a.shape[1] else np.abs(a - b)[:, 2:].max() <= 1e-12
    assert oracle.cal_cor(g[0], g[1], sizes) == out["int8"][0][0, 1]
    assert abs(oracle.cal_cor(g[3], g[100], sizes) - out["e2m1"][1][40, 3]) <= 1e-15


def test_unrepresentable_dosage_needs_int8(gpu_ctx, oracle):
    """The reference computes (c - '0') for ANY byte.  int8 panels reproduce that; an E2M1 panel must
    refuse (not silently round) and the string-level mirror must repack by itself."""
    c = small_case(seed=32, n_snps=160, pop_sizes=(50, 7, 211, 96))
    g = c["g"].copy()
    g[3, 10], g[90, 200] = 5, 7          # '5' and '7' in the genotype strings
    rows_t, rows_u = np.arange(0, 40), np.arange(40, 160)
    p4 = make_panel(gpu_ctx, g, c["pop_sizes"], fmt="e2m1")
    with pytest.raises(gb.GaussB200Error) as e:
        p4.window_distmix(rows_t, rows_u, c["z"][rows_t], c["w"])
    assert e.value.status == gb.api.GB_ERR_UNSUPPORTED
    p8 = make_panel(gpu_ctx, g, c["pop_sizes"], fmt="int8")
    sxy, sx, sxx = p8.gram_counts(np.arange(0, 100), np.arange(60, 160))
    o = oracle.gram_counts(g[0:100], g[60:160], c["pop_sizes"])
    for a, b in zip((sxy, sx, sxx), o):
        np.testing.assert_array_equal(a, b)
    # values E2M1 does hold exactly besides 0/1/2: 3, 4, 6
    g2 = c["g"].copy()
    g2[3, 10], g2[90, 200], g2[91, 3] = 3, 4, 6
    sxy, _, _ = make_panel(gpu_ctx, g2, c["pop_sizes"], fmt="e2m1").gram_counts(np.arange(0, 100), np.arange(60, 160))
    np.testing.assert_array_equal(sxy, oracle.gram_counts(g2[0:100], g2[60:160], c["pop_sizes"])[0])
    # the host mirror of run_distmix gets strings and picks the format itself
    offs = np.concatenate([[0], np.cumsum(c["pop_sizes"])])
    chars = (g.astype(np.int16) + 48).astype(np.uint8)
    strings = [[chars[i, offs[p]:offs[p + 1]].tobytes() for p in range(len(c["pop_sizes"]))] for i in range(len(g))]
    out = gpu_ctx.run_window_strings(c["type"][:160], c["bp"][:160], c["z"][:160], np.ones(160), strings,
                                     c["pop_sizes"], c["w"], c["start_bp"], c["end_bp"])
    r = oracle.run_window(c["type"][:160], c["bp"][:160], c["z"][:160], g, c["pop_sizes"], c["w"],
                          c["start_bp"], c["end_bp"])
    assert out["rc"] == r["rc"] == 0
    assert np.abs(out["z"] - r["z"]).max() <= TOL and np.abs(out["info"] - r["info"]).max() <= TOL


@pytest.mark.parametrize("fmt", FORMATS)
def test_compute_ld_matches_oracle(gpu_ctx, oracle, fmt):
    c = small_case(seed=25, n_snps=300, pop_sizes=(90, 40, 260))
    panel = make_panel(gpu_ctx, c["g"], c["pop_sizes"], fmt=fmt)
    rows = np.arange(17, 290)
    ld, _ = panel.window_ld(rows, c["w"])
    ref = oracle.compute_ld(c["g"][rows], c["pop_sizes"], c["w"])
    assert np.abs(ld - ref).max() <= 1e-12
    if fmt == "int8":
        assert np.abs(ld - ref).max() <= 1e-13 and (ld == ref).mean() > 0.999


def test_host_mirror_of_run_distmix_strings(gpu_ctx, oracle):
    """gb_run_window_strings takes what run_distmix takes: bp-sorted SNPs with type, z and one
    genotype string per population (snp.h:109)."""
    c = small_case(seed=26, n_snps=260, pop_sizes=(50, 7, 211, 96))
    offs = np.concatenate([[0], np.cumsum(c["pop_sizes"])])
    chars = (c["g"].astype(np.int16) + 48).astype(np.uint8)
    strings = [[chars[i, offs[p]:offs[p + 1]].tobytes() for p in range(len(c["pop_sizes"]))]
               for i in range(len(chars))]
    for w in (c["w"], None):
        out = gpu_ctx.run_window_strings(c["type"], c["bp"], c["z"], np.ones(len(chars)), strings, c["pop_sizes"],
                                         w, c["start_bp"], c["end_bp"])
        r = oracle.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
        assert out["rc"] == 0 and out["n_t"] == r["n_t"] and out["n_u"] == r["n_u"]
        assert np.abs(out["z"] - r["z"]).max() <= TOL       # measured / out-of-window SNPs untouched
        assert np.abs(out["info"] - r["info"]).max() <= TOL


def test_too_few_snps_status_codes(gpu_ctx):
    c = small_case(seed=27, n_snps=100)
    panel = make_panel(gpu_ctx, c["g"], c["pop_sizes"])
    z, info, rc = panel.window_distmix(np.arange(10), np.arange(20, 60), np.zeros(10), c["w"],
                                       allow=(gb.api.GB_ERR_TOO_FEW_MEASURED,))
    assert rc == gb.api.GB_ERR_TOO_FEW_MEASURED          # n_t <= 10, distmix.cpp:154
    z, info, rc = panel.window_dist(np.arange(30), np.arange(40, 50), np.zeros(30),
                                    allow=(gb.api.GB_ERR_TOO_FEW_UNMEASURED,))
    assert rc == gb.api.GB_ERR_TOO_FEW_UNMEASURED        # n_u <= 10
    _, rc = panel.window_ld(np.arange(10), c["w"], allow=(gb.api.GB_ERR_TOO_FEW_MEASURED,))
    assert rc == gb.api.GB_ERR_TOO_FEW_MEASURED          # computeLD.cpp:89


def test_make_pos_def_clip_on_the_device(gpu_ctx, oracle):
    """lambda = 0 with a duplicated measured SNP makes B11 singular: the reference's MakePosDef (util.cpp:302-318) clips
    the spectrum at min_abs_eig and carries on.  The Cholesky certificate cannot vouch for such a window, so it takes
    the device eigen-clip path and must come back with the reference's numbers, status GB_OK."""
    c = small_case(seed=28, n_snps=120)
    g = c["g"].copy()
    g[5] = g[4]
    panel = make_panel(gpu_ctx, g, c["pop_sizes"])
    p = gb.Params.default()
    p.lambda_ = 0.0
    meas, unme = np.arange(0, 40), np.arange(40, 120)
    zt = np.asarray(c["z"])[:40]
    t = np.concatenate([np.ones(40, np.int32), np.zeros(80, np.int32)])
    for w in (c["w"], None):
        if w is None:
            z, info, rc = panel.window_dist(meas, unme, zt, p)
        else:
            z, info, rc = panel.window_distmix(meas, unme, zt, w, p)
        assert rc == gb.GB_OK
        r = oracle.run_window(t, np.arange(120, dtype=np.int64), np.concatenate([zt, np.zeros(80)]), g, c["pop_sizes"], w,
                              0, 10 ** 12, lam=0.0)
        assert r["rc"] == 0
        assert np.abs(z - r["z"][40:]).max() <= TOL and np.abs(info - r["info"][40:]).max() <= TOL
    # the same window inside a batch with healthy windows: only it is repaired, the others keep their bits
    t_off, u_off = [0, 40, 70], [0, 80, 130]
    rows_t = np.concatenate([meas, np.arange(10, 40)])
    rows_u = np.concatenate([unme, np.arange(60, 110)])
    zt2 = np.concatenate([zt, zt[10:40]])
    batch = gb.Batch(panel, t_off, rows_t, u_off, rows_u, zt2, c["w"], p)
    batch.run()
    zb, ib, st = batch.fetch()
    assert (st == 0).all()
    z1, i1, _ = panel.window_distmix(meas, unme, zt, c["w"], p)
    z2, i2, _ = panel.window_distmix(np.arange(10, 40), np.arange(60, 110), zt[10:40], c["w"], p)
    np.testing.assert_array_equal(zb, np.concatenate([z1, z2]))
    np.testing.assert_array_equal(ib, np.concatenate([i1, i2]))
    # a monomorphic SNP has zero variance -> NaN correlations (no guard at distmix.cpp:196): nothing to clip, the
    # factorisation breaks down, the window reports it and every result is NaN (the reference returns NaN too)
    BREAKDOWN = 9
    g2 = c["g"].copy()
    g2[7] = 0
    panel2 = make_panel(gpu_ctx, g2, c["pop_sizes"])
    z, info, rc = panel2.window_distmix(meas, unme, np.zeros(40), c["w"], allow=(BREAKDOWN,))
    assert rc == BREAKDOWN and np.isnan(z).all() and np.isnan(info).all()


def test_batch_equals_single_windows_and_placement_invariance(gpu_ctx, oracle):
    """Several windows of different sizes in one batch; results must not depend on batching, on
    the row layout (interleaved vs measured/unmeasured blocks) or on which launch computed them."""
    c = small_case(seed=29, n_snps=900, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(0, 900))
    g, t = c["g"], c["type"]
    panel = make_panel(gpu_ctx, g, c["pop_sizes"])
    wins = [(0, 300), (150, 520), (400, 900), (880, 900), (300, 700)]
    t_rows, u_rows, t_off, u_off = [], [], [0], [0]
    for lo, hi in wins:
        idx = np.arange(lo, hi)
        t_rows.append(idx[t[lo:hi] == 1])
        u_rows.append(idx[t[lo:hi] == 0])
        t_off.append(t_off[-1] + len(t_rows[-1]))
        u_off.append(u_off[-1] + len(u_rows[-1]))
    rows_t, rows_u = np.concatenate(t_rows), np.concatenate(u_rows)
    batch = gb.Batch(panel, t_off, rows_t, u_off, rows_u, c["z"][rows_t], c["w"])
    batch.run()
    z, info, status = batch.fetch()
    assert status[3] in (gb.api.GB_ERR_TOO_FEW_MEASURED, gb.api.GB_ERR_TOO_FEW_UNMEASURED)
    assert np.isnan(z[u_off[3]:u_off[4]]).all()
    for k, (lo, hi) in enumerate(wins):
        if k == 3:
            continue
        assert status[k] == gb.GB_OK
        z1, i1, _ = panel.window_distmix(t_rows[k], u_rows[k], c["z"][t_rows[k]], c["w"])
        np.testing.assert_array_equal(z[u_off[k]:u_off[k + 1]], z1)
        np.testing.assert_array_equal(info[u_off[k]:u_off[k + 1]], i1)
    # same SNPs re-packed as [measured block | unmeasured block]: contiguous TMA path, no gather
    order = np.concatenate([np.where(t == 1)[0], np.where(t == 0)[0]])
    inv = np.empty_like(order)
    inv[order] = np.arange(len(order))
    panel2 = make_panel(gpu_ctx, g[order], c["pop_sizes"])
    k = 1
    z2, i2, _ = panel2.window_distmix(inv[t_rows[k]], inv[u_rows[k]], c["z"][t_rows[k]], c["w"])
    np.testing.assert_array_equal(z[u_off[k]:u_off[k + 1]], z2)
    np.testing.assert_array_equal(info[u_off[k]:u_off[k + 1]], i2)
    # and the oracle agrees on one of them
    lo, hi = wins[1]
    r = oracle.run_window(t[lo:hi], c["bp"][lo:hi], c["z"][lo:hi], g[lo:hi], c["pop_sizes"], c["w"], 0, 10**15)
    assert np.abs(r["z"][t[lo:hi] == 0] - z[u_off[1]:u_off[2]]).max() <= TIGHT


def test_33kg_shaped_window_matches_oracle(gpu_ctx, oracle):
    """BASELINE config #2 shape at a size the oracle finishes in seconds: 21 flagged populations,
    32,147 individuals, PGC2 ancestry weights (sum 1.061, not renormalised)."""
    _, sizes, w = synth.flagged_33kg_pgc2()
    n = 520
    g = synth.make_genotypes(n, sizes, seed=30)
    rng = np.random.default_rng(30)
    t = (rng.random(n) < 0.35).astype(np.int32)
    bp = np.arange(n, dtype=np.int64) * 300 + 1
    zin = rng.standard_normal(n) * 1.34
    meas, unme = np.where(t == 1)[0], np.where(t == 0)[0]
    panel = make_panel(gpu_ctx, g, sizes)
    z, info, rc = panel.window_distmix(meas, unme, zin[meas], w)
    assert rc == gb.GB_OK
    r = oracle.run_window(t, bp, zin, g, sizes, w, 0, 10**12)
    assert np.abs(z - r["z"][unme]).max() <= TOL and np.abs(info - r["info"][unme]).max() <= TOL
    assert np.abs(z - r["z"][unme]).max() <= 1e-9
    # size-independent properties: info in (0, 1], flipping the sign of Z flips the imputed z
    assert (info > 0).all() and (info <= 1.0 + 1e-12).all()
    z_neg, info_neg, _ = panel.window_distmix(meas, unme, -zin[meas], w)
    np.testing.assert_array_equal(z_neg, -z)
    np.testing.assert_array_equal(info_neg, info)
    # linearity in Z1 (dist.cpp:194): z(a + b) == z(a) + z(b) up to rounding
    za, _, _ = panel.window_distmix(meas, unme, np.ones(len(meas)), w)
    zb, _, _ = panel.window_distmix(meas, unme, zin[meas] + 1.0, w)
    assert np.abs(zb - (z + za)).max() <= 1e-9


def test_pipe_matches_single_window_calls(gpu_ctx, oracle):
    """gb_pipe_submit / gb_pipe_wait: host rows in, host results out, copies overlapped with compute.
    Results must be identical to the blocking per-window call, in any submission order, and the
    reference's refusals must come back as the window status."""
    c = small_case(seed=33, n_snps=900, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(0, 900))
    g, t = c["g"].astype(np.int8), c["type"]
    panel = make_panel(gpu_ctx, g, c["pop_sizes"])
    wins = [(0, 300), (150, 520), (400, 900), (880, 900), (300, 700), (100, 400), (0, 900)]
    pipe = gb.Pipe(gpu_ctx, c["pop_sizes"], 900, depth=2)
    tickets, outs = [], []
    for lo, hi in wins:
        idx = np.arange(lo, hi)
        rt, ru = idx[t[lo:hi] == 1], idx[t[lo:hi] == 0]
        tk, z, info = pipe.submit(g[rt], g[ru], c["z"][rt], c["w"])
        tickets.append(tk)
        outs.append((rt, ru, z, info))
        if len(tickets) >= 2:                       # wait lags one window behind, like a genome loop
            k = len(tickets) - 2
            st = pipe.wait(tickets[k])
            assert (st != 0) == (k == 3)
    assert pipe.wait(tickets[-1]) == 0
    for k, (rt, ru, z, info) in enumerate(outs):
        if k == 3:
            assert np.isnan(z).all()
            continue
        z1, i1, _ = panel.window_distmix(rt, ru, c["z"][rt], c["w"])
        np.testing.assert_array_equal(z, z1)
        np.testing.assert_array_equal(info, i1)
    # dist() through the pipe, ASCII rows
    lo, hi = wins[1]
    idx = np.arange(lo, hi)
    rt, ru = idx[t[lo:hi] == 1], idx[t[lo:hi] == 0]
    chars = (g.astype(np.int16) + 48).astype(np.uint8)
    tk, z, info = pipe.submit(chars[rt], chars[ru], c["z"][rt], None)
    assert pipe.wait(tk) == 0
    z1, i1, _ = panel.window_dist(rt, ru, c["z"][rt])
    np.testing.assert_array_equal(z, z1)
    # an unrepresentable dosage in an E2M1 pipe is reported at wait time
    g2 = g.copy()
    g2[rt[0], 3] = 5
    tk, z, info = pipe.submit(g2[rt], g2[ru], c["z"][rt], c["w"])
    assert pipe.wait(tk) == gb.api.GB_ERR_UNSUPPORTED
    pipe8 = gb.Pipe(gpu_ctx, c["pop_sizes"], 900, depth=1, fmt="int8")
    tk, z, info = pipe8.submit(g2[rt], g2[ru], c["z"][rt], c["w"])
    assert pipe8.wait(tk) == 0
    r = oracle.run_window(np.concatenate([np.ones(len(rt), np.int32), np.zeros(len(ru), np.int32)]),
                          np.arange(len(rt) + len(ru), dtype=np.int64), np.concatenate([c["z"][rt], np.zeros(len(ru))]),
                          np.concatenate([g2[rt], g2[ru]]), c["pop_sizes"], c["w"], 0, 10**12)
    assert np.abs(z - r["z"][len(rt):]).max() <= TIGHT


@pytest.mark.parametrize("solver", ["int8", "fp64"])
def test_results_do_not_depend_on_uninitialised_memory(solver, monkeypatch):
    """GB_POISON fills every transient buffer of a batch with 0xFF bytes (NaNs as doubles, -1 digits) before any
    kernel writes it.  Padding that is read but never written -- row n_t of an odd-sized B11, the columns between
    n_u and its multiple of 8 -- must not reach a result: stream-ordered allocations hand back memory full of
    whatever the previous batch left there."""
    c = small_case(seed=33, n_snps=900, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(0, 900))
    g, t = c["g"].astype(np.int8), c["type"]
    monkeypatch.setenv("GB_SOLVE", solver)
    ctx = gb.Context(0)
    try:
        panel = make_panel(ctx, g, c["pop_sizes"])
        for lo, hi in [(0, 300), (150, 520), (0, 900)]:
            idx = np.arange(lo, hi)
            rt, ru = idx[t[lo:hi] == 1], idx[t[lo:hi] == 0]
            monkeypatch.delenv("GB_POISON", raising=False)
            z0, i0, _ = panel.window_distmix(rt, ru, c["z"][rt], c["w"])
            for byte in (0xFF, 0x7F):
                monkeypatch.setenv("GB_POISON", "1023:%d" % byte)
                z1, i1, _ = panel.window_distmix(rt, ru, c["z"][rt], c["w"])
                np.testing.assert_array_equal(z1, z0)
                np.testing.assert_array_equal(i1, i0)
                z2, i2, _ = panel.window_dist(rt, ru, c["z"][rt])
                assert np.isfinite(z2).all() and np.isfinite(i2).all()
    finally:
        ctx.close()


@pytest.mark.parametrize("fmt", FORMATS)
def test_monomorphic_unmeasured_snp_gives_nan_under_both_solvers(oracle, monkeypatch, fmt):
    """An unmeasured SNP without variation has sd = 0: its correlations are 0/0, and the reference's z and info
    for it are NaN (dist.cpp:196-200).  The int8-split solve carries B21 as integer digit planes, which have no NaN:
    the finish pass flags the row instead and the solve returns NaN for it -- and only for it."""
    c = small_case(seed=41, n_snps=420, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(0, 420))
    g, t = c["g"].astype(np.int8).copy(), c["type"]
    idx = np.arange(420)
    rt, ru = idx[t == 1], idx[t == 0]
    g[ru[5]] = 0
    g[ru[140]] = 2
    res = {}
    for solver in ("int8", "fp64"):
        monkeypatch.setenv("GB_SOLVE", solver)
        ctx = gb.Context(0)
        try:
            panel = make_panel(ctx, g, c["pop_sizes"], fmt=fmt)
            res[solver, "mix"] = panel.window_distmix(rt, ru, c["z"][rt], c["w"])[:2]
            res[solver, "pooled"] = panel.window_dist(rt, ru, c["z"][rt])[:2]
        finally:
            ctx.close()
    r = oracle.run_window(t, c["bp"], c["z"], g, c["pop_sizes"], c["w"], 0, 10**15)
    for mode in ("mix", "pooled"):
        for solver in ("int8", "fp64"):
            z, info = res[solver, mode]
            bad = np.isnan(z)
            assert bad.sum() == 2 and bad[5] and bad[140]
            assert np.array_equal(np.isnan(info), bad)
        good = ~np.isnan(res["fp64", mode][0])
        assert np.abs(res["int8", mode][0][good] - res["fp64", mode][0][good]).max() <= TIGHT
        assert np.abs(res["int8", mode][1][good] - res["fp64", mode][1][good]).max() <= TIGHT
    zo = r["z"][t == 0]
    assert np.isnan(zo[5]) and np.isnan(zo[140])
    good = ~np.isnan(zo)
    assert np.abs(res["int8", "mix"][0][good] - zo[good]).max() <= TIGHT


def test_solve_gemm_variants_agree(gpu_ctx, monkeypatch):
    """GB_OZ_KERNEL=2 selects the all-groups-live variant of the digit-plane GEMM (kept as a measured alternative): the
    same exact integer products, recombined in fp64 in another order -- equal to rounding."""
    c = small_case(seed=47, n_snps=900, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.35, core=(0, 900))
    g, t = c["g"].astype(np.int8), c["type"]
    panel = make_panel(gpu_ctx, g, c["pop_sizes"])
    for lo, hi in [(0, 300), (120, 900), (0, 900)]:
        idx = np.arange(lo, hi)
        rt, ru = idx[t[lo:hi] == 1], idx[t[lo:hi] == 0]
        monkeypatch.delenv("GB_OZ_KERNEL", raising=False)
        z1, i1, _ = panel.window_distmix(rt, ru, c["z"][rt], c["w"])
        monkeypatch.setenv("GB_OZ_KERNEL", "2")
        z2, i2, _ = panel.window_distmix(rt, ru, c["z"][rt], c["w"])
        monkeypatch.delenv("GB_OZ_KERNEL", raising=False)
        assert np.abs(z2 - z1).max() <= 1e-12 and np.abs(i2 - i1).max() <= 1e-12
        assert not np.array_equal(z2, np.zeros_like(z2))


def test_windows_beyond_the_digit_plane_limit_take_the_fp64_solve(monkeypatch):
    """More than 2,048 measured SNPs in a window: the int8-split solve's accumulator bound (6 pairs x 128^2 x K) and its
    shared-memory copy of y end there, so the batch is solved by the fp64 triangular solve -- same numbers as with
    GB_SOLVE=fp64, bit for bit."""
    c = small_case(seed=53, n_snps=2500, pop_sizes=(300, 330), measured_frac=0.86, core=(0, 2500))
    g, t = c["g"].astype(np.int8), c["type"]
    idx = np.arange(2500)
    rt, ru = idx[t == 1], idx[t == 0]
    assert len(rt) > 2048 and len(ru) > 10
    res = {}
    for solver in ("int8", "fp64"):
        monkeypatch.setenv("GB_SOLVE", solver)
        ctx = gb.Context(0)
        try:
            panel = make_panel(ctx, g, c["pop_sizes"])
            res[solver] = panel.window_distmix(rt, ru, c["z"][rt], c["w"])[:2]
        finally:
            ctx.close()
    assert np.isfinite(res["int8"][0]).all() and np.isfinite(res["int8"][1]).all()
    np.testing.assert_array_equal(res["int8"][0], res["fp64"][0])
    np.testing.assert_array_equal(res["int8"][1], res["fp64"][1])


def test_overlapped_batch_run_equals_staged_run(gpu_ctx):
    """gb_batch_run puts the factorisation on a side stream beside the B21 Gram tiles once a batch has at least one
    B21 tile per SM; the results must be the bits of the stage-by-stage run, run after run."""
    c = small_case(seed=81, n_snps=7200, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.25, core=(0, 7200))
    g, t = c["g"].astype(np.int8), c["type"]
    panel = make_panel(gpu_ctx, g, c["pop_sizes"])
    t_rows, u_rows, t_off, u_off = [], [], [0], [0]
    for lo, hi in [(0, 1900), (1500, 3600), (3300, 5400), (5000, 7200), (100, 124)]:
        idx = np.arange(lo, hi)
        t_rows.append(idx[t[lo:hi] == 1])
        u_rows.append(idx[t[lo:hi] == 0])
        t_off.append(t_off[-1] + len(t_rows[-1]))
        u_off.append(u_off[-1] + len(u_rows[-1]))
    n_b21_tiles = sum(-(-len(a) // 128) * -(-len(b) // 128) for a, b in zip(t_rows[:4], u_rows[:4]))
    assert n_b21_tiles >= 160          # more than the 148 SMs: the overlapped schedule is taken
    rows_t, rows_u = np.concatenate(t_rows), np.concatenate(u_rows)
    batch = gb.Batch(panel, t_off, rows_t, u_off, rows_u, c["z"][rows_t], c["w"])
    for s in (0, 1, 2, 3):
        batch.run_stage(s)
    z0, i0, s0 = batch.fetch()
    assert (s0[:4] == 0).all() and s0[4] != 0 and np.isfinite(z0[:u_off[4]]).all()
    for _ in range(3):
        batch.run()
        z1, i1, s1 = batch.fetch()
        np.testing.assert_array_equal(z1, z0)
        np.testing.assert_array_equal(i1, i0)
        np.testing.assert_array_equal(s1, s0)


def test_e2m1_count_range_guard(gpu_ctx, oracle):
    """E2M1 panels count in fp32 and re-encode the counts with one shift-add, exact below 2^21: a population block that
    could exceed it (36 per individual with the largest dosage the format holds) is refused, int8 takes it."""
    sizes = np.array([60000, 50], np.int32)
    g = synth.make_genotypes(40, sizes, seed=83).astype(np.int8)
    w = np.array([0.7, 0.3])
    t = np.array([1, 0] * 20, np.int32)
    meas, unme = np.where(t == 1)[0], np.where(t == 0)[0]
    z = np.linspace(-2, 2, 40)
    p4 = gb.Panel(gpu_ctx, sizes, 40, "e2m1")
    p4.append_host(g, is_ascii=False)
    _, _, rc = p4.window_distmix(meas, unme, z[meas], w, allow=(gb.api.GB_ERR_UNSUPPORTED,))
    assert rc == gb.api.GB_ERR_UNSUPPORTED
    zp, ip, rc = p4.window_dist(meas, unme, z[meas])        # pooled: 36 * 60,050 is still below 2^22
    assert rc == gb.GB_OK
    p8 = gb.Panel(gpu_ctx, sizes, 40, "int8")
    p8.append_host(g, is_ascii=False)
    z8, i8, rc = p8.window_distmix(meas, unme, z[meas], w)
    assert rc == gb.GB_OK
    r = oracle.run_window(t, np.arange(40, dtype=np.int64), z, g, sizes, w, 0, 10**9)
    assert np.abs(z8 - r["z"][unme]).max() <= TOL and np.abs(i8 - r["info"][unme]).max() <= TOL
    rp = oracle.run_window(t, np.arange(40, dtype=np.int64), z, g, sizes, None, 0, 10**9)
    assert np.abs(zp - rp["z"][unme]).max() <= TOL
