"""BASELINE.json configurations at their real shapes.  Where the oracle's O(n^2 N) loops would take
minutes, the CUDA path is checked on randomly sampled entries against the oracle's CalWgtCov / CalCor
and through size-independent properties (symmetry, unit diagonal, |r| <= 1, the normal equations)."""
import os

import numpy as np
import pytest

import gauss_b200 as gb
from gauss_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-6
SITES = os.path.join(os.path.dirname(__file__), "golden", "pgc2_chr22_sites.npz")


def mix_cor(oracle, gi, gj, sizes, w):
    return oracle.cal_wgt_cov(gi, gj, sizes, w) / np.sqrt(oracle.cal_wgt_cov(gi, gi, sizes, w) *
                                                          oracle.cal_wgt_cov(gj, gj, sizes, w))


@pytest.mark.parametrize("study_pop", ["EUR", "ALL"])
def test_cfg1_dist_1kg_shaped_window(gpu_ctx, oracle, study_pop):
    """configs[0]: dist() on a chr22 window of the bundled PGC2 positions vs a 1KG-shaped panel:
    study_pop = "EUR" flags 503 of 2,504 individuals (5 populations); "ALL" pools all 26."""
    sizes = synth.flagged_1kg("EUR") if study_pop == "EUR" else np.array([n for _, n, _ in synth.POPS_1KG], np.int32)
    assert int(sizes.sum()) == (503 if study_pop == "EUR" else 2504)
    d = np.load(SITES)
    bp_m = np.unique(d["bp"].astype(np.int64))
    bp, type_, windows = synth.chr22_windows(bp_m, unmeasured_per_mb=900.0)
    win = windows[24]                                   # a mid-size window (n_t = 429)
    meas, unme = win["measured"], win["unmeasured"]
    sites = np.concatenate([meas, unme])
    g = synth.make_genotypes(len(sites), sizes, seed=41)
    rng = np.random.default_rng(41)
    z_t = rng.standard_normal(len(meas)) * 1.34
    panel = gb.Panel(gpu_ctx, sizes, len(g))
    panel.append_host((g.astype(np.int16) + 48).astype(np.uint8), is_ascii=True)   # chars, as ReadGenotype delivers them
    rt, ru = np.arange(len(meas)), np.arange(len(meas), len(g))
    z, info, rc = panel.window_dist(rt, ru, z_t)
    assert rc == gb.GB_OK
    t = np.concatenate([np.ones(len(meas), np.int32), np.zeros(len(unme), np.int32)])
    r = oracle.run_window(t, np.arange(len(g), dtype=np.int64), np.concatenate([z_t, np.zeros(len(unme))]), g, sizes,
                          None, 0, 10**12)
    assert r["rc"] == 0
    assert np.abs(z - r["z"][ru]).max() <= TOL and np.abs(info - r["info"][ru]).max() <= TOL
    assert np.abs(z - r["z"][ru]).max() <= 1e-9


def test_cfg2_full_size_33kg_window(gpu_ctx, oracle):
    """configs[1] at full size: the largest chr22 window (n_t = 1,213 measured, ~3,700 unmeasured) on the
    33KG-shaped panel.  Correlations: 300 sampled entries vs CalWgtCov; imputation: the normal equations
    B11 x = z_t solved in numpy from the GPU's own B11 / B21."""
    _, sizes, w = synth.flagged_33kg_pgc2()
    d = np.load(SITES)
    bp_m, first = np.unique(d["bp"].astype(np.int64), return_index=True)
    bp, type_, windows = synth.chr22_windows(bp_m)
    win = max(windows, key=lambda x: len(x["measured"]))
    n_t, n_u = len(win["measured"]), len(win["unmeasured"])
    assert n_t == 1213 and n_u > 3000
    g = synth.make_genotypes(n_t + n_u, sizes, seed=42)
    rng = np.random.default_rng(42)
    z_t = rng.standard_normal(n_t) * 1.34
    panel = gb.Panel(gpu_ctx, sizes, len(g))
    panel.append_host(g, is_ascii=False)
    rt, ru = np.arange(n_t), np.arange(n_t, n_t + n_u)
    B11, B21 = panel.window_cor(rt, ru, w)
    for _ in range(150):
        i, j = rng.integers(0, n_t, 2)
        if i != j:
            assert abs(B11[i, j] - mix_cor(oracle, g[i], g[j], sizes, w)) <= 1e-11
        u, t = rng.integers(0, n_u), rng.integers(0, n_t)
        assert abs(B21[u, t] - mix_cor(oracle, g[n_t + u], g[t], sizes, w)) <= 1e-11
    np.testing.assert_array_equal(np.diag(B11), np.full(n_t, 1.1))
    np.testing.assert_array_equal(B11, B11.T)
    z, info, rc = panel.window_distmix(rt, ru, z_t, w)
    assert rc == gb.GB_OK
    T = np.linalg.solve(B11, B21.T).T                      # b21 B11^-1   (dist.cpp:193)
    info_ref = np.abs((T * B21).sum(1))
    z_ref = (T @ z_t) / np.sqrt(info_ref)
    assert np.abs(info - info_ref).max() <= 1e-9 and np.abs(z - z_ref).max() <= 1e-8
    assert (info > 0).all() and (info <= 1.0 + 1e-12).all()


def test_cfg2_mean_size_window_against_the_oracle_end_to_end(gpu_ctx, oracle):
    """configs[1] once at the MEAN window size of chr22 (n_t = 742 measured) through the whole oracle -- the restated
    run_distmix with its eigendecomposition and full-pivot LU inverse -- at the full 32,147 individuals; 160 of the
    window's unmeasured SNPs keep the oracle's O(n_t n_u N) loop to ~20 s."""
    _, sizes, w = synth.flagged_33kg_pgc2()
    d = np.load(SITES)
    bp_m = np.unique(d["bp"].astype(np.int64))
    _, _, windows = synth.chr22_windows(bp_m)
    nts = np.array([len(x["measured"]) for x in windows])
    win = windows[int(np.argmin(np.abs(nts - 742)))]
    n_t, n_u = len(win["measured"]), 160
    assert 700 <= n_t <= 790
    g = synth.make_genotypes(n_t + n_u, sizes, seed=43)
    rng = np.random.default_rng(43)
    z_t = rng.standard_normal(n_t) * 1.34
    panel = gb.Panel(gpu_ctx, sizes, len(g))
    panel.append_host(g, is_ascii=False)
    z, info, rc = panel.window_distmix(np.arange(n_t), np.arange(n_t, n_t + n_u), z_t, w)
    assert rc == gb.GB_OK
    t = np.concatenate([np.ones(n_t, np.int32), np.zeros(n_u, np.int32)])
    r = oracle.run_window(t, np.arange(n_t + n_u, dtype=np.int64), np.concatenate([z_t, np.zeros(n_u)]), g, sizes, w, 0, 10 ** 12)
    assert r["rc"] == 0 and r["n_t"] == n_t
    assert np.abs(z - r["z"][n_t:]).max() <= TOL and np.abs(info - r["info"][n_t:]).max() <= TOL
    assert np.abs(z - r["z"][n_t:]).max() <= 1e-8 and np.abs(info - r["info"][n_t:]).max() <= 1e-9


def test_cfg3_compute_ld_5000_snp_block(gpu_ctx, oracle):
    """configs[2]: computeLD() on a dense 5,000-SNP block, ancestry-mixed, 33KG-shaped panel."""
    _, sizes, w = synth.flagged_33kg_pgc2()
    n = 5000
    g = np.concatenate([synth.make_genotypes(1000, sizes, seed=50 + k) for k in range(5)])
    panel = gb.Panel(gpu_ctx, sizes, n)
    panel.append_host(g, is_ascii=False)
    ld, rc = panel.window_ld(np.arange(n), w)
    assert rc == gb.GB_OK and ld.shape == (n, n)
    np.testing.assert_array_equal(np.diag(ld), np.ones(n))          # computeLD.cpp:107: diagonal exactly 1.0
    np.testing.assert_array_equal(ld, ld.T)
    assert np.isfinite(ld).all() and np.abs(ld).max() <= 1.0 + 1e-12
    rng = np.random.default_rng(50)
    for _ in range(300):
        i, j = rng.integers(0, n, 2)
        if i != j:
            assert abs(ld[i, j] - mix_cor(oracle, g[i], g[j], sizes, w)) <= 1e-11
    # a principal sub-block of a correlation matrix is positive semi-definite
    sub = ld[1000:1400, 1000:1400]
    assert np.linalg.eigvalsh(sub).min() > -1e-9
