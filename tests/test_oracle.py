"""CPU tests that pin the oracle: (1) the C restatement against the golden vectors minted from the
reference's own compiled code, (2) directly against oracle/_ref when it is present, (3) against
independent numpy formulas."""
import glob
import os

import numpy as np
import pytest

from helpers import small_case, split_rows

GOLDEN = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
                if not os.path.basename(p).startswith("pgc2_"))


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_port_matches_golden_bitwise(oracle, path):
    d = np.load(path)
    g, m, w = d["g"], d["pop_sizes"], d["w"]
    for (i, j), cor, cov in zip(d["pairs"], d["cal_cor"], d["cal_wgt_cov"]):
        a = oracle.cal_cor(g[i], g[j], m)
        assert a == cor or (np.isnan(a) and np.isnan(cor))
        assert oracle.cal_wgt_cov(g[i], g[j], m, w) == cov
    mix = oracle.run_window(d["type"], d["bp"], d["z"], g, m, w, int(d["start_bp"]), int(d["end_bp"]))
    dist = oracle.run_window(d["type"], d["bp"], d["z"], g, m, None, int(d["start_bp"]), int(d["end_bp"]))
    assert mix["rc"] == 0 and dist["rc"] == 0
    np.testing.assert_array_equal(mix["z"], d["mix_z"])
    np.testing.assert_array_equal(mix["info"], d["mix_info"])
    np.testing.assert_array_equal(dist["z"], d["dist_z"])
    np.testing.assert_array_equal(dist["info"], d["dist_info"])
    np.testing.assert_array_equal(oracle.compute_ld(g[d["meas"]], m, w), d["ld"])


def test_port_matches_compiled_reference(oracle, ref_oracle):
    c = small_case(seed=11, n_snps=260, pop_sizes=(50, 7, 211, 96))
    for w in (c["w"], None):
        a = oracle.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
        b = ref_oracle.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
        assert a["rc"] == b["rc"] == 0 and a["n_t"] == b["n_t"] and a["n_u"] == b["n_u"]
        np.testing.assert_array_equal(a["z"], b["z"])
        np.testing.assert_array_equal(a["info"], b["info"])


def test_too_few_snps_is_an_error(oracle):
    c = small_case(seed=3, n_snps=30, core=(2, 28))
    r = oracle.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], c["w"], 0, 10**12)
    n_t = int((c["type"] == 1).sum())
    assert (r["rc"] == 1) == (n_t <= 10 or (30 - n_t) <= 10)
    r = oracle.run_window(c["type"][:15], c["bp"][:15], c["z"][:15], c["g"][:15], c["pop_sizes"], None, 0, 10**12)
    assert r["rc"] == 1  # dist.cpp:146-151


def test_window_against_numpy(oracle):
    c = small_case(seed=5)
    meas, unme = split_rows(c)
    g, m, w = c["g"].astype(np.float64), c["pop_sizes"], c["w"]
    r = oracle.run_window(c["type"], c["bp"], c["z"], c["g"], m, w, c["start_bp"], c["end_bp"], dump=True)
    offs = np.concatenate([[0], np.cumsum(m)])
    # CalWgtCov in closed form
    cov = np.zeros((len(g), len(g)))
    mean_w = np.zeros(len(g))
    for p, mp in enumerate(m):
        x = g[:, offs[p]:offs[p + 1]]
        s = x.sum(1)
        cov += w[p] * (mp / (mp - 1)) * (mp * x @ x.T - np.outer(s, s)) + w[p] * np.outer(s / mp, s / mp)
        mean_w += w[p] * s / mp
    cov -= np.outer(mean_w, mean_w)
    sd = np.sqrt(np.diag(cov))
    cor = cov / np.outer(sd, sd)
    B11 = cor[np.ix_(meas, meas)].copy()
    np.fill_diagonal(B11, 1.1)
    B21 = cor[np.ix_(unme, meas)]
    np.testing.assert_allclose(r["B11"], B11, atol=1e-12)
    np.testing.assert_allclose(r["B21"], B21, atol=1e-12)
    T = np.linalg.solve(B11, B21.T).T
    info = np.abs((T * B21).sum(1))
    np.testing.assert_allclose(r["info"][unme], info, atol=1e-10)
    np.testing.assert_allclose(r["z"][unme], (T @ c["z"][meas]) / np.sqrt(info), atol=1e-10)


def test_eigen_and_inverse_restatements(oracle):
    rng = np.random.default_rng(0)
    n = 150
    X = rng.standard_normal((n, 40))
    A = X @ X.T / 40 + 0.1 * np.eye(n)   # rank-deficient + ridge, like B11
    ev = np.zeros(n)
    Q = np.zeros((n, n))
    assert oracle.lib.go_sym_eig(np.ascontiguousarray(A), n, ev, Q.ctypes.data) == 0
    np.testing.assert_allclose(ev, np.linalg.eigvalsh(A), atol=1e-12)
    inv = np.zeros((n, n))
    oracle.lib.go_inv_full_piv_lu(inv, np.ascontiguousarray(A), n)
    np.testing.assert_allclose(inv @ A, np.eye(n), atol=1e-10)
    B = A - 0.3 * np.eye(n)              # indefinite -> MakePosDef clips (util.cpp:309-316)
    B2 = B.copy()
    assert oracle.lib.go_make_pos_def(B2, n, 1e-5) == 1
    lam, V = np.linalg.eigh(B)
    np.testing.assert_allclose(B2, (V * np.maximum(lam, 1e-5)) @ V.T, atol=1e-11)
    A2 = A.copy()
    assert oracle.lib.go_make_pos_def(A2, n, 1e-5) == 0 and (A2 == A).all()


def test_gram_counts_bruteforce(oracle):
    c = small_case(seed=7, n_snps=40)
    g, m = c["g"], c["pop_sizes"]
    sxy, sx, sxx = oracle.gram_counts(g[:20], g[20:], m)
    offs = np.concatenate([[0], np.cumsum(m)])
    for p in range(len(m)):
        a = g[:20, offs[p]:offs[p + 1]].astype(np.int64)
        b = g[20:, offs[p]:offs[p + 1]].astype(np.int64)
        np.testing.assert_array_equal(sxy[p], a @ b.T)
        np.testing.assert_array_equal(sx[p], a.sum(1))
        np.testing.assert_array_equal(sxx[p], (a * a).sum(1))
