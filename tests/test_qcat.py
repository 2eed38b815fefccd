"""qcat() / qcatmix() window (reference qcat.cpp:133-238, qcatmix.cpp:140-269).

The reference ships no vectors for it.  The C restatement (oracle/gauss_oracle.c go_run_qcat) is pinned twice on the CPU:
bit for bit against the reference's OWN run_qcat / run_qcatmix bodies compiled by oracle/build_ref.sh (with the restated
Eigen algorithms, like run_dist), and against an independent numpy computation; the CUDA path is checked against the
restatement (GPU tests).  Tolerance: 1e-6 absolute on qcat_t / qcat_chisq (asserted
tighter where it holds); num_eig is an integer and must match exactly."""
import numpy as np
import pytest

import gauss_b200 as gb
from gauss_b200 import api, synth
from helpers import small_case

TOL = 1e-6


def split(c):
    t, bp = c["type"], c["bp"]
    meas = np.where(t == 1)[0]
    core = (bp >= c["start_bp"]) & (bp <= c["end_bp"])
    unme = np.where((t == 0) & core)[0]
    headwing = int(((t == 1) & (bp < c["start_bp"])).sum())
    n_core = int(((t == 1) & core).sum())
    return meas, unme, headwing, n_core


def test_oracle_qcat_against_numpy(oracle):
    c = small_case(seed=5, n_snps=260, measured_frac=0.35, core=(60, 200))
    g, t, bp, z = c["g"], c["type"], c["bp"], c["z"]
    meas, unme, headwing, n_core = split(c)
    r = oracle.run_qcat(t, bp, z, g, c["pop_sizes"], None, c["start_bp"], c["end_bp"])
    assert r["rc"] == 0
    C = np.corrcoef(g.astype(float))
    B11 = C[np.ix_(meas, meas)].copy()
    np.fill_diagonal(B11, 1.1)
    L = np.linalg.cholesky(B11)
    y = np.linalg.solve(L, z[meas])
    ne = int((np.linalg.eigvalsh(B11) >= 0.01).sum())
    assert ne == len(meas)                       # ridge 0.1 > cut-off 0.01: CountPC's branch is dead

    def rr(b):
        return np.corrcoef(y, np.linalg.solve(L, b))[0, 1]

    tested = np.zeros(len(t), bool)
    for u in unme:
        rv = rr(C[u, meas])
        assert abs(r["t"][u] - np.sqrt(ne - 3) * rv) < 1e-10 and abs(r["chisq"][u] - (ne - 3) * rv * rv) < 1e-10
        tested[u] = True
    for k in range(headwing, headwing + n_core):
        rv = rr(B11[k])
        assert abs(r["t"][meas[k]] - np.sqrt(ne - 3) * rv) < 1e-10 and r["m"][meas[k]] == ne
        tested[meas[k]] = True
    assert np.isnan(r["t"][~tested]).all() and np.isfinite(r["t"][tested]).all()
    # CountPC really counts: a cut-off above the ridge removes components
    assert oracle.lib.go_count_pc(np.ascontiguousarray(B11), len(meas), 0.5) < len(meas)
    few = oracle.run_qcat(t[:20], bp[:20], z[:20], g[:20], c["pop_sizes"], None, 0, 10**12)
    assert few["rc"] != 0


@pytest.mark.parametrize("mix", [False, True])
def test_port_matches_compiled_reference(oracle, ref_oracle, mix):
    c = small_case(seed=7 + mix, n_snps=280, measured_frac=0.35, core=(60, 210))
    w = c["w"] if mix else None
    a = oracle.run_qcat(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    b = ref_oracle.run_qcat(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    assert a["rc"] == b["rc"] == 0
    for k in ("m", "t", "chisq"):
        np.testing.assert_array_equal(a[k], b[k])          # same tested set (NaN elsewhere), same bits
    assert np.isfinite(a["t"]).sum() > 100
    # the two entry points differ in what they refuse: run_qcat only looks at the measured count (qcat.cpp:157),
    # run_qcatmix also at the unmeasured one (qcatmix.cpp:168-169)
    t2 = c["type"].copy()
    core = (c["bp"] >= c["start_bp"]) & (c["bp"] <= c["end_bp"])
    t2[(t2 == 0) & core] = 2                                    # no unmeasured SNP left in the prediction window
    a2 = oracle.run_qcat(t2, c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    b2 = ref_oracle.run_qcat(t2, c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    assert a2["rc"] == b2["rc"] and (a2["rc"] != 0) == mix


@pytest.mark.gpu
@pytest.mark.parametrize("mix", [False, True])
@pytest.mark.parametrize("fmt", ["e2m1", "int8"])
def test_qcat_window_matches_oracle(gpu_ctx, oracle, mix, fmt):
    c = small_case(seed=51, n_snps=700, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(150, 560))
    g, t, bp, z = c["g"].astype(np.int8), c["type"], c["bp"], c["z"]
    w = c["w"] if mix else None
    meas, unme, headwing, n_core = split(c)
    assert headwing > 0 and n_core > 20 and headwing + n_core < len(meas)
    panel = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), fmt)
    panel.append_host(g, is_ascii=False)
    out = panel.window_qcat(meas, z[meas], headwing, n_core, unme, w)
    ref = oracle.run_qcat(t, bp, z, g, c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    assert out["rc"] == 0 and ref["rc"] == 0
    core_m = meas[headwing:headwing + n_core]
    assert out["num_eig"] == len(meas) == int(ref["m"][unme[0]])
    for got, want in ((out["t_u"], ref["t"][unme]), (out["chisq_u"], ref["chisq"][unme]),
                      (out["t_m"], ref["t"][core_m]), (out["chisq_m"], ref["chisq"][core_m])):
        assert np.isfinite(got).all()
        assert np.abs(got - want).max() <= TOL
        assert np.abs(got - want).max() <= 1e-8
    # chisq == t^2 and |r| <= 1
    np.testing.assert_allclose(out["chisq_u"], out["t_u"] ** 2, rtol=1e-12)
    assert (np.abs(out["t_m"]) <= np.sqrt(len(meas) - 3) * (1 + 1e-12)).all()
    # flipping Z flips t and keeps chisq
    neg = panel.window_qcat(meas, -z[meas], headwing, n_core, unme, w)
    np.testing.assert_array_equal(neg["t_u"], -out["t_u"])
    np.testing.assert_array_equal(neg["chisq_m"], out["chisq_m"])


@pytest.mark.gpu
def test_qcat_edge_cases(gpu_ctx, oracle):
    c = small_case(seed=52, n_snps=400, measured_frac=0.3, core=(100, 300))
    g, t, bp, z = c["g"].astype(np.int8), c["type"], c["bp"], c["z"]
    meas, unme, headwing, n_core = split(c)
    panel = gb.Panel(gpu_ctx, c["pop_sizes"], len(g))
    panel.append_host(g, is_ascii=False)
    # too few measured SNPs: the reference stops (qcat.cpp:157)
    r = panel.window_qcat(meas[:10], z[meas[:10]], 0, 5, unme, None, allow=(api.GB_ERR_TOO_FEW_MEASURED,))
    assert r["rc"] == api.GB_ERR_TOO_FEW_MEASURED
    # no unmeasured SNPs at all: only the measured ones are tested (qcat has no unmeasured threshold)
    only_m = panel.window_qcat(meas, z[meas], headwing, n_core, np.zeros(0, np.int64), None)
    full = panel.window_qcat(meas, z[meas], headwing, n_core, unme, None)
    np.testing.assert_array_equal(only_m["t_m"], full["t_m"])
    # ... but qcatmix has one (qcatmix.cpp:168-169)
    r = panel.window_qcat(meas, z[meas], headwing, n_core, unme[:10], c["w"], allow=(api.GB_ERR_TOO_FEW_UNMEASURED,))
    assert r["rc"] == api.GB_ERR_TOO_FEW_UNMEASURED
    assert panel.window_qcat(meas, z[meas], headwing, n_core, unme[:10], None)["rc"] == 0
    # nothing measured in the prediction window
    none_m = panel.window_qcat(meas, z[meas], headwing, 0, unme, None)
    np.testing.assert_array_equal(none_m["t_u"], full["t_u"])
    # a cut-off above the ridge cannot be certified by a factorisation: the device eigendecomposition counts the
    # components like CountPC (util.cpp:355-388) and the tests use num_eig - 3 degrees of freedom
    r = panel.window_qcat(meas, z[meas], headwing, n_core, unme, None, eig_cutoff=0.5)
    ref = oracle.run_qcat(t, bp, z, g, c["pop_sizes"], None, c["start_bp"], c["end_bp"], eig_cutoff=0.5)
    assert r["rc"] == 0 and ref["rc"] == 0
    assert r["num_eig"] == int(ref["m"][unme[0]]) < len(meas)
    assert np.abs(r["t_u"] - ref["t"][unme]).max() <= 1e-8 and np.abs(r["chisq_u"] - ref["chisq"][unme]).max() <= 1e-8
    core_m = meas[headwing:headwing + n_core]
    assert np.abs(r["t_m"] - ref["t"][core_m]).max() <= 1e-8
    with pytest.raises(gb.GaussB200Error):
        panel.window_qcat(meas, z[meas], len(meas) - 2, 5, unme, None)     # core range runs past the measured list


@pytest.mark.gpu
def test_qcatmix_33kg_shape(gpu_ctx, oracle):
    """qcatmix on the BASELINE config-2 panel shape: 21 flagged populations, 32,147 individuals, PGC2 weights."""
    _, sizes, w = synth.flagged_33kg_pgc2()
    n = 420
    g = synth.make_genotypes(n, sizes, seed=53)
    rng = np.random.default_rng(53)
    t = (rng.random(n) < 0.35).astype(np.int32)
    bp = np.arange(n, dtype=np.int64) * 300 + 1
    z = rng.standard_normal(n) * 1.34
    c = dict(type=t, bp=bp, start_bp=int(bp[100]), end_bp=int(bp[330]))
    meas, unme, headwing, n_core = split(c)
    panel = gb.Panel(gpu_ctx, sizes, n)
    panel.append_host(g, is_ascii=False)
    out = panel.window_qcat(meas, z[meas], headwing, n_core, unme, w)
    ref = oracle.run_qcat(t, bp, z, g, sizes, w, c["start_bp"], c["end_bp"])
    core_m = meas[headwing:headwing + n_core]
    assert out["num_eig"] == len(meas) == int(ref["m"][core_m[0]])
    assert np.abs(out["t_u"] - ref["t"][unme]).max() <= 1e-8
    assert np.abs(out["t_m"] - ref["t"][core_m]).max() <= 1e-8
    assert np.abs(out["chisq_u"] - ref["chisq"][unme]).max() <= 1e-8


@pytest.mark.gpu
@pytest.mark.parametrize("mix", [False, True])
def test_host_mirror_of_run_qcat_strings(gpu_ctx, oracle, mix):
    """gb_run_qcat_strings receives exactly what run_qcat / run_qcatmix receive (bp-sorted SNPs with type, z and one
    genotype string per flagged population) and must leave what they leave on the Snp objects."""
    c = small_case(seed=55 + mix, n_snps=360, measured_frac=0.33, core=(90, 280))
    g, t, bp, z = c["g"], c["type"].copy(), c["bp"], c["z"]
    t[::17] = 2                                             # SNPs absent from the panel are ignored (type 2)
    w = c["w"] if mix else None
    offs = np.concatenate([[0], np.cumsum(c["pop_sizes"])])
    chars = (g.astype(np.int16) + 48).astype(np.uint8)
    strings = [[chars[i, offs[k]:offs[k + 1]].tobytes() for k in range(len(c["pop_sizes"]))] if t[i] != 2 else None
               for i in range(len(g))]
    strings = [s if s is not None else [b""] * len(c["pop_sizes"]) for s in strings]
    out = api.run_qcat_strings(gpu_ctx, t, bp, z, strings, c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    ref = oracle.run_qcat(t, bp, z, g, c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    assert out["rc"] == 0 and ref["rc"] == 0
    np.testing.assert_array_equal(np.isnan(out["t"]), np.isnan(ref["t"]))          # the same SNPs are tested
    tested = ~np.isnan(ref["t"])
    assert tested.sum() > 100
    np.testing.assert_array_equal(out["m"][tested], ref["m"][tested])
    assert np.abs(out["t"][tested] - ref["t"][tested]).max() <= 1e-8
    assert np.abs(out["chisq"][tested] - ref["chisq"][tested]).max() <= 1e-8
    few = api.run_qcat_strings(gpu_ctx, t[:25], bp[:25], z[:25], strings[:25], c["pop_sizes"], w, 0, 10**12)
    assert few["rc"] == api.GB_ERR_TOO_FEW_MEASURED
