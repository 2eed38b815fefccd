"""jepeg() / jepegmix() per-gene statistics (BASELINE config 5): gb_genes_jepeg against the
reference's own Gene::RunJepegmix / RunJepeg (gene.cpp compiled unmodified into oracle/_ref) on synthetic annotation."""
import ctypes as C

import numpy as np
import pytest

from gauss_b200 import synth

POPS = np.array([61, 103, 40, 25, 2, 330, 97], np.int32)


def synthetic_annotation(n_genes, rng):
    """Genes of 1-10 SNPs (the sizes docs/articles/jepeg_example.md shows), each SNP annotated in 1-3 of the six
    categories with a positive weight -- the shape of JEPEG_SNP_Annotation.v1.0.txt, which is not bundled."""
    sizes = rng.integers(1, 11, n_genes)
    sizes[:4] = [1, 2, 10, 7]
    g_off = np.concatenate([[0], np.cumsum(sizes)])
    n = int(g_off[-1])
    cw = np.full((n, 6), np.nan)
    for i in range(n):
        for c in rng.choice(6, rng.integers(1, 4), replace=False):
            cw[i, c] = rng.uniform(0.2, 1.0)
    return g_off, cw


def ref_gene(lib, g, w, z, info, cw):
    out = np.zeros(8)
    geno = np.ascontiguousarray(g + 48, np.uint8)
    lib.go_gene_jepeg.restype = None
    lib.go_gene_jepeg(geno.ctypes.data_as(C.c_void_p), C.c_int64(len(g)), POPS.ctypes.data_as(C.c_void_p), C.c_int(len(POPS)),
                      None if w is None else w.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p),
                      info.ctypes.data_as(C.c_void_p), np.ascontiguousarray(cw).ctypes.data_as(C.c_void_p),
                      C.c_double(0.1), C.c_double(1e-5), C.c_double(0.8), C.c_int(3), out.ctypes.data_as(C.c_void_p))
    return out


def test_rmath_restatements_match_scipy():
    """R::pnorm5 / R::pchisq are restated in oracle/ref_shim/Rmath.h (nmath is not installed): pin them."""
    from scipy import stats
    from oracle import oracle_py
    if not oracle_py.Oracle.available("reference"):
        pytest.skip("oracle/_ref not built")
    # exercised through a one-SNP gene: chisq = u^2, df = 1, p = pchisq(u^2, 1, upper) = 2 pnorm(-|u|)
    lib = C.CDLL(oracle_py.REF_SO)
    g = synth.make_genotypes(1, POPS, seed=3)
    for zval in (0.3, 1.96, 5.0, 9.0):
        cw = np.full((1, 6), np.nan)
        cw[0, 2] = 0.7
        r = ref_gene(lib, g, None, np.array([zval]), np.array([1.0]), cw)
        assert r[1] == 1 and r[0] == pytest.approx(zval * zval / 1.1, rel=1e-12)
        assert r[2] == pytest.approx(stats.chi2.sf(r[0], 1), rel=1e-10)
        assert r[4] == pytest.approx(2 * stats.norm.sf(abs(zval) / np.sqrt(1.1)), rel=1e-10) or r[3] == 2


@pytest.mark.gpu
@pytest.mark.parametrize("mix", [True, False])
def test_genes_jepeg_matches_the_reference_gene_code(gpu_ctx, ref_oracle, mix):
    import gauss_b200 as gb
    from oracle import oracle_py
    lib = C.CDLL(oracle_py.REF_SO)
    rng = np.random.default_rng(7)
    g_off, cw = synthetic_annotation(300, rng)
    n = int(g_off[-1])
    g = synth.make_genotypes(n, POPS, seed=17)
    g[g_off[5] + 1] = g[g_off[5]]                    # two identical SNPs in one gene: collinear categories / clipped CovX
    z = rng.standard_normal(n) * 1.5
    info = np.where(rng.random(n) < 0.8, 1.0, rng.uniform(0.3, 1.0, n))
    w = rng.dirichlet(np.ones(len(POPS))) * 1.061 if mix else None
    panel = gb.Panel(gpu_ctx, POPS, n)
    panel.append_host(g, is_ascii=False)
    out = panel.genes_jepeg(g_off, np.arange(n), z, info, cw, w)
    n_df = 0
    for gi in range(len(g_off) - 1):
        a, b = g_off[gi], g_off[gi + 1]
        r = ref_gene(lib, g[a:b], w, z[a:b], info[a:b], cw[a:b])
        o = out[gi]
        assert o[1] == r[1], (gi, o[:7], r)                                   # df
        assert o[6] == r[5]                                                   # top SNP
        if r[1] == 0:
            assert o[0] == -1 and o[2] == -1 and o[4] == -1
            continue
        n_df += 1
        assert o[0] == pytest.approx(r[0], rel=1e-8, abs=1e-9), (gi, o[:7], r)  # chisq
        assert o[2] == pytest.approx(r[2], rel=1e-7, abs=1e-300)              # jepeg_pval
        assert o[4] == r[3] and o[5] == pytest.approx(r[4], rel=1e-9)         # top category and its p-value
    assert n_df > 250
