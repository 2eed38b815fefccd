"""Per-gene LD blocks of jepeg()/jepegmix() (BASELINE config 5; reference gene.cpp:300-316, 569-587): many small
correlation matrices in one batch, diagonal forced to 1 + lambda, against the oracle's computeLD / CalCor restatements."""
import numpy as np
import pytest

import gauss_b200 as gb
from gauss_b200 import synth
from helpers import small_case


def test_port_corg_equals_the_references_own_block(oracle, ref_oracle):
    """What the GPU test below compares against (computeLD restatement with the diagonal forced to 1 + lambda, and the
    pooled CalCor) is bit for bit the CorG block Gene::CalJepegmixPval / CalJepegPval build (gene.cpp:569-586, 305-315),
    compiled from the reference by oracle/build_ref.sh."""
    c = small_case(seed=72, n_snps=40)
    for n in (1, 2, 9):
        want = ref_oracle.gene_corg(c["g"][:n], c["pop_sizes"], c["w"], lam=0.1)
        got = oracle.compute_ld(c["g"][:n], c["pop_sizes"], c["w"])
        np.fill_diagonal(got, 1.0 + 0.1)
        np.testing.assert_array_equal(got, want)
        wantp = ref_oracle.gene_corg(c["g"][:n], c["pop_sizes"], None, lam=0.1)
        for i in range(n):
            assert wantp[i, i] == 1.0 + 0.1
            for j in range(i + 1, n):
                assert wantp[i, j] == wantp[j, i] == oracle.cal_cor(c["g"][i], c["g"][j], c["pop_sizes"])


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["e2m1", "int8"])
def test_genes_ld_matches_oracle(gpu_ctx, oracle, fmt):
    c = small_case(seed=71, n_snps=600, pop_sizes=(61, 103, 40, 25, 2, 330, 97))
    g = c["g"].astype(np.int8)
    panel = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), fmt)
    panel.append_host(g, is_ascii=False)
    rng = np.random.default_rng(71)
    sizes = np.concatenate([rng.integers(1, 11, 60), [1, 2, 130, 37]])        # JEPEG genes hold 1-10 SNPs; one > 128
    rows, g_off = [], [0]
    for n in sizes:
        rows.append(np.sort(rng.choice(len(g), int(n), replace=False)))        # a gene's SNPs are not contiguous
        g_off.append(g_off[-1] + int(n))
    rows = np.concatenate(rows)
    mix = panel.genes_ld(g_off, rows, c["w"], diag=1.1)
    pooled = panel.genes_ld(g_off, rows, None, diag=1.1)
    assert len(mix) == len(sizes)
    for k, n in enumerate(sizes):
        r = rows[g_off[k]:g_off[k + 1]]
        want = oracle.compute_ld(g[r], c["pop_sizes"], c["w"])
        np.fill_diagonal(want, 1.1)
        assert mix[k].shape == (n, n)
        assert np.abs(mix[k] - want).max() <= 1e-12
        np.testing.assert_array_equal(mix[k], mix[k].T)
        wantp = np.full((n, n), 1.1)
        for i in range(n):
            for j in range(i + 1, n):
                wantp[i, j] = wantp[j, i] = oracle.cal_cor(g[r[i]], g[r[j]], c["pop_sizes"])
        assert np.abs(pooled[k] - wantp).max() <= 1e-12
    # computeLD's own convention through the same entry: diagonal exactly 1.0
    one = panel.genes_ld([0, 5], rows[:5], c["w"], diag=1.0)[0]
    cm, _ = panel.window_ld(rows[:5], c["w"], allow=(gb.api.GB_ERR_TOO_FEW_MEASURED,))
    assert (np.diag(one) == 1.0).all()
    assert panel.genes_ld([0], np.zeros(0, np.int64), c["w"]) == []
