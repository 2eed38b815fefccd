"""prep_zmix5's pair loop (reference zmix.cpp:151-170, util.cpp:153-169).

CPU: the C restatement equals the reference's own CalCor(std::string&, std::string&) compiled by oracle/build_ref.sh
bit for bit.  GPU: gb_zmix_pair_cor equals the restatement bit for bit (integer counts are exact and the fp64
operation order is the reference's), NaN pattern included."""
import numpy as np
import pytest

import gauss_b200 as gb
from gauss_b200 import synth
from helpers import small_case


def test_port_matches_compiled_reference(oracle, ref_oracle):
    c = small_case(seed=3, n_snps=60)
    a = oracle.zmix_pairs(c["g"], c["pop_sizes"], c["z"])
    b = ref_oracle.zmix_pairs(c["g"], c["pop_sizes"], c["z"])
    assert a.shape == (60 * 59 // 2, 1 + len(c["pop_sizes"]))
    np.testing.assert_array_equal(a, b)
    # column 0 and one entry by hand
    np.testing.assert_array_equal(a[0, 0], c["z"][0] * c["z"][1])
    x, y = c["g"][0, :61].astype(float), c["g"][1, :61].astype(float)
    assert abs(a[0, 1] - np.corrcoef(x, y)[0, 1]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["e2m1", "int8"])
def test_zmix_pairs_bit_exact(gpu_ctx, oracle, fmt):
    c = small_case(seed=61, n_snps=300, pop_sizes=(61, 103, 40, 25, 2, 330, 97))
    g = c["g"].astype(np.int8)
    g[7] = 0                     # monomorphic everywhere: NaN rows like the reference's 0/0
    panel = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), fmt)
    panel.append_host(g, is_ascii=False)
    rows = np.concatenate([np.arange(0, 140), np.arange(200, 270)])   # 210 SNPs, not contiguous
    got = panel.zmix_pair_cor(rows, c["z"][rows])
    want = oracle.zmix_pairs(g[rows], c["pop_sizes"], c["z"][rows])
    assert got.shape == want.shape == (210 * 209 // 2, 8)
    np.testing.assert_array_equal(got, want)
    assert np.isnan(got[:, 1:]).any() and np.isfinite(got[:, 0]).all()
    two = panel.zmix_pair_cor(rows[:2], c["z"][rows[:2]])
    np.testing.assert_array_equal(two, want[:1])
    with pytest.raises(gb.GaussB200Error):
        panel.zmix_pair_cor(rows[:1], c["z"][rows[:1]])


@pytest.mark.gpu
def test_zmix_pairs_33kg_shape(gpu_ctx, oracle):
    names, sizes, _ = synth.flagged_33kg_pgc2()
    g = synth.make_genotypes(48, sizes, seed=62)
    panel = gb.Panel(gpu_ctx, sizes, len(g))
    panel.append_host(g, is_ascii=False)
    z = np.linspace(-3, 3, 48)
    got = panel.zmix_pair_cor(np.arange(48), z)
    np.testing.assert_array_equal(got, oracle.zmix_pairs(g, sizes, z))
