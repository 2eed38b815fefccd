"""Shared test inputs (seeded)."""
import numpy as np

from gauss_b200 import synth


def small_case(seed=0, n_snps=300, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3,
               core=(40, 260)):
    pop_sizes = np.array(pop_sizes, np.int32)
    g = synth.make_genotypes(n_snps, pop_sizes, seed=seed)
    core = (min(core[0], n_snps // 8), min(core[1], n_snps - n_snps // 8))
    type_, bp, start_bp, end_bp = synth.make_window_layout(n_snps, measured_frac, core[0], core[1], seed)
    rng = np.random.default_rng(seed + 1)
    z = rng.standard_normal(n_snps) * 1.34
    w = rng.dirichlet(np.ones(len(pop_sizes))) * 1.061
    return dict(g=g, pop_sizes=pop_sizes, type=type_, bp=bp, start_bp=start_bp, end_bp=end_bp, z=z, w=w)


def split_rows(case):
    t, bp = case["type"], case["bp"]
    meas = np.where(t == 1)[0]
    unme = np.where((t == 0) & (bp >= case["start_bp"]) & (bp <= case["end_bp"]))[0]
    return meas, unme
