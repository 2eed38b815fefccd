"""Mint golden vectors from the REFERENCE'S OWN code (oracle/_ref/libgauss_ref.so, built by
oracle/build_ref.sh from /root/reference/src).  Run in the authoring container only:

    bash oracle/build_ref.sh && python tests/golden/make_golden.py

The fixtures are small .npz files committed beside this script; tests compare the C restatement
(oracle/gauss_oracle.c) and the CUDA path against them.  /root/reference is never read at test time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import small_case, split_rows  # noqa: E402
from oracle.oracle_py import Oracle  # noqa: E402

CASES = {
    "mix7": dict(seed=0, n_snps=300, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(40, 260)),
    "mix3_ragged": dict(seed=1, n_snps=420, pop_sizes=(33, 129, 500), measured_frac=0.36, core=(10, 400)),
    "single_pop": dict(seed=2, n_snps=200, pop_sizes=(503,), measured_frac=0.25, core=(20, 180)),
}


def main():
    ref = Oracle("reference")
    here = os.path.dirname(os.path.abspath(__file__))
    for name, kw in CASES.items():
        c = small_case(**kw)
        meas, unme = split_rows(c)
        mix = ref.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], c["w"], c["start_bp"], c["end_bp"])
        dist = ref.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], None, c["start_bp"], c["end_bp"])
        ld = ref.compute_ld(c["g"][meas], c["pop_sizes"], c["w"])
        assert mix["rc"] == 0 and dist["rc"] == 0
        pairs = np.random.default_rng(5).integers(0, kw["n_snps"], (64, 2))
        cor = np.array([ref.cal_cor(c["g"][i], c["g"][j], c["pop_sizes"]) for i, j in pairs])
        cov = np.array([ref.cal_wgt_cov(c["g"][i], c["g"][j], c["pop_sizes"], c["w"]) for i, j in pairs])
        np.savez_compressed(
            os.path.join(here, f"{name}.npz"), g=c["g"], pop_sizes=c["pop_sizes"], type=c["type"], bp=c["bp"],
            z=c["z"], w=c["w"], start_bp=c["start_bp"], end_bp=c["end_bp"], meas=meas, unme=unme,
            mix_z=mix["z"], mix_info=mix["info"], dist_z=dist["z"], dist_info=dist["info"], ld=ld,
            pairs=pairs, cal_cor=cor, cal_wgt_cov=cov)
        print(name, "n_t", len(meas), "n_u", len(unme), "bytes", os.path.getsize(os.path.join(here, f"{name}.npz")))


if __name__ == "__main__":
    main()
