"""The Rcpp-side patch of INTEGRATION.md section 2, compiled and run: the patched
`run_distmix` / `run_dist` -- the reference's own signature over the reference's own Snp / Arguments classes, calling
libgauss_b200.so -- against the UNPATCHED functions of oracle/_ref on identical std::vector<Snp*>."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import small_case

HERE = os.path.dirname(os.path.abspath(__file__))
PATCHED = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libgauss_patched.so")


def test_patched_library_exists_and_binds_the_product():
    if not os.path.exists(PATCHED):
        pytest.skip("oracle/_ref/libgauss_patched.so not built (needs /root/reference at build time)")
    import subprocess
    syms = subprocess.check_output(["nm", "-D", PATCHED], text=True)
    assert " U gb_run_window_strings" in syms and " T go_run_window_patched" in syms
    # the patch is the reference's own seam: run_distmix / run_dist with the reference's mangled signatures
    assert "_Z11run_distmixRSt6vectorIP3SnpSaIS1_EER9Arguments" in syms and "_Z8run_distRSt6vectorIP3SnpSaIS1_EER9Arguments" in syms


def run_patched(case, w, start_bp, end_bp, min_m=10, min_u=10):
    lib = C.CDLL(PATCHED)
    g = np.ascontiguousarray((case["g"] + 48).astype(np.uint8))
    z = np.array(case["z"], np.float64, copy=True)
    info = np.ones(len(z))
    m = np.ascontiguousarray(case["pop_sizes"], np.int32)
    err = C.create_string_buffer(512)
    lib.go_run_window_patched.restype = C.c_int
    wv = None if w is None else np.ascontiguousarray(w, np.float64)
    rc = lib.go_run_window_patched(np.ascontiguousarray(case["type"], np.int32).ctypes.data_as(C.c_void_p),
                                   np.ascontiguousarray(case["bp"], np.int64).ctypes.data_as(C.c_void_p),
                                   z.ctypes.data_as(C.c_void_p), info.ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p),
                                   C.c_int64(len(z)), m.ctypes.data_as(C.c_void_p), C.c_int(len(m)),
                                   None if wv is None else wv.ctypes.data_as(C.c_void_p), C.c_longlong(start_bp),
                                   C.c_longlong(end_bp), C.c_double(0.1), C.c_double(1e-5), C.c_int(min_m), C.c_int(min_u),
                                   err, C.c_int(len(err)))
    return rc, z, info, err.value.decode()


@pytest.mark.gpu
@pytest.mark.parametrize("mix", [True, False])
def test_patched_run_window_equals_the_unpatched_reference(gpu_ctx, ref_oracle, mix):
    if not os.path.exists(PATCHED):
        pytest.skip("patched library not built")
    c = small_case(seed=41, n_snps=520, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.35, core=(40, 480))
    w = c["w"] if mix else None
    rc, z, info, err = run_patched(c, w, c["start_bp"], c["end_bp"])
    assert rc == 0, err
    r = ref_oracle.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"])
    assert r["rc"] == 0
    imputed = (c["type"] == 0) & (c["bp"] >= c["start_bp"]) & (c["bp"] <= c["end_bp"])
    assert imputed.sum() > 100
    assert np.abs(z - r["z"]).max() <= 1e-6 and np.abs(info[imputed] - r["info"][imputed]).max() <= 1e-6   # GetZ() / GetInfo()
    assert np.abs(z - r["z"]).max() <= 1e-9
    # SNPs the reference leaves untouched stay untouched (measured, type 2, outside the core window)
    np.testing.assert_array_equal(z[~imputed], np.asarray(c["z"])[~imputed])
    # <= 10 SNPs: the same Rcpp::stop message as dist.cpp:150 / distmix.cpp:159
    rc, _, _, err = run_patched(c, w, c["start_bp"], c["end_bp"], min_m=10 ** 6)
    assert rc == 1 and err == f"Not enough number of SNPs loaded - {'DISTMIX' if mix else 'DIST'} not performed"
    r2 = ref_oracle.run_window(c["type"], c["bp"], c["z"], c["g"], c["pop_sizes"], w, c["start_bp"], c["end_bp"],
                               min_measured=10 ** 6)
    assert r2["rc"] != 0          # the unpatched function throws there too
