"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gauss_b200.h declares, and refuses to compute without a GPU (no fallback)."""
import ctypes
import os
import re
import subprocess

import pytest

import gauss_b200 as gb
from gauss_b200 import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_loads():
    path = build.build_library()
    assert os.path.exists(path)
    lib = gb.load_library()
    assert lib.gb_version() == 100


def test_exports_every_declared_symbol():
    lib = gb.load_library()
    declared = gb.exported_symbols()
    assert len(declared) >= 25
    out = subprocess.check_output(["nm", "-D", "--defined-only", gb.library_path()], text=True)
    exported = set(re.findall(r" T (gb_\w+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    for s in declared:
        assert hasattr(lib, s)


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "gauss_b200.h"\nint main(void){ gb_params p; gb_params_default(&p); return p.check_pd==1?0:1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), gb.library_path(), "-Wl,-rpath," + os.path.dirname(gb.library_path())])
    assert subprocess.call([str(exe)]) == 0


def test_params_defaults_match_reference_arguments():
    p = gb.Params.default()  # gauss.cpp:18-35
    assert p.lambda_ == 0.1 and p.min_abs_eig == 1e-5
    assert p.min_num_measured_snp == 10 and p.min_num_unmeasured_snp == 10


def test_sass_is_blackwell_native():
    """The Gram kernel must be tcgen05 + TMA + TMEM, not a legacy mma.sync path."""
    sass = subprocess.check_output(["cuobjdump", "-sass", gb.library_path()], text=True)
    assert "UTCIMMA" in sass      # tcgen05.mma kind::i8
    assert "UTMALDG" in sass      # TMA tensor loads
    assert "LDTM" in sass         # tcgen05.ld (TMEM -> registers)
    assert "HMMA" not in sass and "IMMA.16" not in sass


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(gb.GaussB200Error) as e:
        gb.Context(0)
    assert e.value.status in (2, 3)


def test_product_never_touches_oracle():
    """gauss_b200/ must not import, link or call anything under oracle/."""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "gauss_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"oracle|go_cal_|go_run_|libgauss_ref", txt):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
