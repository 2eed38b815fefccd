import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle("port")


@pytest.fixture(scope="session")
def ref_oracle():
    from oracle.oracle_py import Oracle
    if not Oracle.available("reference"):
        pytest.skip("oracle/_ref/libgauss_ref.so not built (needs /root/reference at build time)")
    return Oracle("reference")


@pytest.fixture(scope="session")
def gpu_ctx():
    import gauss_b200 as gb
    ctx = gb.Context(0)  # raises loudly when the CUDA library or device is missing -- no fallback
    yield ctx
    ctx.close()
