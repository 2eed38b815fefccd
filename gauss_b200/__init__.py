"""gauss_b200 -- B200-native (sm_100a) window hot path of GAUSS: dist() / distmix() / computeLD().

The product is the C-ABI shared library ``gauss_b200/lib/libgauss_b200.so`` (see
``include/gauss_b200.h``); this package is only its ctypes binding, used by tests and bench.py.
There is no CPU fallback: importing works anywhere, every compute call needs a B200.
"""
from .api import (  # noqa: F401
    GB_OK, GaussB200Error, Context, Panel, Batch, Pipe, Genome, Params, load_library, library_path, exported_symbols,
)

__all__ = ["GB_OK", "GaussB200Error", "Context", "Panel", "Batch", "Pipe", "Genome", "Params", "load_library",
           "library_path", "exported_symbols"]
