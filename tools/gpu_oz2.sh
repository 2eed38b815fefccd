#!/bin/bash
# after the padding fix: poison diagnostics, the lagging pipe flow, the whole GPU suite (default solver), chr22 stage times
mkdir -p gpurun_out
timeout 200 python tools/dbg_poison.py 2>&1 | tail -8
timeout 100 python tools/dbg_pipe.py test int8 2>&1 | grep -v "z 0.000e+00 info 0.000e+00" | tail -5
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
GB_SOLVE=fp64 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_qcat.py -m gpu -q -x 2>&1 | tail -4
