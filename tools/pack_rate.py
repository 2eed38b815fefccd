"""K0 rate on the GPU box: pack_rows_kernel (int8 / ASCII rows -> E2M1 or int8 operand rows + per-population sums)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import gauss_b200 as gb  # noqa: E402
from gauss_b200 import synth  # noqa: E402

_, sizes, _ = synth.flagged_33kg_pgc2()
sizes = np.asarray(sizes, np.int32)
N = int(sizes.sum())
n = 40000
ctx = gb.Context(0)
stream = torch.cuda.Stream("cuda:0")
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
g = torch.randint(0, 3, (n, N), dtype=torch.int8, device="cuda:0")
for fmt in ("e2m1", "int8"):
    for ascii_ in (False, True):
        src = (g + 48).to(torch.uint8) if ascii_ else g
        panel = gb.Panel(ctx, sizes, n, fmt)
        ms = []
        for _ in range(4):
            panel.clear()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            panel.append_device_ptr(src.data_ptr(), n, N, is_ascii=ascii_)
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        k_stride = int(((sizes + 127) // 128 * 128).sum()) // (2 if fmt == "e2m1" else 1)
        by = n * (N + k_stride + 8 * len(sizes))
        print(f"{fmt} ascii={ascii_}: {min(ms[1:]):.3f} ms  {by / (min(ms[1:]) / 1e3) / 1e9:.0f} GB/s (read + write)")
        panel.close()
