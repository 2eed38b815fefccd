"""Experiment (GPU box): how much of the Gram kernel's time is the per-population accumulator hand-off?
Times stage 10 (the tensor-core kernel alone) on the chr22-shaped batch for population lists with the same individuals
cut differently: the 21 populations of 33KG/PGC2, only its 10 large ones, the 11 small ones merged pairwise, one pool."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench  # noqa: E402
import gauss_b200 as gb  # noqa: E402
from gauss_b200 import api, synth  # noqa: E402

_, sizes, w = synth.flagged_33kg_pgc2()
sizes = np.asarray(sizes)
w = np.asarray(w, float)
order = np.argsort(sizes)[::-1]
big, small = order[:10], order[10:]
cases = {
    "21 populations": (sizes, w),
    "10 large only": (sizes[big], w[big] / w[big].sum()),
    "x10 large + 6 merged small": (np.concatenate([sizes[big], [sizes[small[i:i + 2]].sum() for i in range(0, 11, 2)]]),
                                  None),
    "10 large + 1 merged small": (np.concatenate([sizes[big], [sizes[small].sum()]]), None),
    "4 equal": (np.full(4, int(sizes.sum()) // 4), None),
}
ctx = gb.Context(0)
stream = torch.cuda.Stream("cuda:0")
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
for name, (sz, ww) in cases.items():
    if name not in ("21 populations", "4 equal"):
        continue
    sz = np.asarray(sz, np.int32)
    ww = np.full(len(sz), 1.0 / len(sz)) if ww is None else ww
    L = bench.chr22_batch_inputs(sz)
    n_all = L["n_all"]
    row5 = api.pack5_row_bytes(sz)
    d_rows5 = torch.empty((n_all, row5), dtype=torch.uint8, device="cuda:0")
    api.synth_pack5_rows_device(ctx, 7, 21, sz, n_all, d_rows5.data_ptr(), row5, sites=L["sites"])
    panel = gb.Panel(ctx, sz, n_all, "e2m1")
    panel.append_pack5_device_ptr(d_rows5.data_ptr(), n_all, row5)
    batch = gb.Batch(panel, L["t_off"], L["rows_t"], L["u_off"], L["rows_u"], L["z_t"], ww)
    work = batch.work()
    batch.run_stage(0)
    ms = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        batch.run_stage(10)
        e1.record(stream)
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = min(ms[2:])
    ft = {}
    for mode in ("1", "2"):
        os.environ["GB_GRAM_FEEDTEST"] = mode
        ms2 = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            batch.run_stage(10)
            e1.record(stream)
            torch.cuda.synchronize()
            ms2.append(e0.elapsed_time(e1))
        ft[mode] = min(ms2[1:])
    os.environ.pop("GB_GRAM_FEEDTEST")
    print(f"    feed test: every CTA reads the same 256 rows {ft['1']:.3f} ms; every CTA its own fixed 256 rows {ft['2']:.3f} ms")
    print(f"{name:28s} K {int(sz.sum()):6d} segs {len(sz):2d}  gram {t:.3f} ms  {work['gram_ops'] / t / 1e9:8.0f} TOP/s  "
          f"ms per 32147 K: {t * 32147 / sz.sum():.3f}", flush=True)
    del batch, panel, d_rows5
