"""Extract the measured-SNP positions and Z-scores of the reference's bundled
data/PGC2_Chr22_ilmn1M_Z.txt (BASELINE.json configs 1-2) into a compact fixture, because
/root/reference does not exist on the GPU box.  Run in the authoring container only."""
import os

import numpy as np

SRC = "/root/reference/data/PGC2_Chr22_ilmn1M_Z.txt"
bp, z = [], []
with open(SRC) as f:
    next(f)
    for line in f:
        t = line.split()
        bp.append(int(t[2]))
        z.append(float(t[5]))
bp, z = np.array(bp, np.int64), np.array(z, np.float64)
order = np.argsort(bp, kind="stable")
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pgc2_chr22_sites.npz")
np.savez_compressed(out, bp=bp[order].astype(np.int32), z=z[order])
print(len(bp), "sites", bp.min(), bp.max(), "sd(z)=%.3f" % z.std(), os.path.getsize(out), "bytes")
