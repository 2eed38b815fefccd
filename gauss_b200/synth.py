"""Synthetic workloads shaped like the reference's panels.  Data generation only.

No reference panel is downloadable here (docs/articles/ref_33KG.md:7 is a Drive link), so tests and
bench.py use seeded synthetic panels with the real population structure:
  * 33KG: 29 populations, 32,953 individuals (docs/articles/ref_33KG.md:15-43)
  * PGC2_SCZ_ANC_Prop weights: 21 populations, sum 1.061 (data/PGC2_SCZ_ANC_Prop.RData, decoded in
    SURVEY.md Appendix C)
  * 1KG phase 3: 26 populations, 2,504 individuals (EUR = 503)
Genotypes are dosages {0,1,2} with haplotype-copy LD so B11 is realistically ill-conditioned.
"""
from __future__ import annotations

import numpy as np

POPS_33KG = [
    ("ACB", 164, "AFR"), ("ASW", 162, "AFR"), ("BEB", 86, "SAS"), ("CCE", 3409, "ASN"), ("CCS", 2613, "ASN"),
    ("CDX", 95, "ASN"), ("CEU", 6360, "EUR"), ("CLM", 98, "AMR"), ("CNE", 2330, "ASN"), ("CSE", 2020, "ASN"),
    ("ESN", 140, "AFR"), ("FIN", 3529, "EUR"), ("GBR", 2020, "EUR"), ("GIH", 110, "SAS"), ("GWD", 113, "AFR"),
    ("IBS", 1309, "EUR"), ("ITU", 95, "SAS"), ("JPT", 107, "ASN"), ("KHV", 226, "ASN"), ("LWK", 99, "AFR"),
    ("MSL", 87, "AFR"), ("MXL", 187, "AMR"), ("ORK", 5772, "EUR"), ("PEL", 110, "AMR"), ("PJL", 121, "SAS"),
    ("PUR", 138, "AMR"), ("STU", 110, "SAS"), ("TSI", 1291, "EUR"), ("YRI", 52, "AFR"),
]
PGC2_SCZ_ANC_PROP = {
    "ACB": .006, "ASW": .036, "BEB": .005, "CCE": .008, "CCS": .004, "CDX": .018, "CEU": .165, "CLM": .025,
    "CNE": .003, "CSE": .012, "FIN": .138, "GBR": .165, "GIH": .006, "IBS": .099, "JPT": .011, "KHV": .017,
    "MXL": .030, "ORK": .166, "PJL": .016, "PUR": .045, "TSI": .086,
}
POPS_1KG = [
    ("CHB", 103, "EAS"), ("JPT", 104, "EAS"), ("CHS", 105, "EAS"), ("CDX", 93, "EAS"), ("KHV", 99, "EAS"),
    ("CEU", 99, "EUR"), ("TSI", 107, "EUR"), ("FIN", 99, "EUR"), ("GBR", 91, "EUR"), ("IBS", 107, "EUR"),
    ("YRI", 108, "AFR"), ("LWK", 99, "AFR"), ("GWD", 113, "AFR"), ("MSL", 85, "AFR"), ("ESN", 99, "AFR"),
    ("ASW", 61, "AFR"), ("ACB", 96, "AFR"), ("MXL", 64, "AMR"), ("PUR", 104, "AMR"), ("CLM", 94, "AMR"),
    ("PEL", 85, "AMR"), ("GIH", 103, "SAS"), ("PJL", 96, "SAS"), ("BEB", 86, "SAS"), ("STU", 102, "SAS"),
    ("ITU", 102, "SAS"),
]


def flagged_33kg_pgc2():
    """Flagged populations (panel order) and their weights, as init_pop_flag_wgt_vec builds them
    (gauss.cpp:1093-1117): 21 populations, 32,147 individuals."""
    sizes, wgts, names = [], [], []
    for name, n, _ in POPS_33KG:
        if name in PGC2_SCZ_ANC_PROP:
            names.append(name)
            sizes.append(n)
            wgts.append(PGC2_SCZ_ANC_PROP[name])
    return names, np.array(sizes, np.int32), np.array(wgts, np.float64)


def flagged_1kg(study_pop="EUR"):
    """init_pop_flag_vec (gauss.cpp:1019-1066): match against population or super-population."""
    sizes = [n for name, n, sup in POPS_1KG if study_pop in (name, sup)]
    return np.array(sizes, np.int32)


def make_genotypes(n_snps: int, pop_sizes, seed: int = 0, copy_lo: float = 0.7, copy_hi: float = 0.95,
                   dtype=np.int8):
    """[n_snps, sum(pop_sizes)] dosages with per-population allele frequencies and haplotype-copy LD."""
    rng = np.random.default_rng(seed)
    pop_sizes = np.asarray(pop_sizes)
    N = int(pop_sizes.sum())
    f = rng.uniform(0.01, 0.5, n_snps)
    rho = rng.uniform(copy_lo, copy_hi, n_snps)
    rho[0] = 0.0
    out = np.empty((n_snps, N), dtype)
    off = 0
    idx = np.arange(n_snps)[:, None]
    for m in pop_sizes:
        m = int(m)
        fp = np.clip(f + rng.normal(0, 0.05, n_snps), 0.005, 0.995)
        dose = np.zeros((n_snps, m), np.int16)
        for _hap in range(2):
            fresh = rng.random((n_snps, m)) < fp[:, None]
            reset = rng.random((n_snps, m)) >= rho[:, None]          # site starts a new haplotype segment
            src = np.maximum.accumulate(np.where(reset, idx, 0), axis=0)  # last reset site at or before i
            dose += np.take_along_axis(fresh, src, axis=0)
        out[:, off:off + m] = dose.astype(dtype)
        off += m
    return out


def make_genotypes_torch(n_snps: int, pop_sizes, device, seed: int = 0, copy_lo: float = 0.7,
                         copy_hi: float = 0.95, chunk: int = 4096):
    """Same model on a CUDA device (bench-sized panels); returns an int8 torch tensor [n_snps, N]."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    N = int(np.sum(pop_sizes))
    f = torch.rand(n_snps, generator=g, device=device) * 0.49 + 0.01
    rho = torch.rand(n_snps, generator=g, device=device) * (copy_hi - copy_lo) + copy_lo
    rho[0] = 0.0
    out = torch.empty((n_snps, N), dtype=torch.int8, device=device)
    idx = torch.arange(n_snps, device=device, dtype=torch.int32)[:, None]
    off = 0
    for m in pop_sizes:
        m = int(m)
        fp = torch.clamp(f + 0.05 * torch.randn(n_snps, generator=g, device=device), 0.005, 0.995)
        for c0 in range(0, m, chunk):
            c = min(chunk, m - c0)
            dose = torch.zeros((n_snps, c), dtype=torch.int8, device=device)
            for _hap in range(2):
                fresh = torch.rand((n_snps, c), generator=g, device=device) < fp[:, None]
                reset = torch.rand((n_snps, c), generator=g, device=device) >= rho[:, None]
                src = torch.cummax(torch.where(reset, idx, torch.zeros_like(idx)), dim=0).values
                dose += torch.gather(fresh, 0, src.long()).to(torch.int8)
            out[:, off + c0:off + c0 + c] = dose
        off += m
    return out


def make_window_layout(n_snps: int, measured_frac: float, core_lo: int, core_hi: int, seed: int = 0):
    """type (0/1) and bp for a bp-sorted SNP list; measured SNPs are interleaved like a real array."""
    rng = np.random.default_rng(seed + 77)
    type_ = (rng.random(n_snps) < measured_frac).astype(np.int32)
    bp = np.sort(rng.choice(np.arange(1, 20 * n_snps), n_snps, replace=False)).astype(np.int64)
    start_bp, end_bp = int(bp[core_lo]), int(bp[core_hi - 1])
    return type_, bp, start_bp, end_bp


def chr22_windows(bp_measured: np.ndarray, unmeasured_per_mb: float = 3700.0, core: int = 1_000_000,
                  wing: int = 500_000, seed: int = 22):
    """Window list of a chromosome-wide distmix run: measured SNPs at the given positions (the bundled
    PGC2 chr22 file), synthetic unmeasured sites at ~3,700 / Mb.  Returns the sorted
    site table (bp, type) and per-window index lists in the reference's sense: measured = type 1 in
    [start-wing, end+wing], unmeasured = type 0 in [start, end] (dist.cpp:132-141)."""
    rng = np.random.default_rng(seed)
    lo, hi = int(bp_measured.min()), int(bp_measured.max())
    n_u = int((hi - lo) / 1e6 * unmeasured_per_mb)
    taken = set(int(b) for b in bp_measured)
    cand = np.array(sorted(set(int(b) for b in rng.integers(lo, hi, size=int(n_u * 1.05))) - taken), np.int64)
    bp_u = np.sort(rng.permutation(cand)[:n_u])
    bp = np.concatenate([np.unique(bp_measured.astype(np.int64)), bp_u])
    type_ = np.concatenate([np.ones(len(bp) - len(bp_u), np.int32), np.zeros(len(bp_u), np.int32)])
    order = np.argsort(bp, kind="stable")
    bp, type_ = bp[order], type_[order]
    windows = []
    start = (lo // core) * core
    while start <= hi:
        end = start + core - 1
        meas = np.where((type_ == 1) & (bp >= start - wing) & (bp <= end + wing))[0]
        unme = np.where((type_ == 0) & (bp >= start) & (bp <= end))[0]
        windows.append(dict(start_bp=start, end_bp=end, measured=meas, unmeasured=unme))
        start += core
    return bp, type_, windows


HG19_MB = [249, 243, 198, 191, 181, 171, 159, 146, 141, 136, 135, 134, 115, 107, 103, 90, 81, 78, 59, 63, 48, 51]


def genome_layout(chrom_mb=HG19_MB, measured_per_mb: float = 417.0, unmeasured_per_mb: float = 3470.0,
                  core: int = 1_000_000, wing: int = 500_000, seed: int = 4, density_sd: float = 0.35,
                  z_sd: float = 1.34):
    """Site tables and window lists of a genome-wide distmix run (BASELINE.json config 4: 22 chromosomes, ~1.2 M measured /
    ~10 M target SNPs, 1 Mb windows with 0.5 Mb wings -> ~2,900 windows, n_t ~ 830, n_u ~ 3,470 per window).

    Measured SNPs follow an inhomogeneous density (log-normal factor per Mb, like the 60x spread of window cost along
    the bundled chr22 file); unmeasured sites are uniform.  Per chromosome the panel rows are laid out as
    [measured block | unmeasured block], each in bp order -- the order a feeder that knows the input Z file uploads them
    in -- so every window is two contiguous row ranges.  `sites[r]` is the bp-order rank of row r (the LD chain of the
    synthetic generator runs along it).  Windows are in the reference's sense (dist.cpp:132-141): measured = type 1 in
    [start - wing, end + wing], unmeasured = type 0 in [start, end]."""
    rng = np.random.default_rng(seed)
    chroms = []
    for ci, mb in enumerate(chrom_mb):
        L = int(mb * core)
        n_mb = max(1, int(np.ceil(L / core)))
        fac = np.exp(rng.normal(0.0, density_sd, n_mb))
        fac *= n_mb / fac.sum()
        cnt = rng.poisson(measured_per_mb * core / 1e6 * fac)
        def sorted_unique(a):                      # (np.unique's hash path is ~20x slower on 10^6 int64 keys)
            a = np.sort(a)
            return a[np.concatenate([[True], a[1:] != a[:-1]])] if len(a) else a

        bp_m = sorted_unique(np.concatenate([rng.integers(i * core, min((i + 1) * core, L), size=c)
                                             for i, c in enumerate(cnt)] + [np.zeros(0, np.int64)]))
        n_u = int(unmeasured_per_mb * mb)
        bp_u = sorted_unique(rng.integers(0, L, size=int(n_u * 1.02) + 8))
        if len(bp_m):
            pos = np.minimum(np.searchsorted(bp_m, bp_u), len(bp_m) - 1)
            bp_u = bp_u[bp_m[pos] != bp_u]
        if len(bp_u) > n_u:
            bp_u = np.sort(rng.choice(bp_u, n_u, replace=False))
        n_m, n_u = len(bp_m), len(bp_u)
        # bp-order rank of every row: rank among the union of both sorted lists
        sites = np.concatenate([np.arange(n_m) + np.searchsorted(bp_u, bp_m),
                                np.arange(n_u) + np.searchsorted(bp_m, bp_u)]).astype(np.int64)
        starts = np.arange(0, L, core, dtype=np.int64)
        t_lo = np.searchsorted(bp_m, starts - wing, "left")
        t_hi = np.searchsorted(bp_m, starts + core - 1 + wing, "right")
        u_lo = np.searchsorted(bp_u, starts, "left")
        u_hi = np.searchsorted(bp_u, starts + core - 1, "right")
        t_off = np.concatenate([[0], np.cumsum(t_hi - t_lo)]).astype(np.int64)
        u_off = np.concatenate([[0], np.cumsum(u_hi - u_lo)]).astype(np.int64)
        rows_t = np.concatenate([np.arange(a, b) for a, b in zip(t_lo, t_hi)] + [np.zeros(0, np.int64)]).astype(np.int64)
        rows_u = (n_m + np.concatenate([np.arange(a, b) for a, b in zip(u_lo, u_hi)] + [np.zeros(0, np.int64)])).astype(np.int64)
        z_m = rng.standard_normal(n_m) * z_sd
        chroms.append(dict(chrom=ci + 1, n_rows=n_m + n_u, n_measured=n_m, sites=sites, t_off=t_off, rows_t=rows_t,
                           u_off=u_off, rows_u=rows_u, z_t=z_m[rows_t], start_bp=starts, bp_m=bp_m, bp_u=bp_u))
    return chroms
