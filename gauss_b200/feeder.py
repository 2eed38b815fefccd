"""Host-side feeder of the window hot path: the reference's file readers on a cached packed panel.

distmix() (distmix.cpp:30-135) spends its non-kernel time in four readers of gauss.cpp -- ReadInputZ (121-190),
ReadReferenceIndex (293-399: allele match / swap -> z negated), MakeSnpVecMix (631-693: the AF filter) and ReadGenotype
(720-785: bgzf_seek + ~33 KB of text per SNP).  BGZF decode, allele matching and the AF filter stay host-side I/O (north
star); what changes is that the genotypes come from a `.gbpack` file (packfile.PackFile5: ternary rows + per-population
allele frequencies + the line's BGZF offset) instead of being re-inflated and re-parsed on every call.  This module is
the mirror of that call sequence; all arithmetic on genotypes happens on the GPU behind the C-ABI.
"""
from __future__ import annotations

import gzip
import math

import numpy as np

from . import api
from .packfile import PackFile5, read_pop_desc


def init_pop_flag_wgt_vec(pops, weights: dict):
    """gauss.cpp:1093-1117 with distmix.cpp:48-54: the caller's population names are upper-cased, the panel's are compared
    as written; flags over all panel populations, weights of the flagged ones in PANEL order."""
    wmap = {k.upper(): float(v) for k, v in weights.items()}
    flags = np.array([p in wmap for p in pops])
    return flags, np.array([wmap[p] for p in pops if p in wmap], np.float64)


def init_pop_flag_vec(pops, super_pops, study_pop: str):
    """gauss.cpp:1019-1066: match against population or super-population name; an unknown name is an error
    ("invalid population name") in the reference."""
    flags = np.array([p == study_pop or s == study_pop for p, s in zip(pops, super_pops)])
    if not flags.any():
        raise ValueError(f"invalid population name '{study_pop}'")
    return flags


def read_input_z(path: str, chr_: int, lo: int, hi: int) -> dict:
    """ReadInputZ (gauss.cpp:121-190): `rsid chr bp a1 a2 z`, header skipped, filtered to the chromosome and
    [start - wing, end + wing]; a later line with the same (chr, bp, a1, a2) replaces an earlier one.  Type 2, info 1."""
    snps = {}
    with open(path) as f:
        next(f, None)
        for line in f:
            t = line.split()
            if len(t) < 6:
                continue
            c, bp = int(t[1]), int(t[2])
            if (chr_ > 0 and c != chr_) or bp < lo or bp > hi:
                continue
            snps[(c, bp, t[3], t[4])] = dict(rsid=t[0], chr=c, bp=bp, a1=t[3], a2=t[4], z=float(t[5]), info=1.0, type=2, fpos=-1)
    return snps


def read_reference_index(path: str, snps: dict, chr_: int, lo: int, hi: int) -> None:
    """ReadReferenceIndex (gauss.cpp:293-399) on the BGZF index text (a BGZF file is a multi-member gzip file)."""
    with gzip.open(path, "rt") as f:
        for line in f:
            t = line.split()
            if len(t) < 7:
                continue
            c, bp = int(t[1]), int(t[2])
            if (chr_ > 0 and c != chr_) or bp < lo or bp > hi:
                continue
            rsid, a1, a2, fpos = t[0], t[3], t[4], int(t[6])
            k1, k2 = (c, bp, a1, a2), (c, bp, a2, a1)
            in1, in2 = k1 in snps, k2 in snps
            if in1 and not in2:                                  # same allele order
                snps[k1].update(rsid=rsid, type=1, fpos=fpos)
            elif in2 and not in1:                                # swapped: z negated, alleles rewritten (gauss.cpp:358-370)
                s = snps.pop(k2)
                s.update(rsid=rsid, a1=a1, a2=a2, z=-s["z"], type=1, fpos=fpos)
                snps[k1] = s
            elif not in1 and not in2:                            # unmeasured SNP of the panel
                snps[k1] = dict(rsid=rsid, chr=c, bp=bp, a1=a1, a2=a2, z=0.0, info=-1.0, type=0, fpos=fpos)
            else:
                raise ValueError("ERROR: input file contains duplicates")   # gauss.cpp:390


def seek_beyond_eof_fails(path: str) -> bool:
    """Whether fseeko(file, 2^48 - 1) fails on the file system holding `path` (ext4 / overlayfs: EINVAL beyond 16 TiB;
    tmpfs / xfs: succeeds).  This decides what the reference does with a type 2 SNP, see distmix_from_files."""
    import os
    fd = os.open(path, os.O_RDONLY)
    try:
        os.lseek(fd, 0xFFFFFFFFFFFF, os.SEEK_SET)
        return False
    except OSError:
        return True
    finally:
        os.close(fd)


def distmix_from_files(ctx: api.Context, pack: PackFile5, input_file: str, index_file: str, chr_: int, start_bp: int,
                       end_bp: int, wing: int, weights: dict, af1_cutoff: float = 0.01, params: api.Params | None = None,
                       type2_reads_next_line: bool | None = None):
    """distmix() on a packed panel: returns the rows the reference's data frame would hold (distmix.cpp:100-133).

    Type 2 SNPs (in the Z file, not in the panel) keep fpos = -1 (snp.cpp:31).  MakeSnpVecMix still calls
    bgzf_seek(fp, -1): that is fseeko(file, 2^48 - 1).  Where the file system accepts the offset the read returns an
    empty line, every frequency parses as 0 and the SNP is dropped by the AF filter; where it refuses (ext4, overlayfs)
    bgzf_seek returns -1 WITHOUT moving, BgzfGetLine reads the line FOLLOWING the previously read SNP's line, and the
    SNP is kept (type 2, input z, info 1) with that other line's allele frequencies in af1mix.  Both are reproduced;
    type2_reads_next_line = None probes the file system of the packed panel."""
    flags, w = init_pop_flag_wgt_vec(pack.pops, weights)
    lo, hi = start_bp - wing, end_bp + wing
    snps = read_input_z(input_file, chr_, lo, hi)
    read_reference_index(index_file, snps, chr_, lo, hi)
    keys = sorted(snps)                                          # std::map order: (chr, bp, a1, a2), gauss.h:77-91
    # MakeSnpVecMix: af1_mix = sum_k af1_k w_k over the flagged populations in panel order; the filter drops everything
    # else -- including type 2 SNPs, whose fpos of -1 reads an empty line (all frequencies 0)
    fpos = np.array([snps[k]["fpos"] for k in keys], np.int64)
    rows = pack.rows_of_fpos(fpos)
    if type2_reads_next_line is None:
        type2_reads_next_line = seek_beyond_eof_fails(pack.path)
    af_rows = rows.copy()                      # the data line whose frequencies MakeSnpVecMix parses for each SNP
    if type2_reads_next_line:
        nxt = 0                                # file position of the reader = the line after the last one it read
        for i in range(len(keys)):
            if fpos[i] < 0:
                af_rows[i] = nxt if nxt < pack.n_rows else -1
            nxt = af_rows[i] + 1 if af_rows[i] >= 0 else nxt
    af = np.zeros((len(keys), int(flags.sum())))
    have = af_rows >= 0
    af[have] = np.asarray(pack.af1[af_rows[have]])[:, flags]
    af1_mix = np.zeros(len(keys))
    for k in range(af.shape[1]):
        af1_mix = af1_mix + af[:, k] * w[k]
    keep = (af1_mix > af1_cutoff) & (af1_mix < (1 - af1_cutoff))
    vec = [dict(snps[k], af1mix=float(a), row=int(r)) for k, a, r, ok in zip(keys, af1_mix, rows, keep) if ok]
    # run_distmix's split (distmix.cpp:141-149)
    meas = [i for i, s in enumerate(vec) if s["type"] == 1]
    unme = [i for i, s in enumerate(vec) if s["type"] == 0 and start_bp <= s["bp"] <= end_bp]
    p = params or api.Params.default()
    if len(meas) <= p.min_num_measured_snp or len(unme) <= p.min_num_unmeasured_snp:
        raise api.GaussB200Error(api.GB_ERR_TOO_FEW_MEASURED, "Not enough number of SNPs loaded - DISTMIX not performed")
    # ReadGenotype replaced: the packed rows of exactly these SNPs, flagged populations only
    sel = np.array([vec[i]["row"] for i in meas + unme], np.int64)
    rows5, sizes = pack.select(sel, flags)
    panel = api.Panel(ctx, sizes, len(sel), "e2m1")
    panel.append_pack5_host(rows5)
    z_u, info_u, _ = panel.window_distmix(np.arange(len(meas)), len(meas) + np.arange(len(unme)),
                                          np.array([vec[i]["z"] for i in meas]), w, params)
    panel.close()
    for j, i in enumerate(unme):
        vec[i]["z"], vec[i]["info"] = float(z_u[j]), float(info_u[j])
    out = [s for s in vec if start_bp <= s["bp"] <= end_bp]
    for s in out:
        s["pval"] = math.erfc(abs(s["z"]) / math.sqrt(2.0))      # 2 * pnorm5(|z|, 0, 1, lower = 0, log = 0)
    return out
