"""File-level parity (SURVEY.md section 4 iv): a synthetic panel written through the reference's own bgzf_write, read by
the reference's own ReadInputZ -> ReadReferenceIndex -> MakeSnpVecMix -> ReadGenotype -> run_distmix (oracle/_ref:
bgzf.c + gauss.cpp compiled unmodified) against the product path: native converter -> .gbpack -> host feeder -> GPU."""
import gzip
import os

import numpy as np
import pytest

from gauss_b200 import api, packfile, synth

POPS = ["ACB", "CEU", "FIN", "GBR", "JPT", "YRI"]
SUPS = ["AFR", "EUR", "EUR", "EUR", "ASN", "AFR"]
SIZES = np.array([37, 161, 88, 95, 23, 52], np.int32)
WEIGHTS = {"ceu": 0.41, "FIN": 0.22, "Gbr": 0.30, "ACB": 0.131}     # distmix upper-cases the caller's names; sum != 1
CHR, START, END, WING = 22, 20_000_000, 21_000_000, 500_000


def ref_lib():
    from oracle import oracle_py
    if not oracle_py.Oracle.available("reference"):
        pytest.skip("oracle/_ref/libgauss_ref.so not built (needs /root/reference at build time)")
    return oracle_py


def make_case(tmp_path, n_snps=900, seed=5):
    """Panel files written by the REFERENCE's writer + an input Z file exercising every branch of ReadReferenceIndex."""
    op = ref_lib()
    rng = np.random.default_rng(seed)
    g = synth.make_genotypes(n_snps, SIZES, seed=seed)
    g[10] = 0                                   # monomorphic everywhere: af1_mix = 0 -> dropped by the AF filter
    g[11, :] = 0
    g[11, :3] = 1                               # af1_mix far below the cut-off
    bp = np.sort(rng.choice(np.arange(START - WING - 100_000, END + WING + 100_000), n_snps, replace=False)).astype(np.int64)
    if n_snps > 201:
        bp[201] = bp[200]                       # two panel SNPs on one position (different alleles)
    alle = np.array(list("ACGT"))
    a1 = alle[rng.integers(0, 4, n_snps)]
    a2 = alle[(np.searchsorted(alle, a1) + rng.integers(1, 4, n_snps)) % 4]
    if n_snps > 201:
        a1[200], a2[200], a1[201], a2[201] = "A", "C", "A", "G"
    rsid = [f"rs{1000 + i}" for i in range(n_snps)]
    offs = np.concatenate([[0], np.cumsum(SIZES)])
    af = np.stack([g[:, offs[k]:offs[k + 1]].mean(1) / 2 for k in range(len(SIZES))], 1)
    desc = tmp_path / "pop_desc.txt"
    desc.write_text("Population\tNumber\tSuper\tDescription\n" +
                    "".join(f"{p}\t{n}\t{s}\tsynthetic\n" for p, n, s in zip(POPS, SIZES, SUPS)))
    data, index = str(tmp_path / "panel_geno.gz"), str(tmp_path / "panel_index.gz")
    fpos = op.ref_write_bgzf_panel(data, index, rsid, np.full(n_snps, CHR), bp, a1, a2, (g + 48).astype(np.uint8), SIZES, af)
    # input Z: every third panel SNP measured; some with swapped alleles (-> z negated), one absent from the panel
    # (type 2), one on another chromosome, one outside the extended window
    meas = np.arange(0, n_snps, 3)
    lines = ["rsid chr bp a1 a2 z\n"]
    zin = rng.standard_normal(n_snps) * 1.34
    for i in meas:
        x1, x2 = (a2[i], a1[i]) if i % 9 == 0 else (a1[i], a2[i])
        lines.append(f"in{rsid[i]} {CHR} {bp[i]} {x1} {x2} {zin[i]:.6f}\n")
    lines.append(f"rsT2 {CHR} {START + 777} A T 1.5\n")
    lines.append(f"rsOther 21 {START + 5} A T 2.5\n")
    lines.append(f"rsFar {CHR} {END + WING + 50} A T 0.5\n")
    zfile = tmp_path / "input_z.txt"
    zfile.write_text("".join(lines))
    return dict(g=g, bp=bp, a1=a1, a2=a2, rsid=rsid, af=af, fpos=fpos, data=data, index=index, desc=str(desc), zfile=str(zfile))


def test_native_converter_reads_what_the_reference_writer_wrote(tmp_path):
    c = make_case(tmp_path)
    out = str(tmp_path / "panel.gbpack")
    info = packfile.convert_reference_panel_native(c["data"], c["desc"], out, threads=3)
    assert info["n_rows"] == len(c["g"]) and info["text_bytes"] == len(gzip.open(c["data"], "rb").read())
    pf = packfile.PackFile5(out, c["desc"])
    np.testing.assert_array_equal(np.asarray(pf.rows), api.pack5_rows_host(SIZES, c["g"], is_ascii=False))
    np.testing.assert_array_equal(pf.fpos, c["fpos"])                      # bgzf_tell of every line, as the index file has it
    np.testing.assert_array_equal(np.asarray(pf.af1), np.round(c["af"], 6))  # the doubles `buffer >> af1` would read
    assert pf.pops == POPS and (pf.rows_of_fpos(c["fpos"][[5, 0, 77]]) == [5, 0, 77]).all()
    assert pf.rows_of_fpos([c["fpos"][3] + 1, -1]).tolist() == [-1, -1]
    flags = np.array([True, False, True, True, False, False])
    rows5, sizes = pf.select([4, 9, 2], flags)
    cols = np.concatenate([np.arange(o, o + m) for o, m, f in zip(np.concatenate([[0], np.cumsum(SIZES)]), SIZES, flags) if f])
    np.testing.assert_array_equal(rows5, api.pack5_rows_host(sizes, c["g"][[4, 9, 2]][:, cols], is_ascii=False))
    # single-threaded and many-threaded runs write the same file
    out1 = str(tmp_path / "panel1.gbpack")
    packfile.convert_reference_panel_native(c["data"], c["desc"], out1, threads=1)
    assert open(out, "rb").read() == open(out1, "rb").read()


def test_native_converter_rejects_what_it_cannot_hold(tmp_path):
    c = make_case(tmp_path, n_snps=40)
    with pytest.raises(api.GaussB200Error):
        packfile.convert_reference_panel_native(c["zfile"], c["desc"], str(tmp_path / "x.gbpack"))   # not a BGZF file
    bad = tmp_path / "short_desc.txt"
    bad.write_text("h\n" + "".join(f"{p}\t{n + 1}\t{s}\n" for p, n, s in zip(POPS, SIZES, SUPS)))
    with pytest.raises(api.GaussB200Error, match="genotypes, expected"):
        packfile.convert_reference_panel_native(c["data"], str(bad), str(tmp_path / "y.gbpack"))


def test_reference_reader_pipeline_runs_on_its_own_files(tmp_path):
    """The reference's readers on files its writer made: branches of ReadReferenceIndex and the AF filter."""
    op = ref_lib()
    c = make_case(tmp_path)
    r = op.ref_file_distmix(c["zfile"], c["index"], c["data"], c["desc"], CHR, START, END, WING, WEIGHTS)
    from gauss_b200 import feeder
    assert r["rc"] > 100 and {0, 1} <= set(r["type"]) <= {0, 1, 2}
    in_core = (c["bp"] >= START) & (c["bp"] <= END)
    assert "rs1010" not in r["rsid"] and "rs1011" not in r["rsid"]           # AF-filtered SNPs
    # type 2 (in the Z file, not in the panel): bgzf_seek(fp, -1) = fseeko(2^48 - 1).  Where the file system refuses that
    # offset the reader does not move and the SNP survives with the NEXT line's frequencies; elsewhere it reads an empty
    # line and is dropped (see feeder.distmix_from_files)
    assert ("rsT2" in r["rsid"]) == feeder.seek_beyond_eof_fails(c["data"])
    # a swapped-allele measured SNP comes back with the panel's alleles and its z negated (gauss.cpp:358-370)
    sw = [i for i in range(0, len(c["bp"]), 9) if in_core[i] and c["rsid"][i] in r["rsid"]]
    assert sw
    zin = {ln.split()[0]: float(ln.split()[5]) for ln in open(c["zfile"]).read().split("\n")[1:] if ln}
    for i in sw[:5]:
        k = r["rsid"].index(c["rsid"][i])
        assert r["type"][k] == 1 and r["a1"][k] == c["a1"][i] and r["z"][k] == -zin["in" + c["rsid"][i]]
    assert list(r["bp"]) == sorted(r["bp"])


@pytest.mark.gpu
def test_file_level_parity_with_the_reference_reader(tmp_path, gpu_ctx):
    from gauss_b200 import feeder
    op = ref_lib()
    c = make_case(tmp_path)
    ref = op.ref_file_distmix(c["zfile"], c["index"], c["data"], c["desc"], CHR, START, END, WING, WEIGHTS)
    assert ref["rc"] > 0
    out = str(tmp_path / "panel.gbpack")
    packfile.convert_reference_panel_native(c["data"], c["desc"], out)
    pf = packfile.PackFile5(out, c["desc"])
    got = feeder.distmix_from_files(gpu_ctx, pf, c["zfile"], c["index"], CHR, START, END, WING, WEIGHTS)
    assert [s["rsid"] for s in got] == ref["rsid"]                           # same SNPs, same (map-key) order
    assert [s["a1"] for s in got] == ref["a1"] and [s["a2"] for s in got] == ref["a2"]
    np.testing.assert_array_equal([s["bp"] for s in got], ref["bp"])
    np.testing.assert_array_equal([s["type"] for s in got], ref["type"])
    np.testing.assert_array_equal([s["af1mix"] for s in got], ref["af1mix"])  # same doubles, same summation order
    dz = np.abs(np.array([s["z"] for s in got]) - ref["z"]).max()
    di = np.abs(np.array([s["info"] for s in got]) - ref["info"]).max()
    assert dz <= 1e-6 and di <= 1e-6                                         # north-star bar
    assert dz <= 1e-9 and di <= 1e-9                                         # achieved
    # too few SNPs: the reference throws "Not enough number of SNPs loaded"; so does the feeder
    few = op.ref_file_distmix(c["zfile"], c["index"], c["data"], c["desc"], CHR, START, START + 2000, 1000, WEIGHTS)
    assert few["rc"] == -1 and "Not enough" in few["error"]
    with pytest.raises(api.GaussB200Error, match="Not enough"):
        feeder.distmix_from_files(gpu_ctx, pf, c["zfile"], c["index"], CHR, START, START + 2000, 1000, WEIGHTS)
