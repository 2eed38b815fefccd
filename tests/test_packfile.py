"""`.gbpack`: converter from the reference's panel text format, flagged-population
selection and the packed-row flip, all host-side; one GPU test that rows selected from the file give the counts of the
raw genotypes."""
import gzip

import numpy as np
import pytest

import gauss_b200 as gb
from gauss_b200 import api, packfile, synth


def write_reference_panel(tmp_path, g, pops, sizes, sups):
    """The reference's data file (gauss.cpp:572-585): per SNP one line of P genotype strings and P allele frequencies;
    written as two gzip members like a BGZF file's blocks."""
    desc = tmp_path / "pop_desc.txt"
    desc.write_text("Population\tNumber\tSuper\tDescription\n" +
                    "".join(f"{p}\t{n}\t{s}\tsynthetic population\n" for p, n, s in zip(pops, sizes, sups)))
    offs = np.concatenate([[0], np.cumsum(sizes)])
    lines = []
    for row in g:
        strs = ["".join(map(str, row[offs[k]:offs[k + 1]])) for k in range(len(sizes))]
        afs = [f"{row[offs[k]:offs[k + 1]].mean() / 2:.6f}" for k in range(len(sizes))]
        lines.append(" ".join(strs + afs) + "\n")
    path = tmp_path / "panel_geno.gz"
    half = len(lines) // 2
    with open(path, "wb") as f:
        f.write(gzip.compress("".join(lines[:half]).encode()))
        f.write(gzip.compress("".join(lines[half:]).encode()))
    return str(path), str(desc)


def write_bgzf(path, data: bytes, block: int = 3000):
    """A BGZF file as bgzf.c writes it: gzip members with the 'BC' extra field, plus the empty EOF block."""
    import struct
    import zlib
    with open(path, "wb") as f:
        chunks = [data[i:i + block] for i in range(0, len(data), block)] + [b""]
        for c in chunks:
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            payload = co.compress(c) + co.flush()
            bsize = 12 + 6 + len(payload) + 8 - 1
            f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize))
            f.write(payload + struct.pack("<II", zlib.crc32(c), len(c)))


def test_bgzf_panel_is_inflated_block_parallel(tmp_path):
    g = synth.make_genotypes(120, SIZES, seed=93).astype(np.int8)
    geno, desc = write_reference_panel(tmp_path, g, POPS, SIZES, SUPS)
    text = gzip.open(geno, "rb").read()
    bg = str(tmp_path / "panel_bgzf_geno.gz")
    write_bgzf(bg, text)
    assert gzip.open(bg, "rb").read() == text                        # a BGZF file is also a valid gzip file
    assert b"".join(packfile.iter_panel_lines(bg, threads=4, batch=7)) == text
    assert b"".join(packfile.iter_panel_lines(geno)) == text          # plain gzip: sequential fallback
    out = str(tmp_path / "p.gbpack")
    packfile.convert_reference_panel(bg, desc, out)
    np.testing.assert_array_equal(np.asarray(packfile.PackFile(out).rows), api.pack2_rows_host(SIZES, g))


POPS = ["CEU", "GBR", "YRI", "JPT", "CHB", "MXL"]
SUPS = ["EUR", "EUR", "AFR", "ASN", "ASN", "AMR"]
SIZES = np.array([61, 130, 7, 128, 33, 2], np.int32)


def test_convert_select_and_flip(tmp_path):
    g = synth.make_genotypes(300, SIZES, seed=91).astype(np.int8)
    geno, desc = write_reference_panel(tmp_path, g, POPS, SIZES, SUPS)
    out = str(tmp_path / "panel.gbpack")
    hdr = packfile.convert_reference_panel(geno, desc, out, chunk_rows=64)
    assert hdr["n_rows"] == 300 and hdr["pops"] == POPS
    pf = packfile.PackFile(out)
    assert pf.n_rows == 300 and pf.row_bytes == api.pack2_row_bytes(SIZES)
    np.testing.assert_array_equal(np.asarray(pf.rows), api.pack2_rows_host(SIZES, g))
    # study_pop = super-population (init_pop_flag_vec) and a weight table (init_pop_flag_wgt_vec)
    offs = np.concatenate([[0], np.cumsum(SIZES)])
    for flags in (pf.flags_for("EUR"), pf.flags_for("jpt"), pf.flags_for(weights={"YRI": .2, "chb": .3, "MXL": .5})):
        assert flags.any()
        idx = np.array([5, 0, 299, 17])
        rows2, sizes = pf.select(idx, flags)
        cols = np.concatenate([np.arange(offs[k], offs[k + 1]) for k in np.where(flags)[0]])
        np.testing.assert_array_equal(sizes, SIZES[flags])
        np.testing.assert_array_equal(rows2, api.pack2_rows_host(SIZES[flags], g[idx][:, cols]))
    np.testing.assert_array_equal(pf.flags_for("EUR"), [True, True, False, False, False, False])
    # flip: dosage -> 2 - dosage on the chosen rows, padding untouched
    rows2 = np.array(pf.rows)
    packfile.flip_rows(rows2, SIZES, [3, 10])
    g2 = g.copy()
    g2[[3, 10]] = 2 - g2[[3, 10]]
    np.testing.assert_array_equal(rows2, api.pack2_rows_host(SIZES, g2))
    # a damaged file is refused
    with open(out, "r+b") as f:
        f.truncate(1000)
    with pytest.raises(ValueError):
        packfile.PackFile(out)
    bad = tmp_path / "bad_geno.gz"
    bad.write_bytes(gzip.compress(b"012 01\n"))
    with pytest.raises(ValueError):
        packfile.convert_reference_panel(str(bad), desc, str(tmp_path / "x.gbpack"))


@pytest.mark.gpu
def test_rows_from_packfile_feed_the_gpu(gpu_ctx, tmp_path):
    g = synth.make_genotypes(200, SIZES, seed=92).astype(np.int8)
    geno, desc = write_reference_panel(tmp_path, g, POPS, SIZES, SUPS)
    out = str(tmp_path / "panel.gbpack")
    packfile.convert_reference_panel(geno, desc, out)
    pf = packfile.PackFile(out)
    flags = pf.flags_for(weights={"CEU": .5, "YRI": .2, "JPT": .3})
    rows2, sizes = pf.select(np.arange(200), flags)
    panel = gb.Panel(gpu_ctx, sizes, 200, "e2m1")
    panel.append_pack2_host(rows2)
    offs = np.concatenate([[0], np.cumsum(SIZES)])
    cols = np.concatenate([np.arange(offs[k], offs[k + 1]) for k in np.where(flags)[0]])
    ref = gb.Panel(gpu_ctx, sizes, 200, "int8")
    ref.append_host(g[:, cols], is_ascii=False)
    a, b = np.arange(0, 150), np.arange(50, 200)
    for x, y in zip(panel.gram_counts(a, b), ref.gram_counts(a, b)):
        np.testing.assert_array_equal(x, y)
