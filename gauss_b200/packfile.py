"""On-disk packed panel (`.gbpack`): the cached form of the reference's BGZF text panel.

The reference re-inflates and re-parses `*_geno.gz` for every call (ReadGenotype, gauss.cpp:720-785: one text line per
SNP with one '0'/'1'/'2' string per population followed by one allele frequency per population, gauss.cpp:572-585).
`convert_reference_panel` does that once and stores every SNP as a pack2 row over ALL populations -- 2 bits per dosage,
each population block on a 128-dosage (32-byte) boundary -- so a later call memory-maps the file, picks the byte ranges
of the populations its `pop_flag_vec` selects (init_pop_flag_vec, gauss.cpp:1019-1066) and hands rows that are already
in the layout `gb_panel_append_pack2_host` / `gb_chrom_run_pack2` take.  Host-side I/O only: no statistics are computed
here.  `flip_rows` is the packed-row form of FlipGenotypeVec (util.cpp), which ReadGenotype has switched off
(gauss.cpp:767-776: the alleles of the Z file are matched to the panel instead); it is kept for callers that flip.

File layout: b"GBPACK2\\n", uint64 little-endian header length, JSON header (version, pops, sizes, super_pops, n_rows,
row_bytes), zero padding to a 4096-byte boundary, then n_rows * row_bytes bytes.
"""
from __future__ import annotations

import gzip
import json
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import api

MAGIC = b"GBPACK2\n"
ALIGN = 4096


def read_pop_desc(path: str, upper: bool = True):
    """Population description file (gauss.cpp:970-985): one header line, then `pop n_subjects super_pop [...]`.
    upper=False keeps the names as written (the reference compares them case-sensitively, gauss.cpp:1040, 1101)."""
    pops, sizes, sups = [], [], []
    with open(path) as f:
        next(f)
        for line in f:
            tok = line.split()
            if len(tok) >= 3:
                pops.append(tok[0].upper() if upper else tok[0])
                sizes.append(int(tok[1]))
                sups.append(tok[2].upper() if upper else tok[2])
    return pops, np.array(sizes, np.int32), sups


def _bgzf_blocks(f):
    """Yield the raw deflate payload of every BGZF block (bgzf.c: a gzip member whose extra field carries 'BC' and the
    block size).  Raises ValueError on a member without that field (a plain gzip file)."""
    while True:
        head = f.read(12)
        if not head:
            return
        if len(head) < 12 or head[:4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF block header")
        xlen = struct.unpack("<H", head[10:12])[0]
        extra = f.read(xlen)
        bsize, i = None, 0
        while i + 4 <= len(extra):
            si, slen = extra[i:i + 2], struct.unpack("<H", extra[i + 2:i + 4])[0]
            if si == b"BC" and slen == 2:
                bsize = struct.unpack("<H", extra[i + 4:i + 6])[0]
            i += 4 + slen
        if bsize is None:
            raise ValueError("gzip member without the BGZF 'BC' field")
        payload = f.read(bsize + 1 - 12 - xlen - 8)
        f.read(8)                                   # CRC32 + ISIZE
        yield payload


def iter_panel_lines(path: str, threads: int = 0, batch: int = 512):
    """Text lines of a reference data file.  A BGZF file (what the reference ships, bgzf.c:486-536) is inflated block-
    parallel on `threads` host threads (zlib releases the GIL); anything else falls back to sequential `gzip`."""
    threads = threads or min(16, os.cpu_count() or 1)
    try:
        with open(path, "rb") as f:
            first = next(_bgzf_blocks(f), None)
        is_bgzf = first is not None
    except ValueError:
        is_bgzf = False
    if not is_bgzf:
        with gzip.open(path, "rb") as f:
            yield from f
        return
    tail = b""
    with open(path, "rb") as f, ThreadPoolExecutor(threads) as pool:
        blocks = _bgzf_blocks(f)
        while True:
            chunk = []
            for payload in blocks:
                chunk.append(payload)
                if len(chunk) == batch:
                    break
            if not chunk:
                break
            data = tail + b"".join(pool.map(lambda p: zlib.decompress(p, -15) if p else b"", chunk))
            lines = data.split(b"\n")
            tail = lines.pop()
            for ln in lines:
                yield ln + b"\n"
    if tail:
        yield tail


def _block_bytes(sizes) -> np.ndarray:
    return (np.asarray(sizes, np.int64) + 127) // 128 * 32


def convert_reference_panel(geno_gz: str, pop_desc: str, out_path: str, chunk_rows: int = 4096) -> dict:
    """Reference data file (BGZF, inflated block-parallel; plain gzip also accepted) -> `.gbpack`.  Returns the header."""
    pops, sizes, sups = read_pop_desc(pop_desc)
    P, N = len(pops), int(sizes.sum())
    row_bytes = api.pack2_row_bytes(sizes)
    header = dict(version=1, pops=pops, sizes=[int(x) for x in sizes], super_pops=sups, n_rows=0, row_bytes=row_bytes)

    def header_blob(h):
        js = json.dumps(h).encode()
        blob = MAGIC + np.uint64(len(js)).tobytes() + js
        return blob + b"\0" * (-len(blob) % ALIGN)

    # the header's length must not change when n_rows is filled in: reserve digits
    header["n_rows"] = 10 ** 15
    data_off = len(header_blob(header))
    n_rows = 0
    buf = np.empty((chunk_rows, N), np.uint8)
    with open(out_path, "wb") as out:
        f = iter_panel_lines(geno_gz)
        out.write(b"\0" * data_off)
        fill = 0

        def flush():
            nonlocal fill
            if fill:
                out.write(api.pack2_rows_host(sizes, buf[:fill], is_ascii=True).tobytes())
                fill = 0

        for line in f:
            tok = line.split()
            if not tok:
                continue
            if len(tok) < P:
                raise ValueError(f"line {n_rows + 1}: {len(tok)} fields, expected at least {P} genotype strings")
            off = 0
            for k in range(P):
                s = tok[k]
                if len(s) != sizes[k]:
                    raise ValueError(f"line {n_rows + 1}: population {pops[k]} has {len(s)} genotypes, expected {sizes[k]}")
                buf[fill, off:off + len(s)] = np.frombuffer(s, np.uint8)
                off += len(s)
            fill += 1
            n_rows += 1
            if fill == chunk_rows:
                flush()
        flush()
        header["n_rows"] = n_rows
        blob = header_blob(header)
        blob = blob + b"\0" * (data_off - len(blob))
        assert len(blob) == data_off
        out.seek(0)
        out.write(blob)
    return header


class PackFile:
    """Memory-mapped `.gbpack`."""

    def __init__(self, path: str):
        with open(path, "rb") as f:
            if f.read(len(MAGIC)) != MAGIC:
                raise ValueError(f"{path} is not a GBPACK2 file")
            hlen = int(np.frombuffer(f.read(8), np.uint64)[0])
            self.header = json.loads(f.read(hlen))
        self.pops = self.header["pops"]
        self.super_pops = self.header["super_pops"]
        self.sizes = np.array(self.header["sizes"], np.int32)
        self.n_rows, self.row_bytes = int(self.header["n_rows"]), int(self.header["row_bytes"])
        data_off = -(-(len(MAGIC) + 8 + hlen) // ALIGN) * ALIGN
        if os.path.getsize(path) != data_off + self.n_rows * self.row_bytes:
            raise ValueError(f"{path}: size does not match its header")
        self.rows = np.memmap(path, np.uint8, "r", offset=data_off, shape=(self.n_rows, self.row_bytes))
        self._boff = np.concatenate([[0], np.cumsum(_block_bytes(self.sizes))])

    def flags_for(self, study_pop: str | None = None, weights: dict | None = None) -> np.ndarray:
        """pop_flag_vec as the reference builds it: by population or super-population name (init_pop_flag_vec,
        gauss.cpp:1019-1066) or by the populations a weight table names (init_pop_flag_wgt_vec, gauss.cpp:1093-1117)."""
        if weights is not None:
            names = {k.upper() for k in weights}
            return np.array([p in names for p in self.pops])
        sp = study_pop.upper()
        flags = np.array([p == sp or s == sp for p, s in zip(self.pops, self.super_pops)])
        if not flags.any():
            raise ValueError(f"invalid population name '{study_pop}'")   # the reference stops here too (gauss.cpp:1060)
        return flags

    def weights_for(self, weights: dict) -> np.ndarray:
        """pop_wgt_vec as init_pop_flag_wgt_vec builds it (gauss.cpp:1093-1117): the weights of the flagged populations in
        PANEL order -- the order every pop_wgt argument of the C-ABI expects."""
        wmap = {k.upper(): float(v) for k, v in weights.items()}
        return np.array([wmap[p] for p in self.pops if p in wmap], np.float64)

    def select(self, row_idx, flags, out: np.ndarray | None = None):
        """pack2 rows of the listed SNPs restricted to the flagged populations -> (rows2, flagged sizes).

        A population block is a whole number of 32-byte units in both layouts, so this is a byte-range gather."""
        flags = np.asarray(flags, bool)
        idx = np.asarray(row_idx, np.int64)
        sizes = self.sizes[flags]
        rb = api.pack2_row_bytes(sizes)
        if out is None:
            out = np.empty((len(idx), rb), np.uint8)
        assert out.shape == (len(idx), rb)
        src = self.rows[idx] if len(idx) else np.empty((0, self.row_bytes), np.uint8)
        o = 0
        for k in np.where(flags)[0]:
            n = int(self._boff[k + 1] - self._boff[k])
            out[:, o:o + n] = src[:, self._boff[k]:self._boff[k + 1]]
            o += n
        out[:, o:] = 0
        return out, sizes


_FLIP = np.zeros(256, np.uint8)
for _b in range(256):
    _v = 0
    for _q in range(4):
        _c = (_b >> (2 * _q)) & 3
        _v |= ((2 - _c) & 3 if _c < 3 else 3) << (2 * _q)
    _FLIP[_b] = _v


def flip_rows(rows2: np.ndarray, sizes, which) -> None:
    """In place: dosage x -> 2 - x for the rows in `which` (FlipGenotypeVec, util.cpp; disabled in the reference's
    ReadGenotype, gauss.cpp:767-776).  Padding stays zero."""
    which = np.asarray(which)
    if which.dtype != bool:
        m = np.zeros(rows2.shape[0], bool)
        m[which] = True
        which = m
    if not which.any():
        return
    sub = _FLIP[rows2[which]]
    boff = np.concatenate([[0], np.cumsum(_block_bytes(sizes))])
    for k, m in enumerate(np.asarray(sizes, np.int64)):
        full, rem = divmod(int(m), 4)
        start = int(boff[k]) + full
        if rem:
            sub[:, start] &= (1 << (2 * rem)) - 1      # dosages past the population's size in its last byte
            start += 1
        sub[:, start:int(boff[k + 1])] = 0               # padding bytes of the block
    sub[:, int(boff[-1]):] = 0
    rows2[which] = sub


# ---- native converter + ternary file (gb_packfile_convert, gauss_b200/csrc/gb_packfile.cu) ------------------------------
MAGIC5 = b"GBPACK5\n"


def convert_reference_panel_native(geno_gz: str, pop_desc: str, out_path: str, threads: int = 0) -> dict:
    """Reference data file (BGZF) -> `.gbpack` with ternary rows, per-population allele frequencies and the BGZF
    virtual offset of every line, through the library's threaded C++ converter.  -> dict(n_rows, text_bytes, seconds)."""
    import ctypes as C
    _, sizes, _ = read_pop_desc(pop_desc)
    lib = api.load_library()
    n, tb, sec = C.c_int64(), C.c_double(), C.c_double()
    err = C.create_string_buffer(512)
    rc = lib.gb_packfile_convert(geno_gz.encode(), len(sizes), sizes.ctypes.data, out_path.encode(), int(threads),
                                 C.byref(n), C.byref(tb), C.byref(sec), err, len(err))
    if rc != api.GB_OK:
        raise api.GaussB200Error(rc, err.value.decode())
    return dict(n_rows=n.value, text_bytes=tb.value, seconds=sec.value)


class PackFile5:
    """Memory-mapped ternary `.gbpack` (layout: gb_packfile.cu).  rows[r] is the pack5 row of data line r over ALL
    populations, af1[r] its per-population allele frequencies, fpos[r] the BGZF virtual offset of the line -- the key
    the reference's index file refers to a SNP by (gauss.cpp:324-330)."""

    def __init__(self, path: str, pop_desc: str | None = None):
        with open(path, "rb") as f:
            head = f.read(72)
        if head[:8] != MAGIC5:
            raise ValueError(f"{path} is not a GBPACK5 file")
        self.path = path
        ver, n_rows, n_pops, row_bytes, off_sizes, off_rows, off_fpos, off_af1 = np.frombuffer(head[8:], np.uint64)
        self.n_rows, self.n_pops, self.row_bytes = int(n_rows), int(n_pops), int(row_bytes)
        self.sizes = np.array(np.memmap(path, np.int32, "r", offset=int(off_sizes), shape=(self.n_pops,)))
        if self.row_bytes != api.pack5_row_bytes(self.sizes):
            raise ValueError(f"{path}: row size does not match its population sizes")
        self.rows = np.memmap(path, np.uint8, "r", offset=int(off_rows), shape=(self.n_rows, self.row_bytes))
        self.fpos = np.array(np.memmap(path, np.int64, "r", offset=int(off_fpos), shape=(self.n_rows,)))
        self.af1 = np.memmap(path, np.float64, "r", offset=int(off_af1), shape=(self.n_rows, self.n_pops))
        self.pops = self.super_pops = None
        if pop_desc is not None:
            self.pops, sizes, self.super_pops = read_pop_desc(pop_desc, upper=False)
            if not np.array_equal(sizes, self.sizes):
                raise ValueError("population description does not match the packed file")
        nb = (self.sizes.astype(np.int64) + 4) // 5
        self._boff = np.concatenate([[0], np.cumsum((nb + 3) // 4 * 4)])
        self._order = np.argsort(self.fpos, kind="stable")

    def rows_of_fpos(self, fpos) -> np.ndarray:
        """Row index of every virtual offset (-1 when the file has no line there)."""
        fpos = np.asarray(fpos, np.int64)
        srt = self.fpos[self._order]
        i = np.minimum(np.searchsorted(srt, fpos), max(self.n_rows - 1, 0))
        hit = (srt[i] == fpos) if self.n_rows else np.zeros(len(fpos), bool)
        return np.where(hit, self._order[i], -1)

    def select(self, row_idx, flags, out: np.ndarray | None = None):
        """pack5 rows of the listed SNPs restricted to the flagged populations -> (rows5, flagged sizes).  Population
        blocks are whole 4-byte units in both layouts, so this is a byte-range gather."""
        flags = np.asarray(flags, bool)
        idx = np.asarray(row_idx, np.int64)
        sizes = self.sizes[flags]
        rb = api.pack5_row_bytes(sizes)
        if out is None:
            out = np.zeros((len(idx), rb), np.uint8)
        assert out.shape == (len(idx), rb)
        src = self.rows[idx] if len(idx) else np.empty((0, self.row_bytes), np.uint8)
        o = 0
        for k in np.where(flags)[0]:
            n = int(self._boff[k + 1] - self._boff[k])
            out[:, o:o + n] = src[:, self._boff[k]:self._boff[k + 1]]
            o += n
        out[:, o:] = 0
        return out, sizes
