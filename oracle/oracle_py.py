"""ctypes binding of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  gauss_b200/ never does.

Two back-ends with the same Python surface:
  * ``port``      -- oracle/libgauss_oracle.so, the plain-C restatement (gauss_oracle.c)
  * ``reference`` -- oracle/_ref/libgauss_ref.so, the reference's own CalCor / CalWgtCov /
                     run_dist / run_distmix compiled from /root/reference by build_ref.sh
                     (present only if that build was run in the authoring container;
                     the prebuilt .so travels to the GPU box, the sources do not).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(_HERE, "libgauss_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libgauss_ref.so")

_c_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_c_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_c_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_c_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


class GoArgs(C.Structure):
    _fields_ = [
        ("start_bp", C.c_longlong),
        ("end_bp", C.c_longlong),
        ("lambda_", C.c_double),
        ("min_abs_eig", C.c_double),
        ("min_num_measured_snp", C.c_int),
        ("min_num_unmeasured_snp", C.c_int),
    ]


def build_port(force: bool = False) -> str:
    """Compile the plain-C restatement (gcc only)."""
    src = os.path.join(_HERE, "gauss_oracle.c")
    if force or not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", src, "-o", PORT_SO, "-lm"]
        )
    return PORT_SO


def _bind(lib):
    lib.go_cal_cor.restype = C.c_double
    lib.go_cal_cor.argtypes = [_c_u8p, _c_u8p, _c_i32p, C.c_int]
    lib.go_cal_wgt_cov.restype = C.c_double
    lib.go_cal_wgt_cov.argtypes = [_c_u8p, _c_u8p, _c_i32p, C.c_int, _c_f64p]
    lib.go_run_window.restype = C.c_int
    lib.go_run_window.argtypes = [
        _c_i32p, _c_i64p, _c_f64p, _c_f64p, _c_u8p, C.c_int64, _c_i32p, C.c_int,
        C.c_void_p, C.POINTER(GoArgs), C.POINTER(C.c_int), C.POINTER(C.c_int),
        C.c_void_p, C.c_void_p,
    ]
    lib.go_compute_ld.restype = None
    lib.go_compute_ld.argtypes = [_c_u8p, C.c_int64, _c_i32p, C.c_int, _c_f64p, _c_f64p]
    lib.go_last_sample_pairs.restype = C.c_double
    lib.go_zmix_pairs.restype = None
    lib.go_zmix_pairs.argtypes = [_c_u8p, C.c_int64, _c_i32p, C.c_int, _c_f64p, _c_f64p]
    return lib


class Oracle:
    """kind = 'port' (C restatement) or 'reference' (reference's own code, if built)."""

    def __init__(self, kind: str = "port"):
        self.kind = kind
        if kind == "port":
            self.lib = _bind(C.CDLL(build_port()))
            lib = self.lib
            lib.go_gram_counts.restype = None
            lib.go_gram_counts.argtypes = [
                _c_u8p, C.c_int64, _c_u8p, C.c_int64, _c_i32p, C.c_int,
                C.c_void_p, C.c_void_p, C.c_void_p,
            ]
            lib.go_sym_eig.restype = C.c_int
            lib.go_sym_eig.argtypes = [_c_f64p, C.c_int, _c_f64p, C.c_void_p]
            lib.go_make_pos_def.restype = C.c_int
            lib.go_make_pos_def.argtypes = [_c_f64p, C.c_int, C.c_double]
            lib.go_count_pc.restype = C.c_int
            lib.go_count_pc.argtypes = [_c_f64p, C.c_int, C.c_double]
            lib.go_inv_full_piv_lu.restype = None
            lib.go_inv_full_piv_lu.argtypes = [_c_f64p, _c_f64p, C.c_int]
        elif kind == "reference":
            if not os.path.exists(REF_SO):
                raise FileNotFoundError(REF_SO + " (run oracle/build_ref.sh where /root/reference exists)")
            self.lib = _bind(C.CDLL(REF_SO))
        else:
            raise ValueError(kind)

    @staticmethod
    def available(kind: str) -> bool:
        return kind == "port" or os.path.exists(REF_SO)

    # -- scalar statistics ---------------------------------------------------
    def cal_cor(self, x, y, m):
        m = np.ascontiguousarray(m, np.int32)
        return self.lib.go_cal_cor(_chars(x), _chars(y), m, len(m))

    def cal_wgt_cov(self, x, y, m, w):
        m = np.ascontiguousarray(m, np.int32)
        w = np.ascontiguousarray(w, np.float64)
        return self.lib.go_cal_wgt_cov(_chars(x), _chars(y), m, len(m), w)

    def gram_counts(self, geno_a, geno_b, m):
        """-> (sxy[P,na,nb], sx_a[P,na], sxx_a[P,na]) int32, brute force."""
        assert self.kind == "port"
        m = np.ascontiguousarray(m, np.int32)
        ga, gb = _chars(geno_a), _chars(geno_b)
        na, nb, P = ga.shape[0], gb.shape[0], len(m)
        sxy = np.zeros((P, na, nb), np.int32)
        sx = np.zeros((P, na), np.int32)
        sxx = np.zeros((P, na), np.int32)
        self.lib.go_gram_counts(ga, na, gb, nb, m, P, sxy.ctypes.data, sx.ctypes.data, sxx.ctypes.data)
        return sxy, sx, sxx

    # -- window kernels --------------------------------------------------------
    def run_window(self, type_, bp, z, geno, m, w=None, start_bp=0, end_bp=0, lam=0.1,
                   min_abs_eig=1e-5, min_measured=10, min_unmeasured=10, dump=False):
        """run_dist (w=None) / run_distmix.  Returns dict(rc, z, info, n_t, n_u[, B11, B21])."""
        type_ = np.ascontiguousarray(type_, np.int32)
        bp = np.ascontiguousarray(bp, np.int64)
        zz = np.array(z, np.float64, copy=True)
        info = np.ones_like(zz)
        g = _chars(geno)
        m = np.ascontiguousarray(m, np.int32)
        a = GoArgs(int(start_bp), int(end_bp), lam, min_abs_eig, min_measured, min_unmeasured)
        nt, nu = C.c_int(0), C.c_int(0)
        wp = None
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
            wp = w.ctypes.data
        B11 = B21 = None
        b11p = b21p = None
        if dump:
            core = (bp >= start_bp) & (bp <= end_bp)
            n_t = int((type_ == 1).sum())
            n_u = int(((type_ == 0) & core).sum())
            B11 = np.zeros((n_t, n_t), np.float64)
            B21 = np.zeros((n_u, n_t), np.float64)
            b11p, b21p = B11.ctypes.data, B21.ctypes.data
        rc = self.lib.go_run_window(type_, bp, zz, info, g, g.shape[0], m, len(m), wp, C.byref(a),
                                    C.byref(nt), C.byref(nu), b11p, b21p)
        out = dict(rc=rc, z=zz, info=info, n_t=nt.value, n_u=nu.value,
                   sample_pairs=self.lib.go_last_sample_pairs())
        if dump:
            out["B11"], out["B21"] = B11, B21
        return out

    def run_qcat(self, type_, bp, z, geno, m, w=None, start_bp=0, end_bp=0, lam=0.1, eig_cutoff=0.01,
                 min_measured=10):
        """run_qcat (w=None) / run_qcatmix.  Returns dict(rc, m, t, chisq) indexed like the SNP list (NaN = untested)."""
        type_ = np.ascontiguousarray(type_, np.int32)
        bp = np.ascontiguousarray(bp, np.int64)
        zz = np.ascontiguousarray(z, np.float64)
        g = _chars(geno)
        m = np.ascontiguousarray(m, np.int32)
        a = GoArgs(int(start_bp), int(end_bp), lam, 1e-5, min_measured, 10)
        wp = None
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
            wp = w.ctypes.data
        qm, qt, qc = (np.full(len(zz), np.nan) for _ in range(3))
        self.lib.go_run_qcat.restype = C.c_int
        self.lib.go_run_qcat.argtypes = [_c_i32p, _c_i64p, _c_f64p, _c_u8p, C.c_int64, _c_i32p, C.c_int, C.c_void_p,
                                         C.POINTER(GoArgs), C.c_double, _c_f64p, _c_f64p, _c_f64p]
        rc = self.lib.go_run_qcat(type_, bp, zz, g, g.shape[0], m, len(m), wp, C.byref(a), eig_cutoff, qm, qt, qc)
        return dict(rc=rc, m=qm, t=qt, chisq=qc)

    def zmix_pairs(self, geno, m, z):
        """prep_zmix5 pair loop -> [n(n-1)/2, 1 + P] (column 0 = z_i z_j, then one Pearson r per population)."""
        g = _chars(geno)
        m = np.ascontiguousarray(m, np.int32)
        z = np.ascontiguousarray(z, np.float64)
        n = g.shape[0]
        out = np.zeros((1 + len(m), n * (n - 1) // 2), np.float64)   # column-major [pairs][1+P]
        self.lib.go_zmix_pairs(g, n, m, len(m), z, out)
        return out.T

    def gene_corg(self, geno, m, w=None, lam=0.1):
        """CorG of one gene exactly as Gene::CalJepegmixPval (w given) / Gene::CalJepegPval (w None) build it
        (gene.cpp:569-586 / 305-315); reference build only -- the port's equivalent is compute_ld / cal_cor + diagonal."""
        assert self.kind == "reference"
        g = _chars(geno)
        m = np.ascontiguousarray(m, np.int32)
        n = g.shape[0]
        out = np.zeros((n, n), np.float64)
        wp = None
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
            wp = w.ctypes.data
        self.lib.go_gene_corg.restype = None
        self.lib.go_gene_corg.argtypes = [_c_u8p, C.c_int64, _c_i32p, C.c_int, C.c_void_p, C.c_double, _c_f64p]
        self.lib.go_gene_corg(g, n, m, len(m), wp, lam, out)
        return out

    def compute_ld(self, geno, m, w):
        g = _chars(geno)
        m = np.ascontiguousarray(m, np.int32)
        w = np.ascontiguousarray(w, np.float64)
        n = g.shape[0]
        cm = np.zeros((n, n), np.float64)
        self.lib.go_compute_ld(g, n, m, len(m), w, cm)
        return cm  # symmetric, so row/col-major agree


def _chars(a):
    """Accept int8 dosages (0/1/2) or uint8 ASCII chars; return ASCII uint8, C-contiguous."""
    a = np.asarray(a)
    if a.dtype == np.int8:
        a = (a.astype(np.int16) + 48).astype(np.uint8)
    elif a.dtype != np.uint8:
        raise TypeError("genotypes must be int8 dosages or uint8 ASCII")
    return np.ascontiguousarray(a)


# ---- file-level reference path (oracle/ref_files.cpp: the reference's own bgzf.c + gauss.cpp readers) ----------------
def ref_write_bgzf_panel(data_path, index_path, rsid, chr_, bp, a1, a2, geno, pop_sizes, af):
    """Index + data file pair written through the reference's bgzf_write (real bgzf_tell offsets in the fpos column).
    geno: [n, N] uint8 chars over ALL populations; af: [n, P] values printed with 6 decimals.  -> fpos per SNP."""
    lib = C.CDLL(REF_SO)
    n = len(bp)
    geno = np.ascontiguousarray(geno, np.uint8)
    m = np.ascontiguousarray(pop_sizes, np.int32)
    af = np.ascontiguousarray(af, np.float64)
    fpos = np.zeros(n, np.int64)

    def strs(v):
        arr = (C.c_char_p * n)()
        arr[:] = [s.encode() for s in v]
        return arr

    lib.go_write_bgzf_panel.restype = C.c_int
    rc = lib.go_write_bgzf_panel(data_path.encode(), index_path.encode(), C.c_int64(n), strs(rsid),
                                 np.ascontiguousarray(chr_, np.int32).ctypes.data_as(C.c_void_p),
                                 np.ascontiguousarray(bp, np.int64).ctypes.data_as(C.c_void_p), strs(a1), strs(a2),
                                 geno.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), C.c_int(len(m)),
                                 af.ctypes.data_as(C.c_void_p), fpos.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"go_write_bgzf_panel failed: {rc}")
    return fpos


def ref_file_distmix(input_file, index_file, data_file, pop_desc_file, chr_, start_bp, end_bp, wing, weights: dict,
                     af1_cutoff=0.01, cap=200000):
    """The reference's distmix() body on files (read_ref_desc -> ... -> ReadGenotype -> run_distmix -> output rows)."""
    lib = C.CDLL(REF_SO)
    names = list(weights)
    arr = (C.c_char_p * len(names))()
    arr[:] = [s.encode() for s in names]
    vals = np.array([weights[k] for k in names], np.float64)
    bp = np.zeros(cap, np.int64)
    af, z, info = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    typ = np.zeros(cap, np.int32)
    sbuf = C.create_string_buffer(cap * 48)
    err = C.create_string_buffer(512)
    nm, na = C.c_int(0), C.c_int(0)
    lib.go_file_distmix.restype = C.c_int
    rc = lib.go_file_distmix(input_file.encode(), index_file.encode(), data_file.encode(), pop_desc_file.encode(),
                             C.c_int(chr_), C.c_longlong(start_bp), C.c_longlong(end_bp), C.c_longlong(wing), arr,
                             vals.ctypes.data_as(C.c_void_p), C.c_int(len(names)), C.c_double(af1_cutoff), C.c_int(cap),
                             bp.ctypes.data_as(C.c_void_p), af.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p),
                             info.ctypes.data_as(C.c_void_p), typ.ctypes.data_as(C.c_void_p), sbuf, C.c_int(len(sbuf)),
                             err, C.c_int(len(err)), C.byref(nm), C.byref(na))
    if rc < 0:
        return dict(rc=rc, error=err.value.decode(), n_measured=nm.value, n_all=na.value)
    toks = [ln.split(" ") for ln in sbuf.value.decode().split("\n") if ln]
    return dict(rc=rc, rsid=[t[0] for t in toks], a1=[t[1] for t in toks], a2=[t[2] for t in toks], bp=bp[:rc].copy(),
                af1mix=af[:rc].copy(), z=z[:rc].copy(), info=info[:rc].copy(), type=typ[:rc].copy(),
                n_measured=nm.value, n_all=na.value)
