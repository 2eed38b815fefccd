"""ctypes binding of include/gauss_b200.h.  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
import weakref
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GB_OK = 0
GB_ERR_BAD_ARG, GB_ERR_CUDA, GB_ERR_NO_DEVICE, GB_ERR_OOM = 1, 2, 3, 4
GB_ERR_TOO_FEW_MEASURED, GB_ERR_TOO_FEW_UNMEASURED, GB_ERR_NOT_PD, GB_ERR_UNSUPPORTED = 5, 6, 7, 8


PANEL_FORMATS = {"int8": 0, "e2m1": 1}   # enum gb_panel_format


class GaussB200Error(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"gauss_b200 status {status}: {msg}")
        self.status = status


class Params(C.Structure):
    """struct gb_params (hidden arguments of the reference, gauss.cpp:18-35)."""
    _fields_ = [("lambda_", C.c_double), ("min_abs_eig", C.c_double), ("min_num_measured_snp", C.c_int),
                ("min_num_unmeasured_snp", C.c_int), ("check_pd", C.c_int), ("reserved", C.c_int)]

    @staticmethod
    def default() -> "Params":
        p = Params()
        load_library().gb_params_default(C.byref(p))
        return p


def library_path() -> str:
    # GB_LIBRARY_PATH: tuning builds of the same sources (e.g. a different pipeline depth)
    return os.environ.get("GB_LIBRARY_PATH") or os.path.join(_HERE, "lib", "libgauss_b200.so")


def header_path() -> str:
    return os.path.join(os.path.dirname(_HERE), "include", "gauss_b200.h")


def exported_symbols() -> list[str]:
    """Names declared GB_API in include/gauss_b200.h."""
    txt = open(header_path()).read()
    return sorted(set(re.findall(r"GB_API[^;(]*?\b(gb_\w+)\s*\(", txt)))


def load_library():
    """Load the CUDA library; fail loudly if it has not been built (no fallback exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: run `python -m gauss_b200.build` (needs nvcc). "
                          "gauss_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(path)
    vp, i64, i32p, dblp, i64p = C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p
    sig = {
        "gb_params_default": (None, [C.POINTER(Params)]),
        "gb_version": (C.c_int, []),
        "gb_status_string": (C.c_char_p, [C.c_int]),
        "gb_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "gb_ctx_destroy": (None, [vp]),
        "gb_ctx_set_stream": (C.c_int, [vp, vp]),
        "gb_ctx_synchronize": (C.c_int, [vp]),
        "gb_last_error": (C.c_char_p, [vp]),
        "gb_ctx_launch_count": (i64, [vp]),
        "gb_panel_create": (C.c_int, [vp, C.c_int, i32p, i64, C.POINTER(vp)]),
        "gb_panel_create_fmt": (C.c_int, [vp, C.c_int, i32p, i64, C.c_int, C.POINTER(vp)]),
        "gb_panel_format_of": (C.c_int, [vp]),
        "gb_panel_destroy": (None, [vp]),
        "gb_panel_clear": (C.c_int, [vp]),
        "gb_panel_num_rows": (i64, [vp]),
        "gb_panel_num_samples": (i64, [vp]),
        "gb_panel_append_strings": (C.c_int, [vp, i64, vp]),
        "gb_panel_append_host": (C.c_int, [vp, i64, vp, i64, C.c_int]),
        "gb_panel_append_device": (C.c_int, [vp, i64, vp, i64, C.c_int]),
        "gb_gram_counts": (C.c_int, [vp, vp, i64, i64p, i64, i64p, vp, vp, vp]),
        "gb_window_dist": (C.c_int, [vp, vp, i64, i64p, i64, i64p, dblp, C.POINTER(Params), dblp, dblp]),
        "gb_window_distmix": (C.c_int, [vp, vp, i64, i64p, i64, i64p, dblp, dblp, C.POINTER(Params), dblp, dblp]),
        "gb_window_ld": (C.c_int, [vp, vp, i64, i64p, dblp, dblp]),
        "gb_window_cor": (C.c_int, [vp, vp, i64, i64p, i64, i64p, dblp, C.POINTER(Params), dblp, dblp]),
        "gb_genes_ld": (C.c_int, [vp, vp, i64, i64p, i64p, dblp, C.c_double, dblp]),
        "gb_genes_jepeg": (C.c_int, [vp, vp, i64, i64p, i64p, dblp, dblp, dblp, dblp, C.c_double, C.c_double, C.c_double,
                                     C.c_int, dblp]),
        "gb_zmix_pair_cor": (C.c_int, [vp, vp, i64, i64p, dblp, dblp]),
        "gb_window_qcat": (C.c_int, [vp, vp, i64, i64p, dblp, i64, i64, i64, i64p, dblp, C.POINTER(Params), C.c_double,
                                     C.POINTER(C.c_int), dblp, dblp, dblp, dblp]),
        "gb_batch_create": (C.c_int, [vp, vp, i64, i64p, i64p, i64p, i64p, dblp, dblp, C.POINTER(Params),
                                      C.POINTER(vp)]),
        "gb_batch_create_ld": (C.c_int, [vp, vp, i64, i64p, i64p, dblp, C.c_double, C.POINTER(vp)]),
        "gb_batch_destroy": (None, [vp]),
        "gb_batch_run": (C.c_int, [vp]),
        "gb_batch_run_stage": (C.c_int, [vp, C.c_int]),
        "gb_batch_fetch": (C.c_int, [vp, dblp, dblp, vp]),
        "gb_batch_work": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "gb_pack2_row_bytes": (i64, [C.c_int, i32p]),
        "gb_pack2_rows_host": (C.c_int, [C.c_int, i32p, i64, vp, i64, C.c_int, vp, i64]),
        "gb_panel_append_pack2_host": (C.c_int, [vp, i64, vp, i64]),
        "gb_chrom_run_pack2": (C.c_int, [vp, vp, i64, vp, i64, i64, i64p, i64p, i64p, i64p, dblp, dblp,
                                         C.POINTER(Params), C.c_int, dblp, dblp, vp]),
        "gb_pack5_row_bytes": (i64, [C.c_int, i32p]),
        "gb_pack5_rows_host": (C.c_int, [C.c_int, i32p, i64, vp, i64, C.c_int, vp, i64]),
        "gb_panel_append_pack5_host": (C.c_int, [vp, i64, vp, i64]),
        "gb_panel_append_pack5_device": (C.c_int, [vp, i64, vp, i64]),
        "gb_chrom_run_pack5": (C.c_int, [vp, vp, i64, vp, i64, i64, i64p, i64p, i64p, i64p, dblp, dblp,
                                         C.POINTER(Params), C.c_int, dblp, dblp, vp]),
        "gb_pipe_create": (C.c_int, [vp, C.c_int, i32p, i64, C.c_int, C.c_int, C.POINTER(vp)]),
        "gb_pipe_destroy": (None, [vp]),
        "gb_pipe_submit": (C.c_int, [vp, i64, vp, i64, vp, i64, C.c_int, dblp, dblp, C.POINTER(Params), dblp, dblp,
                                     C.POINTER(C.c_int64)]),
        "gb_pipe_wait": (C.c_int, [vp, i64, C.POINTER(C.c_int)]),
        "gb_genome_create": (C.c_int, [C.c_int, i32p, C.c_int, i32p, dblp, C.POINTER(Params), C.POINTER(vp)]),
        "gb_genome_destroy": (None, [vp]),
        "gb_genome_last_error": (C.c_char_p, [vp]),
        "gb_genome_add_chromosome": (C.c_int, [vp, i64, vp, i64, i64, i64p, i64p, i64p, i64p, dblp, i64p]),
        "gb_genome_plan": (C.c_int, [vp, C.c_int, C.c_int]),
        "gb_genome_num_chromosomes": (C.c_int, [vp]),
        "gb_genome_shard_info": (C.c_int, [vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                           C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "gb_genome_upload": (C.c_int, [vp, C.c_int]),
        "gb_genome_download_rows": (C.c_int, [vp, C.c_int, C.c_int, i64, i64, vp, i64]),
        "gb_genome_resident_ranges": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, i64p, i64p, C.POINTER(C.c_int)]),
        "gb_genome_set_host_rows": (C.c_int, [vp, C.c_int, i64, i64, vp, i64]),
        "gb_genome_fill_synthetic": (C.c_int, [vp, C.c_uint64, C.c_int]),
        "gb_genome_submit": (C.c_int, [vp, vp, vp, vp]),
        "gb_genome_wait": (C.c_int, [vp, dblp, dblp]),
        "gb_genome_run": (C.c_int, [vp, vp, vp, vp, dblp]),
        "gb_genome_launch_count": (i64, [vp]),
        "gb_partition_windows": (C.c_int, [i64, i64p, i64p, i64, C.POINTER(Params), C.c_int, i64p, dblp]),
        "gb_synth_pack5_rows": (C.c_int, [vp, C.c_uint64, C.c_int, i64, i64p, i64, C.c_int, i32p, vp, i64, C.c_int]),
        "gb_packfile_convert": (C.c_int, [C.c_char_p, C.c_int, i32p, C.c_char_p, C.c_int, C.POINTER(C.c_int64),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_char_p, C.c_int]),
        "gb_probe_peak": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_double)]),
        "gb_run_qcat_strings": (C.c_int, [vp, i64, vp, vp, dblp, vp, C.c_int, vp, dblp, C.c_longlong, C.c_longlong,
                                          C.POINTER(Params), C.c_double, dblp, dblp, dblp, C.POINTER(C.c_int),
                                          C.POINTER(C.c_int)]),
        "gb_run_window_strings": (C.c_int, [vp, i64, vp, vp, dblp, dblp, vp, C.c_int, vp, dblp, C.c_longlong,
                                            C.c_longlong, C.POINTER(Params), C.POINTER(C.c_int),
                                            C.POINTER(C.c_int)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def pack2_row_bytes(pop_sizes) -> int:
    """Bytes per SNP row of the 2-bit host format (gb_pack2_row_bytes)."""
    ps = np.ascontiguousarray(pop_sizes, np.int32)
    return int(load_library().gb_pack2_row_bytes(len(ps), _ptr(ps)))


def pack5_row_bytes(pop_sizes) -> int:
    """Bytes per SNP row of the ternary host format, five dosages per byte (gb_pack5_row_bytes)."""
    ps = np.ascontiguousarray(pop_sizes, np.int32)
    return int(load_library().gb_pack5_row_bytes(len(ps), _ptr(ps)))


def pack5_rows_host(pop_sizes, rows: np.ndarray, is_ascii: bool | None = None, out: np.ndarray | None = None):
    """HOST-side packer of the ternary format (gb_pack5_rows_host); same contract as pack2_rows_host."""
    return pack2_rows_host(pop_sizes, rows, is_ascii, out, _fmt=5)


def pack2_rows_host(pop_sizes, rows: np.ndarray, is_ascii: bool | None = None, out: np.ndarray | None = None,
                    _fmt: int = 2):
    """HOST-side packer: [n, n_samples] int8 dosages / uint8 chars -> [n, pack2_row_bytes] uint8 (CPU threads)."""
    lib = load_library()
    ps = np.ascontiguousarray(pop_sizes, np.int32)
    assert rows.ndim == 2 and rows.itemsize == 1
    if rows.size and rows.strides[1] != 1:
        rows = np.ascontiguousarray(rows)
    if is_ascii is None:
        is_ascii = rows.dtype == np.uint8
        if is_ascii and rows.size and int(rows.max()) < 32:
            raise ValueError("uint8 rows holding values below 32 look like numeric dosages, not characters: pass is_ascii explicitly")
    rb = pack5_row_bytes(ps) if _fmt == 5 else pack2_row_bytes(ps)
    if out is None:
        out = np.empty((rows.shape[0], rb), np.uint8)
    assert out.shape == (rows.shape[0], rb) and (out.size == 0 or out.strides[1] == 1)
    if rows.shape[0] == 0:
        return out
    fn = lib.gb_pack5_rows_host if _fmt == 5 else lib.gb_pack2_rows_host
    rc = fn(len(ps), _ptr(ps), rows.shape[0], rows.ctypes.data, rows.strides[0],
                                int(bool(is_ascii)), out.ctypes.data, out.strides[0])
    if rc != GB_OK:
        raise GaussB200Error(rc, lib.gb_status_string(rc).decode())
    return out


class Context:
    """gb_ctx: one per GPU."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.gb_ctx_create(device, C.byref(h))
        if rc != GB_OK:
            raise GaussB200Error(rc, self.lib.gb_last_error(None).decode() or
                                 self.lib.gb_status_string(rc).decode())
        self.h = h
        self.device = device
        self._children = weakref.WeakSet()   # panels / batches / pipes: destroyed before the context they point into

    def check(self, rc: int, allow=()):
        if rc != GB_OK and rc not in allow:
            raise GaussB200Error(rc, self.lib.gb_last_error(self.h).decode() or
                                 self.lib.gb_status_string(rc).decode())
        return rc

    def set_stream(self, cuda_stream: int | None):
        self.check(self.lib.gb_ctx_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def synchronize(self):
        self.check(self.lib.gb_ctx_synchronize(self.h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.gb_ctx_launch_count(self.h))

    def probe_peak(self, which: str, reps: int = 3) -> float:
        """gb_probe_peak: "i8" / "mxf4" (TOP/s), "fp64" (TFLOP/s), "copy" (GB/s) measured on this GPU now."""
        v = C.c_double()
        self.check(self.lib.gb_probe_peak(self.h, {"i8": 0, "mxf4": 1, "fp64": 2, "copy": 3}[which], reps, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "h", None):
            for child in list(getattr(self, "_children", ())):
                child.close()
            self.lib.gb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host-side mirror of run_dist / run_distmix (dist.cpp:129-227, distmix.cpp:138-253) ----
    def run_window_strings(self, type_, bp, z, info, pop_strings, pop_sizes, pop_wgt, start_bp, end_bp,
                           params: Params | None = None):
        """pop_strings: list (per SNP) of list (per population) of bytes.  Updates z/info copies."""
        type_ = np.ascontiguousarray(type_, np.int32)
        bp = np.ascontiguousarray(bp, np.int64)
        z = np.array(z, np.float64, copy=True)
        info = np.array(info, np.float64, copy=True)
        pop_sizes = np.ascontiguousarray(pop_sizes, np.int32)
        n, P = len(type_), len(pop_sizes)
        arr = (C.c_char_p * (n * P))()
        keep = []
        for i in range(n):
            for k in range(P):
                s = pop_strings[i][k] if pop_strings[i] is not None else None
                keep.append(s)
                arr[i * P + k] = s
        w = None if pop_wgt is None else _f64(pop_wgt)
        nt, nu = C.c_int(0), C.c_int(0)
        rc = self.lib.gb_run_window_strings(self.h, n, _ptr(type_), _ptr(bp), _ptr(z), _ptr(info),
                                            C.cast(arr, C.c_void_p), P, _ptr(pop_sizes), _ptr(w), int(start_bp),
                                            int(end_bp), C.byref(params) if params else None, C.byref(nt),
                                            C.byref(nu))
        return dict(rc=rc, z=z, info=info, n_t=nt.value, n_u=nu.value)


def _string_table(pop_strings, n, P):
    arr = (C.c_char_p * (n * P))()
    keep = []
    for i in range(n):
        for k in range(P):
            s = pop_strings[i][k] if pop_strings[i] is not None else None
            keep.append(s)
            arr[i * P + k] = s
    return arr, keep


def run_qcat_strings(ctx: "Context", type_, bp, z, pop_strings, pop_sizes, pop_wgt, start_bp, end_bp,
                     params: Params | None = None, eig_cutoff: float = 0.01):
    """gb_run_qcat_strings: host mirror of run_qcat / run_qcatmix.  -> dict(rc, m, t, chisq) (NaN = untested)."""
    type_ = np.ascontiguousarray(type_, np.int32)
    bp = np.ascontiguousarray(bp, np.int64)
    z = _f64(z)
    pop_sizes = np.ascontiguousarray(pop_sizes, np.int32)
    n, P = len(type_), len(pop_sizes)
    arr, keep = _string_table(pop_strings, n, P)
    w = None if pop_wgt is None else _f64(pop_wgt)
    qm, qt, qc = (np.full(n, np.nan) for _ in range(3))
    nt, nu = C.c_int(0), C.c_int(0)
    rc = ctx.lib.gb_run_qcat_strings(ctx.h, n, _ptr(type_), _ptr(bp), _ptr(z), C.cast(arr, C.c_void_p), P,
                                     _ptr(pop_sizes), _ptr(w), int(start_bp), int(end_bp),
                                     C.byref(params) if params else None, float(eig_cutoff), _ptr(qm), _ptr(qt),
                                     _ptr(qc), C.byref(nt), C.byref(nu))
    return dict(rc=rc, m=qm, t=qt, chisq=qc, n_t=nt.value, n_u=nu.value)


class Panel:
    """gb_panel: HBM-resident packed panel (int8 rows, per-population blocks padded to 32)."""

    def __init__(self, ctx: Context, pop_sizes, capacity_rows: int, fmt: str | None = None):
        """fmt: None (library default: E2M1 nibbles), "int8" or "e2m1" (enum gb_panel_format)."""
        self.ctx = ctx
        self.pop_sizes = np.ascontiguousarray(pop_sizes, np.int32)
        h = C.c_void_p()
        if fmt is None:
            ctx.check(ctx.lib.gb_panel_create(ctx.h, len(self.pop_sizes), _ptr(self.pop_sizes),
                                              int(capacity_rows), C.byref(h)))
        else:
            ctx.check(ctx.lib.gb_panel_create_fmt(ctx.h, len(self.pop_sizes), _ptr(self.pop_sizes),
                                                  int(capacity_rows), PANEL_FORMATS[fmt], C.byref(h)))
        self.h = h
        self.ctx._children.add(self)

    @property
    def format(self) -> str:
        return {v: k for k, v in PANEL_FORMATS.items()}[int(self.ctx.lib.gb_panel_format_of(self.h))]

    @property
    def n_rows(self) -> int:
        return int(self.ctx.lib.gb_panel_num_rows(self.h))

    @property
    def n_samples(self) -> int:
        return int(self.ctx.lib.gb_panel_num_samples(self.h))

    def clear(self):
        self.ctx.check(self.ctx.lib.gb_panel_clear(self.h))

    def append_host(self, rows: np.ndarray, is_ascii: bool | None = None):
        """rows: [n, n_samples] int8 dosages or uint8 ASCII chars (HOST numpy array)."""
        rows = np.ascontiguousarray(rows)
        if is_ascii is None:
            is_ascii = rows.dtype == np.uint8
            if is_ascii and rows.size and int(rows.max()) < 32:
                raise ValueError("uint8 rows holding values below 32 look like numeric dosages, not characters: "
                                 "pass is_ascii explicitly")
        assert rows.ndim == 2 and rows.itemsize == 1
        self.ctx.check(self.ctx.lib.gb_panel_append_host(self.h, rows.shape[0], rows.ctypes.data,
                                                         rows.strides[0], int(bool(is_ascii))))

    def append_pack2_host(self, rows2: np.ndarray):
        """rows2: [n, pack2_row_bytes] uint8 rows of the 2-bit host format (E2M1 panels only)."""
        assert rows2.ndim == 2 and rows2.dtype == np.uint8 and rows2.strides[1] == 1
        self.ctx.check(self.ctx.lib.gb_panel_append_pack2_host(self.h, rows2.shape[0], rows2.ctypes.data,
                                                               rows2.strides[0]))

    def append_pack5_host(self, rows5: np.ndarray):
        """rows5: [n, pack5_row_bytes] uint8 rows of the ternary host format (E2M1 panels only)."""
        assert rows5.ndim == 2 and rows5.dtype == np.uint8 and rows5.strides[1] == 1
        self.ctx.check(self.ctx.lib.gb_panel_append_pack5_host(self.h, rows5.shape[0], rows5.ctypes.data,
                                                               rows5.strides[0]))

    def append_pack5_device_ptr(self, ptr: int, n_rows: int, row_stride: int):
        self.ctx.check(self.ctx.lib.gb_panel_append_pack5_device(self.h, int(n_rows), C.c_void_p(ptr), int(row_stride)))

    def chrom_run_pack5(self, rows5_ptr: int, n_rows: int, row_stride: int, t_off, rows_t, u_off, rows_u, z_t,
                        pop_wgt=None, params: Params | None = None, n_groups: int = 4, z=None, info=None):
        """gb_chrom_run_pack5: as chrom_run_pack2 on ternary host rows."""
        return self.chrom_run_pack2(rows5_ptr, n_rows, row_stride, t_off, rows_t, u_off, rows_u, z_t, pop_wgt, params,
                                    n_groups, z, info, _fmt=5)

    def chrom_run_pack2(self, rows2_ptr: int, n_rows: int, row_stride: int, t_off, rows_t, u_off, rows_u, z_t,
                        pop_wgt=None, params: Params | None = None, n_groups: int = 4, z=None, info=None, _fmt: int = 2):
        """gb_chrom_run_pack2: one chromosome from pack2 HOST rows to HOST results (clears this panel)."""
        t_off, u_off, rt, ru, zt = _i64(t_off), _i64(u_off), _i64(rows_t), _i64(rows_u), _f64(z_t)
        w = None if pop_wgt is None else _f64(pop_wgt)
        nw = len(t_off) - 1
        n = int(u_off[-1])
        z = np.zeros(n) if z is None else z
        info = np.zeros(n) if info is None else info
        status = np.zeros(nw, np.int32)
        fn = self.ctx.lib.gb_chrom_run_pack5 if _fmt == 5 else self.ctx.lib.gb_chrom_run_pack2
        self.ctx.check(fn(
            self.ctx.h, self.h, int(n_rows), C.c_void_p(rows2_ptr), int(row_stride), nw, _ptr(t_off), _ptr(rt),
            _ptr(u_off), _ptr(ru), _ptr(zt), _ptr(w), C.byref(params) if params else None, int(n_groups), _ptr(z),
            _ptr(info), _ptr(status)))
        return z, info, status

    def append_host_ptr(self, ptr: int, n_rows: int, row_stride: int, is_ascii: bool):
        self.ctx.check(self.ctx.lib.gb_panel_append_host(self.h, n_rows, C.c_void_p(ptr), row_stride,
                                                         int(bool(is_ascii))))

    def append_device_ptr(self, ptr: int, n_rows: int, row_stride: int, is_ascii: bool):
        self.ctx.check(self.ctx.lib.gb_panel_append_device(self.h, n_rows, C.c_void_p(ptr), row_stride,
                                                           int(bool(is_ascii))))

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):     # a closed context has already taken its children with it
                self.ctx.lib.gb_panel_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- parity surface ------------------------------------------------------------------
    def gram_counts(self, rows_a, rows_b):
        ra, rb = _i64(rows_a), _i64(rows_b)
        P = len(self.pop_sizes)
        sxy = np.zeros((P, len(ra), len(rb)), np.int32)
        sx = np.zeros((P, len(ra)), np.int32)
        sxx = np.zeros((P, len(ra)), np.int32)
        self.ctx.check(self.ctx.lib.gb_gram_counts(self.ctx.h, self.h, len(ra), _ptr(ra), len(rb), _ptr(rb),
                                                   _ptr(sxy), _ptr(sx), _ptr(sxx)))
        return sxy, sx, sxx

    def window_cor(self, rows_t, rows_u, pop_wgt=None, params: Params | None = None):
        rt, ru = _i64(rows_t), _i64(rows_u)
        w = None if pop_wgt is None else _f64(pop_wgt)
        B11 = np.zeros((len(rt), len(rt)))
        B21 = np.zeros((len(ru), len(rt)))
        self.ctx.check(self.ctx.lib.gb_window_cor(self.ctx.h, self.h, len(rt), _ptr(rt), len(ru), _ptr(ru),
                                                  _ptr(w), C.byref(params) if params else None, _ptr(B11),
                                                  _ptr(B21)))
        return B11, B21

    # ---- one window ---------------------------------------------------------------------------
    def window_dist(self, rows_t, rows_u, z_t, params: Params | None = None, allow=()):
        return self._impute(rows_t, rows_u, z_t, None, params, allow)

    def window_distmix(self, rows_t, rows_u, z_t, pop_wgt, params: Params | None = None, allow=()):
        return self._impute(rows_t, rows_u, z_t, _f64(pop_wgt), params, allow)

    def _impute(self, rows_t, rows_u, z_t, w, params, allow):
        rt, ru, zt = _i64(rows_t), _i64(rows_u), _f64(z_t)
        z = np.zeros(len(ru))
        info = np.zeros(len(ru))
        lib, pp = self.ctx.lib, (C.byref(params) if params else None)
        if w is None:
            rc = lib.gb_window_dist(self.ctx.h, self.h, len(rt), _ptr(rt), len(ru), _ptr(ru), _ptr(zt), pp,
                                    _ptr(z), _ptr(info))
        else:
            rc = lib.gb_window_distmix(self.ctx.h, self.h, len(rt), _ptr(rt), len(ru), _ptr(ru), _ptr(zt),
                                       _ptr(w), pp, _ptr(z), _ptr(info))
        self.ctx.check(rc, allow)
        return z, info, rc

    def genes_ld(self, g_off, rows, pop_wgt=None, diag: float = 1.1):
        """gb_genes_ld: list of per-gene correlation matrices (CorG of jepeg / jepegmix)."""
        go, r = _i64(g_off), _i64(rows)
        w = None if pop_wgt is None else _f64(pop_wgt)
        sizes = np.diff(go)
        out = np.zeros(int((sizes * sizes).sum()))
        self.ctx.check(self.ctx.lib.gb_genes_ld(self.ctx.h, self.h, len(go) - 1, _ptr(go), _ptr(r), _ptr(w), float(diag),
                                                _ptr(out)))
        offs = np.concatenate([[0], np.cumsum(sizes * sizes)])
        return [out[offs[g]:offs[g + 1]].reshape(sizes[g], sizes[g]) for g in range(len(sizes))]

    def genes_jepeg(self, g_off, rows, z, info, categ_wgt, pop_wgt=None, lam: float = 0.1, min_abs_eig: float = 1e-5,
                    categ_cor_cutoff: float = 0.8, denorm_norm_w: int = 3):
        """gb_genes_jepeg: [n_genes, 16] statistics of jepeg() (pop_wgt None) / jepegmix() for every gene."""
        go, r = _i64(g_off), _i64(rows)
        w = None if pop_wgt is None else _f64(pop_wgt)
        zz, ii, cw = _f64(z), _f64(info), _f64(categ_wgt)
        assert cw.shape == (len(r), 6) and len(zz) == len(r) == len(ii)
        out = np.zeros((len(go) - 1, 16))
        self.ctx.check(self.ctx.lib.gb_genes_jepeg(self.ctx.h, self.h, len(go) - 1, _ptr(go), _ptr(r), _ptr(w), _ptr(zz), _ptr(ii),
                                                   _ptr(cw), float(lam), float(min_abs_eig), float(categ_cor_cutoff),
                                                   int(denorm_norm_w), _ptr(out)))
        return out

    def zmix_pair_cor(self, rows, z):
        """gb_zmix_pair_cor: prep_zmix5's pair matrix [n(n-1)/2, 1 + P] (column 0 = z_i z_j)."""
        r, zz = _i64(rows), _f64(z)
        n, P = len(r), len(self.pop_sizes)
        out = np.zeros((1 + P, n * (n - 1) // 2))
        self.ctx.check(self.ctx.lib.gb_zmix_pair_cor(self.ctx.h, self.h, n, _ptr(r), _ptr(zz), _ptr(out)))
        return out.T

    def window_qcat(self, rows_t, z_t, core_first, n_core, rows_u, pop_wgt=None, params: Params | None = None,
                    eig_cutoff: float = 0.01, allow=()):
        """gb_window_qcat: qcat (pop_wgt None) / qcatmix.  -> dict(rc, num_eig, t_m, chisq_m, t_u, chisq_u)."""
        rt, ru, zt = _i64(rows_t), _i64(rows_u), _f64(z_t)
        w = None if pop_wgt is None else _f64(pop_wgt)
        t_m, c_m = np.zeros(n_core), np.zeros(n_core)
        t_u, c_u = np.zeros(len(ru)), np.zeros(len(ru))
        ne = C.c_int(0)
        rc = self.ctx.lib.gb_window_qcat(self.ctx.h, self.h, len(rt), _ptr(rt), _ptr(zt), int(core_first), int(n_core),
                                         len(ru), _ptr(ru), _ptr(w), C.byref(params) if params else None,
                                         float(eig_cutoff), C.byref(ne), _ptr(t_m), _ptr(c_m), _ptr(t_u), _ptr(c_u))
        self.ctx.check(rc, allow)
        return dict(rc=rc, num_eig=ne.value, t_m=t_m, chisq_m=c_m, t_u=t_u, chisq_u=c_u)

    def window_ld(self, rows, pop_wgt, allow=()):
        r, w = _i64(rows), _f64(pop_wgt)
        cm = np.zeros((len(r), len(r)))
        rc = self.ctx.lib.gb_window_ld(self.ctx.h, self.h, len(r), _ptr(r), _ptr(w), _ptr(cm))
        self.ctx.check(rc, allow)
        return cm, rc


class Batch:
    """gb_batch: many windows on one resident panel."""

    def __init__(self, panel: Panel, t_off, rows_t, u_off, rows_u, z_t, pop_wgt=None,
                 params: Params | None = None, ld_diag: float | None = None):
        """ld_diag given -> computeLD blocks (gb_batch_create_ld): u_off / rows_u / z_t are ignored."""
        self.panel, self.ctx = panel, panel.ctx
        self.t_off = _i64(t_off)
        self.rows_t = _i64(rows_t)
        self.w = None if pop_wgt is None else _f64(pop_wgt)
        self.n_windows = len(self.t_off) - 1
        h = C.c_void_p()
        if ld_diag is not None:
            self.u_off = np.zeros(self.n_windows + 1, np.int64)
            self.ctx.check(self.ctx.lib.gb_batch_create_ld(self.ctx.h, panel.h, self.n_windows, _ptr(self.t_off),
                                                           _ptr(self.rows_t), _ptr(self.w), float(ld_diag), C.byref(h)))
            self.h = h
            self.ctx._children.add(self)
            return
        self.u_off = _i64(u_off)
        self.rows_u, self.z_t = _i64(rows_u), _f64(z_t)
        self.ctx.check(self.ctx.lib.gb_batch_create(
            self.ctx.h, panel.h, self.n_windows, _ptr(self.t_off), _ptr(self.rows_t), _ptr(self.u_off),
            _ptr(self.rows_u), _ptr(self.z_t), _ptr(self.w), C.byref(params) if params else None, C.byref(h)))
        self.h = h
        self.ctx._children.add(self)

    def run(self):
        self.ctx.check(self.ctx.lib.gb_batch_run(self.h))

    def run_stage(self, stage: int):
        self.ctx.check(self.ctx.lib.gb_batch_run_stage(self.h, stage))

    def fetch(self, z=None, info=None):
        n = int(self.u_off[-1])
        z = np.zeros(n) if z is None else z
        info = np.zeros(n) if info is None else info
        status = np.zeros(self.n_windows, np.int32)
        self.ctx.check(self.ctx.lib.gb_batch_fetch(self.h, _ptr(z), _ptr(info), _ptr(status)))
        return z, info, status

    def work(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.ctx.check(self.ctx.lib.gb_batch_work(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return dict(gram_ops=a.value, solve_flops=b.value, panel_bytes=c.value)

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):     # a closed context has already taken its children with it
                self.ctx.lib.gb_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Pipe:
    """gb_pipe: asynchronous per-window pipeline on host buffers (H2D of window w+1 overlaps window w)."""

    def __init__(self, ctx: Context, pop_sizes, max_rows_per_window: int, depth: int = 2, fmt: str | None = None):
        self.ctx = ctx
        self.pop_sizes = np.ascontiguousarray(pop_sizes, np.int32)
        h = C.c_void_p()
        ctx.check(ctx.lib.gb_pipe_create(ctx.h, len(self.pop_sizes), _ptr(self.pop_sizes), int(max_rows_per_window),
                                         int(depth), -1 if fmt is None else PANEL_FORMATS[fmt], C.byref(h)))
        self.h = h
        self.ctx._children.add(self)
        self._keep = {}

    def submit_ptr(self, ptr_t: int, n_t: int, ptr_u: int, n_u: int, row_stride: int, is_ascii: bool, z_t, pop_wgt,
                   z_u: np.ndarray, info_u: np.ndarray, params: Params | None = None) -> int:
        """Host pointers to n_t measured / n_u unmeasured rows; z_u / info_u are written by the time
        wait(ticket) returns.  Buffers must stay alive until then (they are kept referenced here)."""
        zt = _f64(z_t)
        w = None if pop_wgt is None else _f64(pop_wgt)
        t = C.c_int64(-1)
        self.ctx.check(self.ctx.lib.gb_pipe_submit(self.h, n_t, C.c_void_p(ptr_t), n_u, C.c_void_p(ptr_u), row_stride,
                                                   int(bool(is_ascii)), _ptr(zt), _ptr(w),
                                                   C.byref(params) if params else None, _ptr(z_u), _ptr(info_u),
                                                   C.byref(t)))
        self._keep[t.value] = (zt, w, z_u, info_u)
        return t.value

    def submit(self, rows_t: np.ndarray, rows_u: np.ndarray, z_t, pop_wgt, z_u=None, info_u=None,
               params: Params | None = None):
        rows_t, rows_u = np.ascontiguousarray(rows_t), np.ascontiguousarray(rows_u)
        assert rows_t.itemsize == 1 and rows_u.itemsize == 1
        z_u = np.zeros(len(rows_u)) if z_u is None else z_u
        info_u = np.zeros(len(rows_u)) if info_u is None else info_u
        t = self.submit_ptr(rows_t.ctypes.data, len(rows_t), rows_u.ctypes.data, len(rows_u), rows_t.shape[1],
                            rows_t.dtype == np.uint8, z_t, pop_wgt, z_u, info_u, params)
        self._keep[t] = self._keep[t] + (rows_t, rows_u)
        return t, z_u, info_u

    def wait(self, ticket: int) -> int:
        st = C.c_int(0)
        self.ctx.check(self.ctx.lib.gb_pipe_wait(self.h, ticket, C.byref(st)))
        self._keep.pop(ticket, None)
        return st.value

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):     # a closed context has already taken its children with it
                self.ctx.lib.gb_pipe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def partition_windows(n_t, n_u, n_samples: int, n_parts: int, params: Params | None = None):
    """gb_partition_windows: the cost-balanced contiguous cuts gb_genome_plan uses (pure host code).
    -> (cuts [n_parts + 1], cost per window)."""
    lib = load_library()
    nt, nu = _i64(n_t), _i64(n_u)
    cuts = np.zeros(n_parts + 1, np.int64)
    cost = np.zeros(len(nt))
    rc = lib.gb_partition_windows(len(nt), _ptr(nt), _ptr(nu), int(n_samples), C.byref(params) if params else None,
                                  int(n_parts), _ptr(cuts), _ptr(cost))
    if rc != GB_OK:
        raise GaussB200Error(rc, lib.gb_status_string(rc).decode())
    return cuts, cost


def unpack5_rows(rows5: np.ndarray, pop_sizes) -> np.ndarray:
    """Ternary host rows -> [n, sum(pop_sizes)] int8 dosages (numpy; formatting only: the inverse of
    pack5_rows_host, for callers that need the dosages of a packed or synthetic panel back)."""
    ps = np.ascontiguousarray(pop_sizes, np.int64)
    rows5 = np.ascontiguousarray(rows5, np.uint8)
    n = rows5.shape[0]
    out = np.empty((n, int(ps.sum())), np.int8)
    pw = 3 ** np.arange(5)
    boff = 0
    col = 0
    for m in ps:
        nb = (int(m) + 4) // 5
        blk = rows5[:, boff:boff + nb].astype(np.int64)
        dig = ((blk[:, :, None] // pw[None, None, :]) % 3).reshape(n, nb * 5)[:, :int(m)]
        out[:, col:col + int(m)] = dig.astype(np.int8)
        col += int(m)
        boff += (nb + 3) // 4 * 4
    return out


def synth_pack5_rows(ctx: "Context", seed: int, chrom: int, pop_sizes, n_rows: int, sites=None, first_site: int = 0,
                     out: np.ndarray | None = None):
    """gb_synth_pack5_rows into a HOST array [n_rows, pack5_row_bytes] (the device generator of the bench panels)."""
    ps = np.ascontiguousarray(pop_sizes, np.int32)
    rb = pack5_row_bytes(ps)
    if out is None:
        out = np.zeros((n_rows, rb), np.uint8)
    assert out.shape == (n_rows, rb) and out.dtype == np.uint8 and (out.size == 0 or out.strides == (rb, 1))
    st = None if sites is None else _i64(sites)
    ctx.check(ctx.lib.gb_synth_pack5_rows(ctx.h, int(seed), int(chrom), int(n_rows), _ptr(st), int(first_site), len(ps),
                                          _ptr(ps), out.ctypes.data, rb, 0))
    return out


def synth_pack5_rows_device(ctx: "Context", seed: int, chrom: int, pop_sizes, n_rows: int, dev_ptr: int, row_stride: int,
                            sites=None, first_site: int = 0):
    """gb_synth_pack5_rows straight into DEVICE memory (dev_ptr: n_rows * row_stride bytes on ctx's GPU)."""
    ps = np.ascontiguousarray(pop_sizes, np.int32)
    st = None if sites is None else _i64(sites)
    ctx.check(ctx.lib.gb_synth_pack5_rows(ctx.h, int(seed), int(chrom), int(n_rows), _ptr(st), int(first_site), len(ps),
                                          _ptr(ps), C.c_void_p(dev_ptr), int(row_stride), 1))


class Genome:
    """gb_genome: genome-wide dist()/distmix() from one process on n_gpus GPUs (one host thread per GPU inside the
    library).  chromosomes are added as dicts / keyword arguments, then plan() -> upload() or fill_synthetic() -> run()."""

    def __init__(self, n_gpus: int, pop_sizes, pop_wgt=None, params: Params | None = None, devices=None):
        self.lib = load_library()
        self.pop_sizes = np.ascontiguousarray(pop_sizes, np.int32)
        self.w = None if pop_wgt is None else _f64(pop_wgt)
        dv = None if devices is None else np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        rc = self.lib.gb_genome_create(int(n_gpus), _ptr(dv), len(self.pop_sizes), _ptr(self.pop_sizes), _ptr(self.w),
                                       C.byref(params) if params else None, C.byref(h))
        if rc != GB_OK:
            raise GaussB200Error(rc, self.lib.gb_last_error(None).decode() or self.lib.gb_status_string(rc).decode())
        self.h = h
        self.n_gpus = int(n_gpus)
        self.chroms = []        # (n_u_total, n_windows) per chromosome
        self._keep = []

    def check(self, rc: int, allow=()):
        if rc != GB_OK and rc not in allow:
            raise GaussB200Error(rc, self.lib.gb_genome_last_error(self.h).decode() or
                                 self.lib.gb_status_string(rc).decode())
        return rc

    def add_chromosome(self, n_rows: int, t_off, rows_t, u_off, rows_u, z_t, rows5_ptr: int | None = None,
                       row_stride: int = 0, sites=None):
        t_off, u_off, rt, ru, zt = _i64(t_off), _i64(u_off), _i64(rows_t), _i64(rows_u), _f64(z_t)
        st = None if sites is None else _i64(sites)
        self.check(self.lib.gb_genome_add_chromosome(
            self.h, int(n_rows), C.c_void_p(rows5_ptr) if rows5_ptr else None, int(row_stride), len(t_off) - 1,
            _ptr(t_off), _ptr(rt), _ptr(u_off), _ptr(ru), _ptr(zt), _ptr(st)))
        self.chroms.append((int(u_off[-1]), len(t_off) - 1))
        return len(self.chroms) - 1

    def plan(self, n_parts: int | None = None, first_part: int = 0):
        self.check(self.lib.gb_genome_plan(self.h, int(n_parts or self.n_gpus), int(first_part)))

    def shard_info(self, gpu: int) -> dict:
        a, b, c, d, e = (C.c_int64() for _ in range(5))
        f = C.c_int()
        g, hh = C.c_double(), C.c_double()
        self.check(self.lib.gb_genome_shard_info(self.h, gpu, C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e),
                                                 C.byref(f), C.byref(g), C.byref(hh)))
        return dict(first_window=a.value, n_windows=b.value, resident_rows=c.value, n_batches=d.value,
                    n_imputed=e.value, e2m1_resident=f.value == 100, expanded_pct=f.value, gram_ops=g.value,
                    solve_flops=hh.value)

    def upload(self, wait: bool = True):
        self.check(self.lib.gb_genome_upload(self.h, int(bool(wait))))

    def resident_ranges(self, gpu: int, chrom: int):
        """[(lo, hi)] panel rows of `chrom` GPU `gpu` keeps resident (what a feeder has to supply)."""
        lo, hi = np.zeros(64, np.int64), np.zeros(64, np.int64)
        n = C.c_int()
        self.check(self.lib.gb_genome_resident_ranges(self.h, gpu, chrom, 64, _ptr(lo), _ptr(hi), C.byref(n)))
        return [(int(lo[i]), int(hi[i])) for i in range(min(n.value, 64))]

    def download_rows(self, gpu: int, chrom: int, row_lo: int, n_rows: int, ptr: int, row_stride: int) -> int:
        """Resident ternary rows back to host memory; returns the status (8 = this GPU keeps its rows expanded)."""
        return self.check(self.lib.gb_genome_download_rows(self.h, gpu, chrom, int(row_lo), int(n_rows), C.c_void_p(ptr),
                                                           int(row_stride)), allow=(GB_ERR_UNSUPPORTED,))

    def set_host_rows(self, chrom: int, row_lo: int, n_rows: int, ptr: int, row_stride: int):
        self.check(self.lib.gb_genome_set_host_rows(self.h, chrom, int(row_lo), int(n_rows), C.c_void_p(ptr), int(row_stride)))

    def fill_synthetic(self, seed: int, wait: bool = True):
        self.check(self.lib.gb_genome_fill_synthetic(self.h, int(seed), int(bool(wait))))

    def _out_tables(self, z, info, status):
        nc = len(self.chroms)
        z = [np.zeros(n) for n, _ in self.chroms] if z is None else z
        info = [np.zeros(n) for n, _ in self.chroms] if info is None else info
        status = [np.zeros(nw, np.int32) for _, nw in self.chroms] if status is None else status
        tabs = []
        for arrs in (z, info, status):
            t = (C.c_void_p * nc)()
            for i, a in enumerate(arrs):
                t[i] = a.ctypes.data if a is not None and a.size else None
            tabs.append(t)
        self._keep = [z, info, status, tabs]
        return z, info, status, tabs

    def submit(self, z=None, info=None, status=None):
        z, info, status, tabs = self._out_tables(z, info, status)
        self.check(self.lib.gb_genome_submit(self.h, tabs[0], tabs[1], tabs[2]))
        return z, info, status

    def wait(self):
        ms = np.zeros(self.n_gpus)
        up = np.zeros(self.n_gpus)
        self.check(self.lib.gb_genome_wait(self.h, _ptr(ms), _ptr(up)))
        return ms, up

    def run(self, z=None, info=None, status=None):
        """-> (z, info, status) lists per chromosome and the device-timed ms per GPU."""
        z, info, status = self.submit(z, info, status)
        ms, _ = self.wait()
        return z, info, status, ms

    @property
    def launch_count(self) -> int:
        return int(self.lib.gb_genome_launch_count(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.gb_genome_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
