#!/usr/bin/env python
"""bench.py -- distmix imputed SNPs/sec on a 33KG-shaped synthetic panel (BASELINE.json metric).

Workloads (`--workload`):
  genome  (default) BASELINE config 4: genome-wide distmix, 22 chromosomes, ~2,900 one-Mb windows (0.5 Mb wings),
          ~1.2 M measured / ~10 M target SNPs, 21 flagged populations / 32,147 individuals, PGC2 weights.  The SAME
          job at every N (strong scaling): the window list is cut into N contiguous cost-balanced shards
          (gb_genome_plan / gb_partition_windows), each GPU keeps its rows resident and the results are gathered on
          the host.  One process per GPU under torchrun (each rank drives a one-GPU gb_genome with its part of the
          partition); `--single-process` drives all N GPUs from ONE process through the same C-ABI (one host thread
          per GPU inside the library).
  chr22   BASELINE config 2 (the round-1 line): one chromosome-22-shaped batch, 36 windows at the bundled PGC2 positions.
  ld5000  BASELINE config 3: computeLD of a dense 5,000-SNP block (Gram + mixture epilogue only).
  dist1kg BASELINE config 1: dist() on chr22 against a 1KG-shaped panel (EUR, 503 individuals; and all 2,504).

  value     whole-job imputed SNPs/s with the panel rows resident in HBM; device-timed (CUDA events inside the library,
            around everything the step enqueues incl. the D2H of z / info), max over ranks
  e2e       the same job from pinned HOST rows (ternary packed panel): host -> device upload of every GPU's rows inside
            the timed region, batches starting as their rows land, results written to host arrays; wall clock
  roofline  the longest kernel of the step (gram_seg_kernel, tensor-bound) against the tensor-pipe rate of its instruction
            kind measured on this box at bench time (gb_probe_peak); roofline_solve (the int8-split solve GEMM against the
            kind::i8 probe), roofline_linv / roofline_chol (fp64 probe), roofline_finish / roofline_expand5 (HBM) beside it.
            GB_SOLVE=fp64 brings back the DMMA triangular solve, which is then the longest kernel and the `roofline`
  cpu_baseline / --impl reference   the reference's own CPU code (oracle/_ref) on a bounded sample, see cpu_model()
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gauss_b200 import synth  # noqa: E402

METRIC = "distmix imputed SNPs/sec"
UNIT = "SNPs/s"
SITES = os.path.join(ROOT, "tests", "golden", "pgc2_chr22_sites.npz")
SEED = 20260101


# ------------------------------------------------------------------------------------------------ workloads
def chr22_layout():
    d = np.load(SITES)
    bp_m, first = np.unique(d["bp"].astype(np.int64), return_index=True)
    z_m = d["z"][first]
    bp, type_, windows = synth.chr22_windows(bp_m)
    return bp, type_, windows, bp_m, z_m


def genome_chroms(args):
    mb = synth.HG19_MB if args.genome_mb == "all" else [int(x) for x in args.genome_mb.split(",")]
    return synth.genome_layout(mb)


def genome_config(chroms, world, mode):
    nw = sum(len(c["t_off"]) - 1 for c in chroms)
    n_m = sum(c["n_measured"] for c in chroms)
    n_u = sum(c["n_rows"] - c["n_measured"] for c in chroms)
    mb = sum(len(c["start_bp"]) for c in chroms)
    return dict(workload=f"genome-wide distmix (BASELINE config 4): {len(chroms)} chromosomes / {mb} Mb (hg19 lengths), {nw} x 1 Mb "
                         f"windows with 0.5 Mb wings, {n_m / 1e6:.2f} M measured + {n_u / 1e6:.2f} M target SNPs (synthetic sites, "
                         "log-normal measured density), 33KG-shaped panel: 21 flagged pops / 32,147 indiv (of 29 / 32,953), "
                         "PGC2_SCZ_ANC_Prop weights, lambda=0.1, PD certificate on",
                windows=nw,
                l2="inputs >> L2 (126 MB): every step streams the resident panel rows (65 GB as ternary rows at N=1)",
                parallelism=f"{world} contiguous cost-balanced window shard(s), rows resident per GPU, host gather, no collective ({mode})")


def chr22_config():
    return dict(workload="distmix chr22 (BASELINE config 2), 36 x 1 Mb windows (0.5 Mb wings), measured SNPs at PGC2_Chr22_ilmn1M_Z "
                         "positions + 3.7k synthetic unmeasured/Mb, 33KG-shaped panel: 21 flagged pops / 32,147 "
                         "indiv (of 29 / 32,953), PGC2_SCZ_ANC_Prop weights, lambda=0.1, PD certificate on",
                windows=36, l2="inputs >> L2 (126 MB): each step streams the packed panel slice (2.3 GB as E2M1 nibbles)",
                parallelism="one chromosome-shaped batch per GPU, no collective")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm_gbs=j["hbm_gbs"], bf16_tflops=j["bf16_tflops"],
                    bf16_tflops_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]), source="MEASURED_PEAKS.json")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md), sampled through
    NVML every 5 ms from a host thread; falls back to `nvidia-smi -lms` when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.proc = None
        self.path = None
        self.thread = None
        self.stop_flag = False
        self.samples = []

    def _nvml_loop(self, nv, h):
        bits = dict(hw_slowdown=nv.nvmlClocksThrottleReasonHwSlowdown,
                    hw_thermal_slowdown=nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    sw_thermal_slowdown=nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    sw_power_cap=nv.nvmlClocksThrottleReasonSwPowerCap)
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                self.samples.append((sm, [k for k, b in bits.items() if rs & b], pw))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.device
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[self.device])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=float(np.median([x[0] for x in self.samples])), sm_max_mhz=self.max_mhz,
                           samples=len(self.samples), power_w_max=float(max(x[2] for x in self.samples)),
                           source="nvml, 5 ms period, timed region only")
                out["reasons"] = sorted({r for x in self.samples for r in x[1]})
            return out
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1]))
                mx.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(names, t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm), source="nvidia-smi")
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------ CPU reference arm
def window_sizes(workload: str, args):
    """(n_t, n_u) of every window the reference would accept (dist.cpp:146) and the flagged population sizes / weights."""
    _, sizes, w = synth.flagged_33kg_pgc2()
    if workload == "genome":
        ch = genome_chroms(args)
        nt = np.concatenate([np.diff(c["t_off"]) for c in ch])
        nu = np.concatenate([np.diff(c["u_off"]) for c in ch])
    else:
        _, _, windows, _, _ = chr22_layout()
        nt = np.array([len(x["measured"]) for x in windows])
        nu = np.array([len(x["unmeasured"]) for x in windows])
    ok = (nt > 10) & (nu > 10)
    return nt[ok], nu[ok], sizes, w


def _eig_lu_seconds(n: int) -> float:
    """MakePosDef (symmetric eigensolver) + InvMat (full-pivot LU inverse) of an n x n correlation matrix: the O(n^3)
    part of a window (util.cpp:298-318).  Eigen is absent here, so this times the restated algorithms the checker
    uses (Householder + implicit QL; Gaussian elimination with complete pivoting)."""
    from oracle.oracle_py import Oracle
    o = Oracle("port")
    rng = np.random.default_rng(n)
    A = np.corrcoef(rng.standard_normal((n, 2 * n)))
    A[np.diag_indices(n)] = 1.1
    A = np.ascontiguousarray(A)
    inv = np.zeros_like(A)
    t0 = time.perf_counter()
    o.lib.go_make_pos_def(A, n, 1e-5)
    o.lib.go_inv_full_piv_lu(inv, A, n)
    return time.perf_counter() - t0


def _pair_rate_worker(arg):
    """One bounded sample window through the reference's own run_distmix (oracle/_ref): n_t measured x n_u unmeasured
    SNPs at the full 32,147 individuals.  Returns (sample-pairs, seconds, kind)."""
    seed, n_t, n_u = arg
    from oracle.oracle_py import Oracle
    kind = "reference" if Oracle.available("reference") else "port"
    orc = Oracle(kind)
    _, sizes, w = synth.flagged_33kg_pgc2()
    g = synth.make_genotypes(n_t + n_u, sizes, seed=seed)
    t = np.concatenate([np.ones(n_t, np.int32), np.zeros(n_u, np.int32)])
    zz = np.concatenate([np.random.default_rng(seed).standard_normal(n_t), np.zeros(n_u)])
    bpp = np.arange(n_t + n_u, dtype=np.int64)
    t0 = time.perf_counter()
    r = orc.run_window(t, bpp, zz, g, sizes, w, 0, 10 ** 12)
    dt = time.perf_counter() - t0
    assert r["rc"] == 0
    pairs = (n_t * (n_t - 1) / 2 + n_u * n_t + (n_t + n_u)) * float(sizes.sum())
    return pairs, dt, kind, n_t


def cpu_model(workload, args, pair_samples, eig_times):
    """CPU time of the whole step on ONE core, assembled from measurements of the reference's own code:
       t(window) = sample-pairs(window) / pair rate  +  eig+LU(n_t)
    pair rate: sample windows through run_distmix at the full N (the correlation loops are linear in sample-pairs,
    SURVEY.md section 8d); eig+LU: timed at the workload's min / mean / (capped) max n_t, interpolated in between and
    cubic beyond.  pair_samples: (sample-pairs, seconds, kind, n_t) per sample window."""
    nt, nu, sizes, _ = window_sizes(workload, args)
    N = float(sizes.sum())
    ns = np.array(sorted(eig_times))
    ts = np.array([eig_times[k] for k in ns])
    coef = float(np.mean(ts / ns.astype(float) ** 3))          # s per n^3

    def eig(n):
        if n <= ns[0]:
            return float(ts[0] * (n / ns[0]) ** 3)
        return float(np.interp(n, ns, ts)) if n <= ns[-1] else float(ts[-1] * (n / ns[-1]) ** 3)

    # the sample windows' own eig+LU share is taken out of their time before the pair rate is formed
    rate = sum(s[0] for s in pair_samples) / sum(max(s[1] - eig(s[3]), 1e-3) for s in pair_samples)
    pairs = (nt * (nt - 1) / 2 + nu * nt + (nt + nu)) * N
    t_pairs = float(pairs.sum() / rate)
    t_eig = float(sum(eig(int(n)) for n in nt))
    return dict(core_seconds=t_pairs + t_eig, pair_seconds=t_pairs, eig_lu_seconds=t_eig, pair_rate=rate,
                eig_lu_ns_per_n3=coef * 1e9, n_imputed=int(nu.sum()), nt_min=int(nt.min()), nt_mean=float(nt.mean()),
                nt_max=int(nt.max()), eig_points={int(k): float(v) for k, v in eig_times.items()})


def eig_points(nt, cap=1200):
    return sorted({int(nt.min()), int(round(nt.mean())), int(min(nt.max(), cap))})


def cpu_baseline_one_core(workload, args):
    nt, _, _, _ = window_sizes(workload, args)
    n_s = int(min(max(nt.min(), 200), 320))
    samples = [_pair_rate_worker((5, n_s, 96))]
    eig = {n: _eig_lu_seconds(n) for n in eig_points(nt)}
    m = cpu_model(workload, args, samples, eig)
    return dict(value=m["n_imputed"] / m["core_seconds"], unit=UNIT, cores=1, kind=samples[0][2],
                sample=(f"run_distmix of oracle/_ref on 1 window (n_t={n_s} x 96 unmeasured SNPs, N=32147, 21 pops: {samples[0][1]:.1f} s "
                        f"-> {m['pair_rate']:.3g} sample-pairs/s) + eig+LU timed at n_t={list(m['eig_points'])} "
                        f"({', '.join(f'{v:.2f} s' for v in m['eig_points'].values())}; cubic beyond the largest); step = "
                        f"{m['pair_seconds'] / 3600:.1f} core-hours of correlation loops + {m['eig_lu_seconds'] / 3600:.2f} of "
                        f"eig+LU over windows with n_t {m['nt_min']}/{m['nt_mean']:.0f}/{m['nt_max']} (min/mean/max)"))


def run_reference(args):
    """Reference arm: the reference's own CPU code path (oracle/_ref = its CalWgtCov / run_distmix compiled from
    /root/reference/src; the C port only if that library is absent) on ALL host cores of the box.  The reference is
    single-threaded, so "all the host threads it can use" = one window per core, the way a user fans the R loop out.
    A step = one bounded sample window per core (pair-loop rate at the full N); the O(n_t^3) eig + LU part is timed
    once, in the warm-up, at the workload's min / mean / (capped) max n_t and added per window (cpu_model)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workload = "genome" if args.workload in ("auto", "genome") else "chr22"
    nt, nu, sizes, _ = window_sizes(workload, args)
    n_s = int(min(max(nt.min(), 200), 320))
    vals, last_ms, kind, model = [], 0.0, "port", None
    with mp.get_context("spawn").Pool(cores) as pool:
        eig_async = pool.map_async(_eig_lu_seconds, eig_points(nt))
        eig = None
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_pair_rate_worker, [(5 + i * cores + k, n_s, 48) for k in range(cores)])
            wall = time.perf_counter() - t0
            if eig is None:
                eig = dict(zip(eig_points(nt), eig_async.get()))
            # the workers overlap in time: per-core time = the model's core-seconds / cores
            m = cpu_model(workload, args, res, eig)
            if i >= args.warmup:
                vals.append(m["n_imputed"] / (m["core_seconds"] / cores))
            last_ms, kind, model = wall * 1e3, res[0][2], m
    v = float(np.mean(vals))
    step_ms = model["core_seconds"] / cores * 1e3
    desc = (f"{cores} windows at once, one per core (n_t={n_s}, 48 unmeasured SNPs each, N=32147, 21 pops) through oracle/_ref's "
            f"run_distmix: {model['pair_rate']:.3g} sample-pairs/s per core; eig+LU timed at n_t={list(model['eig_points'])} "
            f"({', '.join(f'{t:.2f} s' for t in model['eig_points'].values())}); step = {model['pair_seconds'] / 3600:.1f} core-hours "
            f"of correlation loops + {model['eig_lu_seconds'] / 3600:.2f} of eig+LU (windows n_t {model['nt_min']}/"
            f"{model['nt_mean']:.0f}/{model['nt_max']}), divided over {cores} cores")
    cfg = genome_config(genome_chroms(args), args.gpus, "reference") if workload == "genome" else chr22_config()
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=step_ms, sample_ms_per_step=last_ms, higher_is_better=True,
                scaling="strong" if workload == "genome" else "weak", vs_baseline=None,
                dtype="f64", data="synthetic", impl="reference", config=cfg,
                cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind=kind, sample=desc),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(index: int) -> str:
    """Multi-rank runs: keep this process on the CPUs NVML reports as local to its GPU.  Host-side plumbing only."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return "nvml reports no local cpus inside this process's mask"
        os.sched_setaffinity(0, cpus)
        return f"bound to {len(cpus)} cpus local to GPU {index}"
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__}: {e})"


# ------------------------------------------------------------------------------------------------ GPU arm
class Env:
    """Rank / device / timing plumbing shared by the workloads."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = bind_to_gpu_numa_node(self.local) if self.world > 1 else "not bound (1 rank)"
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            from datetime import timedelta
            # a rank that dies must not leave the others spinning for NCCL's default 10 minutes
            dist.init_process_group("nccl", device_id=self.dev, timeout=timedelta(seconds=240))
            self.dist = dist

    def barrier(self, collective=True):
        if self.dist is not None and collective:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, vals, op):
        t = self.torch.tensor(list(vals), device=self.dev, dtype=self.torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return [float(x) for x in t]


def u64_checksum(arrs):
    """Order-independent exact checksum of float64 results: sum of the bit patterns mod 2^64 (placement of windows on
    GPUs / ranks / batches cannot change it unless a result changes)."""
    s = np.uint64(0)
    with np.errstate(over="ignore"):
        for a in arrs:
            s = s + np.ascontiguousarray(a).view(np.uint64).sum(dtype=np.uint64)
    return int(s)


def measure_probes(ctx):
    pk = peaks()
    out = dict(source="gb_probe_peak on this GPU at bench time (tensor pipe alone: one thread per SM issuing back-to-back "
                      "tcgen05.mma 128x128 from shared memory, no TMA, no epilogue; fp64: mma.sync m8n8k4 from registers, "
                      "16 warps per SM; copy: 1 GiB device copy, read + write bytes)")
    for k in ("i8", "mxf4", "fp64", "copy"):
        try:
            out[k] = ctx.probe_peak(k)
        except Exception as e:  # noqa: BLE001
            out[k] = None
            out[k + "_error"] = str(e)
    out["units"] = dict(i8="TOP/s", mxf4="TOP/s", fp64="TFLOP/s", copy="GB/s")
    out["measured_peaks_json"] = pk
    try:
        json.dump(out, open(os.path.join(ROOT, "MEASURED_PEAKS_TENSOR.json"), "w"), indent=1)
    except OSError:
        pass
    return out


# NCU figures of the kernels below (one `ncu --set full` capture each, profiles/, not re-measured here): tensor-pipe
# utilisation and DRAM bytes per launch on the chr22-shaped batch
NCU = dict(
    gram=dict(tensor_pipe_pct=None, dram_bytes=3.676e9, file="profiles/r01_final_ncu_full.md"),
    trsm=dict(dram_bytes=5.96e9, file="profiles/r01_final_ncu_full.md"),
)
_ncu_path = os.path.join(ROOT, "profiles", "r02_ncu_figures.json")
if os.path.exists(_ncu_path):
    try:
        for _k, _v in json.load(open(_ncu_path)).items():
            NCU.setdefault(_k, {}).update(_v)
    except Exception:  # noqa: BLE001
        pass


def chr22_batch_inputs(sizes):
    bp, type_, windows, bp_m, z_m = chr22_layout()
    n_m = int((type_ == 1).sum())
    n_all = len(bp)
    # panel layout: [measured block | unmeasured block], each in bp order -> every window is two
    # contiguous row ranges and TMA reads the panel directly (no gather)
    pos = np.empty(n_all, np.int64)
    pos[type_ == 1] = np.arange(n_m)
    pos[type_ == 0] = n_m + np.arange(n_all - n_m)
    meas_idx = np.where(type_ == 1)[0]
    z_by_site = np.zeros(n_all)
    z_by_site[meas_idx] = z_m
    t_off, u_off, rows_t, rows_u, z_t = [0], [0], [], [], []
    for x in windows:
        rows_t.append(pos[x["measured"]])
        rows_u.append(pos[x["unmeasured"]])
        z_t.append(z_by_site[x["measured"]])
        t_off.append(t_off[-1] + len(x["measured"]))
        u_off.append(u_off[-1] + len(x["unmeasured"]))
    sites = np.concatenate([meas_idx, np.where(type_ == 0)[0]]).astype(np.int64)     # bp-order rank of every row
    return dict(n_all=n_all, n_m=n_m, windows=windows, pos=pos, z_by_site=z_by_site, sites=sites,
                t_off=np.array(t_off), u_off=np.array(u_off), rows_t=np.concatenate(rows_t),
                rows_u=np.concatenate(rows_u), z_t=np.concatenate(z_t))


def converter_rate(sizes, g):
    """gb_packfile_convert on a BGZF file in the reference's data format (gauss.cpp:572-585), written here with zlib
    the way bgzf.c:280-330 frames its blocks: GB/s of panel text through inflate + parse + ternary pack on host threads."""
    import struct
    import zlib
    from gauss_b200 import packfile
    offs = np.concatenate([[0], np.cumsum(sizes)])
    chars = (g + 48).astype(np.uint8)
    lines = []
    for row in chars:
        strs = [row[offs[k]:offs[k + 1]].tobytes() for k in range(len(sizes))]
        afs = [b"%.6f" % (float(g_.mean()) / 2) for g_ in (row[offs[k]:offs[k + 1]].astype(np.float64) - 48 for k in range(len(sizes)))]
        lines.append(b" ".join(strs + afs) + b"\n")
    data = b"".join(lines)
    tmp = tempfile.mkdtemp()
    geno, desc, outp = os.path.join(tmp, "p_geno.gz"), os.path.join(tmp, "desc.txt"), os.path.join(tmp, "p.gbpack")
    with open(desc, "w") as f:
        f.write("pop n sup\n" + "".join(f"P{k} {int(m)} S\n" for k, m in enumerate(sizes)))
    with open(geno, "wb") as f:
        for i in list(range(0, len(data), 0xff00)) + [None]:
            c = b"" if i is None else data[i:i + 0xff00]
            co = zlib.compressobj(6, zlib.DEFLATED, -15)
            payload = co.compress(c) + co.flush()
            f.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, 12 + 6 + len(payload) + 8 - 1))
            f.write(payload + struct.pack("<II", zlib.crc32(c), len(c)))
    info = packfile.convert_reference_panel_native(geno, desc, outp)
    pf = packfile.PackFile5(outp)
    ok = bool(np.array_equal(np.asarray(pf.rows), __import__("gauss_b200").api.pack5_rows_host(sizes, g, is_ascii=False)))
    for pth in (geno, desc, outp):
        os.unlink(pth)
    os.rmdir(tmp)
    return dict(rows=int(info["n_rows"]), text_gb=info["text_bytes"] / 1e9, seconds=info["seconds"], threads=os.cpu_count(),
                text_gbs=info["text_bytes"] / info["seconds"] / 1e9, compressed_mb=None, rows_equal_host_packer=ok,
                note="gb_packfile_convert: BGZF inflate + line parse + ternary pack, block-parallel on host threads; "
                     "the reference re-inflates and re-parses this text on every call (ReadGenotype, gauss.cpp:720-785)")


def stage_times(env, batch, stream, steps, stages):
    torch = env.torch
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(steps)]
    for k in range(steps):
        for i, s in enumerate(stages):
            evs[k][i].record(stream)
            batch.run_stage(s)
        evs[k][len(stages)].record(stream)
    torch.cuda.synchronize()
    return np.array([[evs[k][s].elapsed_time(evs[k][s + 1]) for s in range(len(stages))] for k in range(steps)])


def bench_chr22(env, args, ctx, stream, sizes, w, probes, with_e2e, fmt=None, tag="chr22", collective=True):
    """One chromosome-22-shaped batch resident in HBM: value, per-stage times, rooflines, host-buffer legs.
    collective=False: only this rank runs it (the kernel block of the genome workload) -- no cross-rank barrier."""
    import gauss_b200 as gb
    from gauss_b200 import api
    torch = env.torch
    dev = env.dev
    N = int(sizes.sum())
    L = chr22_batch_inputs(sizes)
    n_all = L["n_all"]
    out = {}
    # ---- synthetic panel generated on the device as ternary rows, expanded once (K0c)
    row5 = api.pack5_row_bytes(sizes)
    t0 = time.time()
    d_rows5 = torch.empty((n_all, row5), dtype=torch.uint8, device=dev)
    api.synth_pack5_rows_device(ctx, SEED + 22 + 1000 * env.rank, 21, sizes, n_all, d_rows5.data_ptr(), row5, sites=L["sites"])
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    panel = gb.Panel(ctx, sizes, n_all, "e2m1")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    exp_ms = []
    for _ in range(3):
        panel.clear()
        ev0.record(stream)
        panel.append_pack5_device_ptr(d_rows5.data_ptr(), n_all, row5)
        ev1.record(stream)
        torch.cuda.synchronize()
        exp_ms.append(ev0.elapsed_time(ev1))
    k_stride = ((sizes + 127) // 128 * 128).sum() // 2
    exp_bytes = n_all * (row5 + k_stride + 8 * len(sizes))
    pk = peaks()
    out["roofline_expand5"] = dict(bound="hbm", kernel="expand5_rows_kernel", achieved=exp_bytes / (min(exp_ms) / 1e3) / 1e9,
                                   peak=pk["hbm_gbs"], unit="GB/s", frac=exp_bytes / (min(exp_ms) / 1e3) / 1e9 / pk["hbm_gbs"],
                                   traffic=NCU.get("expand5", {}).get("dram_bytes"), ms=min(exp_ms), rows=n_all,
                                   note=f"ternary resident rows ({row5} B) -> E2M1 operand rows ({int(k_stride)} B) + per-population "
                                        f"sum x / sum x^2: algorithmic bytes {exp_bytes / 1e9:.2f} GB per launch; peak = {pk['source']} hbm_gbs "
                                        f"(copy probe on this box: {probes.get('copy')})")
    if fmt == "int8":
        # the int8 arm: same rows as one signed byte per dosage (kind::i8, int32 accumulate)
        host5 = torch.empty((n_all, row5), dtype=torch.uint8)
        host5.copy_(d_rows5)
        g8 = torch.from_numpy(api.unpack5_rows(host5.numpy(), sizes))
        panel = gb.Panel(ctx, sizes, n_all, "int8")
        panel.append_host(g8.numpy(), is_ascii=False)
        del g8, host5
    batch = gb.Batch(panel, L["t_off"], L["rows_t"], L["u_off"], L["rows_u"], L["z_t"], w)
    work = batch.work()
    windows = L["windows"]
    n_imputed = int(sum(len(x["unmeasured"]) for x in windows if len(x["measured"]) > 10 and len(x["unmeasured"]) > 10))
    n_u_all = int(L["u_off"][-1])
    z_pin = torch.empty(n_u_all, dtype=torch.float64, pin_memory=True)
    i_pin = torch.empty(n_u_all, dtype=torch.float64, pin_memory=True)

    def step():
        batch.run()
        return batch.fetch(z_pin.numpy(), i_pin.numpy())     # D2H of z / info / statuses: part of the resident step (SURVEY 8d i)

    for _ in range(args.warmup):
        step()
    env.barrier(collective)
    sampler = ClockSampler(env.local)
    sampler.start()
    launches0 = ctx.launch_count
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(stream)
    for _ in range(args.steps):
        z, info, status = step()
    e_end.record(stream)
    torch.cuda.synchronize()
    env.barrier(collective)
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    total_ms = e_start.elapsed_time(e_end)
    z, info = z.copy(), info.copy()
    n_ok = int((status == 0).sum())
    # ---- the same K steps stage by stage (serialised, every kernel on all SMs)
    # row statistics | Gram tensor-core kernel | Gram finish pass | factorisation chain (Cholesky with the rows of L^-1
    # growing beside its steps on an auxiliary stream: int8-split solve only) | solve
    STAGES = [0, 10, 11, 2, 3]
    st = stage_times(env, batch, stream, args.steps, STAGES)
    gram_ms, fin_ms, chain_ms, trsm_ms = (float(st[:, i].mean()) for i in (1, 2, 3, 4))
    # ... and the two halves of the chain alone, one after the other (what each costs without the overlap)
    st2 = stage_times(env, batch, stream, args.steps, [20, 21])
    chol_ms, trtri_ms = float(st2[:, 0].mean()), float(st2[:, 1].mean())
    ozaki = os.environ.get("GB_SOLVE", "ozaki") != "fp64"
    nts = np.array([len(x["measured"]) for x in windows], float)
    nus = np.array([len(x["unmeasured"]) for x in windows], float)
    okw = (nts > 10) & (nus > 10)
    trsm_flops = float((nts ** 2 * nus + nts ** 2 + 4 * nts * nus)[okw].sum())
    chol_flops = float((nts ** 3 / 3)[okw].sum())
    mhz = clocks.get("sm_max_mhz") or 1965.0
    fp64_peak = probes.get("fp64") or 148 * 128 * mhz * 1e6 / 1e12
    fp64_src = "gb_probe_peak fp64 (mma.sync m8n8k4, measured on this box now)" if probes.get("fp64") else "148 SM x 128 flop x SM clock"
    serial_ms = float(st.sum(1).mean())
    out.update(
        value=n_imputed * args.steps / (total_ms / 1e3), ms_per_step=total_ms / args.steps, launches=int(launches), clocks=clocks,
        n_imputed=n_imputed, windows_ok=n_ok, panel_gen_s=gen_s, work=work, z=z, info=info, status=status,
        stage_ms_serial=serial_ms,
        stage_ms=dict(row_stats=float(st[:, 0].mean()), gram=gram_ms, gram_finish=fin_ms, factorisation_chain=chain_ms, solve=trsm_ms),
        stage_ms_alone=dict(cholesky=chol_ms, linv=trtri_ms,
                            note="the chain's halves run one after the other: in the step the rows of L^-1 are formed on an auxiliary stream "
                                 "while the factorisation takes its next steps, so factorisation_chain < cholesky + linv"),
        solver=("int8-split GEMM on tcgen05 (kind::i8, six signed 8-bit digit planes per operand, the 26 digit pairs of weight >= 2^-64) "
                "behind an explicit L^-1") if ozaki else "fp64 DMMA triangular solve",
        roofline_chol=dict(bound="tensor", kernel="chol_diag/panel/update chain", achieved=chol_flops / (chol_ms / 1e3) / 1e12,
                           peak=fp64_peak, unit="TFLOP/s", frac=chol_flops / (chol_ms / 1e3) / 1e12 / fp64_peak, ms=chol_ms,
                           note=f"sum n_t^3 / 3 = {chol_flops:.4g} flops over the whole factorisation chain (latency-bound dependent launches)"),
    )
    is_fp4 = panel.format == "e2m1"
    gram_peak = probes.get("mxf4" if is_fp4 else "i8")
    rate = 4.0 if is_fp4 else 2.0
    ach = work["gram_ops"] / (gram_ms / 1e3) / 1e12
    rg = dict(bound="tensor", kernel="gram_seg_kernel<%s>" % ("kind::mxf4" if is_fp4 else "kind::i8"), achieved=ach,
              peak=gram_peak or rate * pk["bf16_tflops"], unit="TOP/s",
              frac=ach / (gram_peak or rate * pk["bf16_tflops"]), ms=gram_ms, share_of_step=gram_ms / serial_ms,
              frac_vs_bf16_burst_x=ach / (rate * pk["bf16_tflops"]), frac_vs_bf16_sustained_x=ach / (rate * pk["bf16_tflops_sustained"]),
              frac_vs_int8_spec_4500=ach / 4500.0, ncu_tensor_pipe_pct=NCU["gram"].get("tensor_pipe_pct"), ncu_file=NCU["gram"].get("file"),
              traffic=NCU["gram"].get("dram_bytes"),
              note=(f"algorithmic ops 2 N (n_u n_t + n_t (n_t + 1) / 2) summed over windows = {work['gram_ops']:.4g} per launch; peak = "
                    f"{'tensor-pipe probe of this instruction kind measured on this box now (gb_probe_peak)' if gram_peak else str(rate) + ' x burst bf16 of MEASURED_PEAKS.json (probe failed)'}; "
                    f"frac_vs_bf16_*_x = against {rate:g} x the bf16 figures of {pk['source']}; traffic = dram bytes per launch "
                    f"({NCU['gram'].get('file')}) vs {work['panel_bytes'] / (2 if is_fp4 else 1) / 1e9:.2f} GB of operand rows"))
    out["roofline_gram"] = rg
    if ozaki:
        # the solve as it runs now: explicit L^-1 (stage 21) + digit-plane GEMM (stage 3); the GEMM's own work is the 26 digit-pair
        # products over the triangular K range of L^-1
        n_pairs = 26
        gemm_ops = float(2 * n_pairs * (nus * nts * (nts + 1) / 2)[okw].sum())
        i8_peak = probes.get("i8") or 2.0 * pk["bf16_tflops"]
        out["roofline_solve"] = dict(
            bound="tensor", kernel="ozaki_solve_kernel (+ oz_slice_x_kernel)", achieved=gemm_ops / (trsm_ms / 1e3) / 1e12, peak=i8_peak,
            unit="TOP/s", frac=gemm_ops / (trsm_ms / 1e3) / 1e12 / i8_peak, ms=trsm_ms, share_of_step=trsm_ms / serial_ms,
            fp64_equivalent_tflops=trsm_flops / (trsm_ms / 1e3) / 1e12, fp64_probe_tflops=fp64_peak,
            ncu_tensor_pipe_pct=NCU.get("ozaki", {}).get("tensor_pipe_pct"), traffic=NCU.get("ozaki", {}).get("dram_bytes"),
            ncu_file=NCU.get("ozaki", {}).get("file"),
            note=(f"int8 ops of the GEMM itself: 2 x {n_pairs} digit pairs x sum n_u n_t (n_t + 1) / 2 = {gemm_ops:.4g} per launch against the "
                  f"kind::i8 pipe probe of this box; the stage time also holds the slicing of L^-1 into digit planes.  "
                  f"fp64_equivalent_tflops = the solve's algorithmic fp64 flops ({trsm_flops:.4g}, what the DMMA triangular solve spent "
                  f"at 0.67 of the fp64 probe) over the same time: the reason this path exists"))
        linv_flops = chol_flops   # n_t^3 / 3 per window: the triangular solve against the identity with the zero blocks skipped
        out["roofline_linv"] = dict(bound="tensor", kernel="linv_row_kernel (X = L^-1 row block by row block, y = L^-1 z)",
                                    achieved=linv_flops / (trtri_ms / 1e3) / 1e12,
                                    peak=fp64_peak, unit="TFLOP/s", frac=linv_flops / (trtri_ms / 1e3) / 1e12 / fp64_peak, ms=trtri_ms,
                                    ms_added_to_the_chain=chain_ms - chol_ms,
                                    note=f"sum n_t^3 / 3 = {linv_flops:.4g} flops; timed alone (one launch per 64-row block, each waiting for the "
                                         "previous: latency-bound); inside the step the launches ride beside the factorisation's own "
                                         "steps and add ms_added_to_the_chain")
        b21 = float((nts * nus)[okw].sum())
        b11 = float((nts * (nts + 1) / 2)[okw].sum())
        fin_bytes = b21 * (8 + 6) + b11 * 16
        if is_fp4:
          out["roofline_finish"] = dict(bound="hbm", kernel="gram_finalize_kernel", achieved=fin_bytes / (fin_ms / 1e3) / 1e9, peak=pk["hbm_gbs"],
                                        unit="GB/s", frac=fin_bytes / (fin_ms / 1e3) / 1e9 / pk["hbm_gbs"], ms=fin_ms,
                                        traffic=NCU.get("finish", {}).get("dram_bytes"), ncu_file=NCU.get("finish", {}).get("file"),
                                        note=f"B21: 8 B read + 6 digit-plane bytes written per entry ({b21:.4g} entries), B11: 8 + 8 B per lower-triangle "
                                             f"entry ({b11:.4g}); 21 DFMA + the digit split per entry keep it issue-bound, not HBM-bound (ncu: "
                                             f"{NCU.get('finish', {}).get('issue_active_pct')} % issue slots, {NCU.get('finish', {}).get('dram_pct')} % DRAM)")
        out["roofline"] = dict(rg)
        out["roofline"]["note"] = "longest kernel of the step (event-timed alone on all SMs); " + rg["note"]
        out["solve"] = dict(flops_per_step=work["solve_flops"], ms=chain_ms + trsm_ms,
                            tflops=work["solve_flops"] / ((chain_ms + trsm_ms) / 1e3) / 1e12,
                            note="factorisation chain (Cholesky + explicit L^-1) + int8-split GEMM; algorithmic fp64 flops of the reference's formulation over their time")
    else:
        out["roofline"] = dict(bound="tensor", kernel="trsm_finalize_kernel", achieved=trsm_flops / (trsm_ms / 1e3) / 1e12, peak=fp64_peak,
                               unit="TFLOP/s", frac=trsm_flops / (trsm_ms / 1e3) / 1e12 / fp64_peak,
                               traffic=NCU["trsm"].get("dram_bytes"), ms=trsm_ms, share_of_step=trsm_ms / serial_ms,
                               note=(f"longest kernel of the step ({tag} batch, event-timed alone); algorithmic flops sum(n_t^2 n_u + n_t^2 + "
                                     f"4 n_t n_u) = {trsm_flops:.4g} per launch (no Cholesky term); pipe = fp64 tensor core (DMMA m8n8k4); peak = "
                                     f"{fp64_src}; traffic = dram bytes per launch from {NCU['trsm'].get('file')} vs {8 * float((nts * nts + nts * nus)[okw].sum()) / 1e9:.2f} GB algorithmic"))
        out["solve"] = dict(flops_per_step=work["solve_flops"], ms=chol_ms + trsm_ms, tflops=work["solve_flops"] / ((chol_ms + trsm_ms) / 1e3) / 1e12)
    out["dtype"] = ("e2m1 x e2m1 -> f32 (exact integer counts, tcgen05 kind::mxf4)" if is_fp4 else "int8 x int8 -> int32 (tcgen05 kind::i8)") + " + f64 fold/solve"
    if not with_e2e:
        batch.close()
        panel.close()
        return out

    # ---- host-buffer legs (N = 1 only) ------------------------------------------------------------------------
    host5 = torch.empty((n_all, row5), dtype=torch.uint8, pin_memory=True)
    host5.copy_(d_rows5)
    torch.cuda.synchronize()
    del d_rows5
    n_groups = int(os.environ.get("GB_E2E_GROUPS", "6"))
    panel2 = gb.Panel(ctx, sizes, n_all, "e2m1")
    ok_u = np.concatenate([np.full(L["u_off"][i + 1] - L["u_off"][i], status[i] == 0) for i in range(len(windows))])
    e2e_steps = max(1, min(args.steps, 3))

    def chrom_step(rows_ptr):
        _, _, s2 = panel2.chrom_run_pack5(rows_ptr, n_all, row5, L["t_off"], L["rows_t"], L["u_off"], L["rows_u"], L["z_t"], w,
                                          n_groups=n_groups, z=z_pin.numpy(), info=i_pin.numpy())
        return s2

    s2 = chrom_step(host5.data_ptr())
    eq_chrom = bool(int((s2 == 0).sum()) == n_ok and np.array_equal(z_pin.numpy()[ok_u], z[ok_u]) and
                    np.array_equal(i_pin.numpy()[ok_u], info[ok_u]))
    chrom_step(host5.data_ptr())
    env.barrier(collective)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        chrom_step(host5.data_ptr())
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    c_h2d = n_all * row5 + (len(L["rows_t"]) + len(L["rows_u"])) * 4 + len(L["z_t"]) * 8 + len(w) * 8
    c_d2h = n_u_all * 16 + (2 * len(windows) + 3 * n_groups) * 4
    out["e2e"] = dict(value=n_imputed / e2e_s, unit=UNIT, h2d_bytes_per_step=int(c_h2d), d2h_bytes_per_step=int(c_d2h),
                      steps=e2e_steps, ms_per_step=e2e_s * 1e3, host_row_bytes=int(row5), equal_to_resident_run=eq_chrom,
                      note=f"gb_chrom_run_pack5 (C-ABI chromosome driver): pinned HOST ternary panel rows ({row5} B per SNP) -> H2D in "
                           f"{n_groups} chunks on a copy stream -> expand -> window batches as their rows land -> D2H of z/info; wall clock "
                           "around the blocking call; results compared bit for bit with the resident run (equal_to_resident_run); the packed rows are a cached packed panel")
    # cold: the host pack (int8 rows -> ternary rows, CPU threads) inside the timed region
    g8 = api.unpack5_rows(host5.numpy(), sizes)
    cold5 = torch.empty((n_all, row5), dtype=torch.uint8, pin_memory=True)
    t0 = time.perf_counter()
    api.pack5_rows_host(sizes, g8, is_ascii=False, out=cold5.numpy())
    pack_s = time.perf_counter() - t0
    chrom_step(cold5.data_ptr())
    cold_s = time.perf_counter() - t0
    out["e2e_cold"] = dict(value=n_imputed / cold_s, unit=UNIT, ms_per_step=cold_s * 1e3, host_pack_ms=pack_s * 1e3,
                           h2d_bytes_per_step=int(c_h2d), d2h_bytes_per_step=int(c_d2h), steps=1,
                           note=f"same call with the host-side pack INSIDE the timed region: gb_pack5_rows_host (int8 dosages -> ternary rows, "
                                f"{os.cpu_count()} CPU threads, {n_all * N / pack_s / 1e9:.2f} GB/s of dosage bytes) + gb_chrom_run_pack5")
    del cold5
    # per-window calls on pinned int8 host rows (the seam run_distmix has today)
    host8 = torch.empty((n_all, N), dtype=torch.int8, pin_memory=True)
    host8.numpy()[:] = g8
    # K0 on the device (chars / int8 rows -> packed operand rows + per-population sums), and the native converter of the
    # reference's BGZF text panel (host threads) -- the "packed-panel stream" of the north star
    out["pack"] = {}
    try:
        d8 = host8.to(env.dev)
        pk_panel = gb.Panel(ctx, sizes, n_all, "e2m1")
        pms = []
        for _ in range(3):
            pk_panel.clear()
            ev0.record(stream)
            pk_panel.append_device_ptr(d8.data_ptr(), n_all, N, is_ascii=False)
            ev1.record(stream)
            torch.cuda.synchronize()
            pms.append(ev0.elapsed_time(ev1))
        pbytes = n_all * (N + k_stride + 8 * len(sizes))
        out["pack"]["pack_rows_kernel"] = dict(ms=min(pms), gbs=pbytes / (min(pms) / 1e3) / 1e9, hbm_peak_gbs=pk["hbm_gbs"],
                                               frac=pbytes / (min(pms) / 1e3) / 1e9 / pk["hbm_gbs"],
                                               note="one byte per dosage in, E2M1 nibbles + sum x / sum x^2 out (read + write bytes)")
        pk_panel.close()
        del d8
        out["pack"]["converter"] = converter_rate(sizes, g8[:1500])
    except Exception as e:  # noqa: BLE001
        out["pack"]["error"] = f"{type(e).__name__}: {e}"
    max_rows = int(max(len(x["measured"]) + len(x["unmeasured"]) for x in windows))
    pipe = gb.Pipe(ctx, sizes, max_rows, depth=3)
    res_z = [np.zeros(len(x["unmeasured"])) for x in windows]
    res_i = [np.zeros(len(x["unmeasured"])) for x in windows]
    cnt = dict(h2d=0, d2h=0)

    def pw_step(count=False):
        pending = []
        for wi, x in enumerate(windows):
            a, b = len(x["measured"]), len(x["unmeasured"])
            if a <= 10 or b <= 10:
                continue
            m0, u0 = int(L["pos"][x["measured"][0]]), int(L["pos"][x["unmeasured"][0]])
            pending.append(pipe.submit_ptr(host8.data_ptr() + m0 * N, a, host8.data_ptr() + u0 * N, b, N, False,
                                           L["z_by_site"][x["measured"]], w, res_z[wi], res_i[wi]))
            if len(pending) >= 3:
                assert pipe.wait(pending.pop(0)) == 0
            if count:
                cnt["h2d"] += (a + b) * N + a * 8 + (a + b) * 8 + len(w) * 8
                cnt["d2h"] += b * 16 + 8
        for t in pending:
            assert pipe.wait(t) == 0

    pw_step()
    t0 = time.perf_counter()
    pw_step(count=True)
    pw_s = time.perf_counter() - t0
    out["e2e_per_window"] = dict(value=n_imputed / pw_s, unit=UNIT, h2d_bytes_per_step=int(cnt["h2d"]), d2h_bytes_per_step=int(cnt["d2h"]),
                                 steps=1, ms_per_step=pw_s * 1e3,
                                 note="per-window gb_pipe_submit / gb_pipe_wait (depth 3) on pinned host int8 rows: one window per call, "
                                      "8 bits per dosage over PCIe")
    pipe.close()
    del host8
    # the literal seam: gb_run_window_strings on '0'/'1'/'2' strings, one blocking call per window
    if not args.no_strings:
        P = len(sizes)
        offs = np.concatenate([[0], np.cumsum(sizes + 1)])[:-1]            # every population string is NUL-terminated
        chars = np.zeros((n_all, int((sizes + 1).sum())), np.uint8)
        col = 0
        for p_, m in enumerate(sizes):
            chars[:, offs[p_]:offs[p_] + m] = g8[:, col:col + m] + 48
            col += m
        base = chars.ctypes.data
        st_s, st_done, st_ok = 0.0, 0, True
        import ctypes as C
        for wi, x in enumerate(windows):
            a, b = len(x["measured"]), len(x["unmeasured"])
            if a <= 10 or b <= 10:
                continue
            rows = np.concatenate([L["pos"][x["measured"]], L["pos"][x["unmeasured"]]])
            ptrs = (base + rows[:, None] * chars.strides[0] + offs[None, :]).astype(np.uint64).ravel()
            type_ = np.concatenate([np.ones(a, np.int32), np.zeros(b, np.int32)])
            bpw = np.arange(a + b, dtype=np.int64)
            zz = np.concatenate([L["z_by_site"][x["measured"]], np.zeros(b)])
            inf = np.ones(a + b)
            nt_, nu_ = C.c_int(0), C.c_int(0)
            wv = np.ascontiguousarray(w, np.float64)
            t0 = time.perf_counter()
            rc = ctx.lib.gb_run_window_strings(ctx.h, a + b, type_.ctypes.data, bpw.ctypes.data, zz.ctypes.data, inf.ctypes.data,
                                               ptrs.ctypes.data, P, sizes.ctypes.data, wv.ctypes.data, 0, 10 ** 12, None,
                                               C.byref(nt_), C.byref(nu_))
            st_s += time.perf_counter() - t0
            lo = int(L["u_off"][wi])
            st_ok = st_ok and rc == 0 and nt_.value == a and nu_.value == b and np.array_equal(zz[a:], z[lo:lo + b])
            st_done += b
        out["e2e_strings"] = dict(value=st_done / st_s, unit=UNIT, steps=1, ms_per_step=st_s * 1e3, equal_to_resident_run=bool(st_ok),
                                  h2d_bytes_per_step=int(cnt["h2d"]), d2h_bytes_per_step=int(cnt["d2h"]),
                                  note="gb_run_window_strings, one blocking call per window on NUL-terminated '0'/'1'/'2' strings per population "
                                       "(what Snp::genotype_vec_ holds, snp.h:109): host concatenation + pinned staging + H2D + pack + window "
                                       "+ D2H inside the timed region; results compared bit for bit with the resident batch")
        del chars
    batch.close()
    panel.close()
    panel2.close()
    return out


def bench_genome(env, args, sizes, w):
    from gauss_b200 import api
    torch = env.torch
    single = args.single_process and env.world == 1
    n_local = args.gpus if single else 1
    devices = list(range(n_local)) if single else [env.local]
    n_parts = n_local * env.world
    t0 = time.time()
    chroms = genome_chroms(args)
    layout_s = time.time() - t0
    g = api.Genome(n_local, sizes, w, devices=devices)
    for c in chroms:
        g.add_chromosome(c["n_rows"], c["t_off"], c["rows_t"], c["u_off"], c["rows_u"], c["z_t"], sites=c["sites"])
    t0 = time.time()
    g.plan(n_parts, env.rank * n_local)
    plan_s = time.time() - t0
    t0 = time.time()
    g.fill_synthetic(SEED)
    fill_s = time.time() - t0
    infos = [g.shard_info(i) for i in range(n_local)]
    n_imp_local = sum(x["n_imputed"] for x in infos)
    z = [np.zeros(int(c["u_off"][-1])) for c in chroms]
    info = [np.zeros(int(c["u_off"][-1])) for c in chroms]
    status = [np.full(len(c["t_off"]) - 1, -1, np.int32) for c in chroms]
    for _ in range(args.warmup):
        g.run(z, info, status)
    env.barrier()
    sampler = ClockSampler(env.local)
    sampler.start()
    l0 = g.launch_count
    dev_ms, t0 = 0.0, time.perf_counter()
    per_gpu = np.zeros(n_local)
    for _ in range(args.steps):
        _, _, _, ms = g.run(z, info, status)
        dev_ms += float(ms.max())
        per_gpu += ms
    wall_ms = (time.perf_counter() - t0) * 1e3
    env.barrier()
    clocks = sampler.stop()
    launches = g.launch_count - l0
    dev_ms_max, wall_ms_max = env.reduce([dev_ms, wall_ms], "MAX")
    n_imp, launches_all, gram_ops, solve_flops = env.reduce(
        [n_imp_local, launches, sum(x["gram_ops"] for x in infos), sum(x["solve_flops"] for x in infos)], "SUM")
    # what this process computed: its windows only (status >= 0)
    mine_z, mine_i, bad = [], [], 0
    for c, zz, ii, st in zip(chroms, z, info, status):
        for wdx in np.where(st >= 0)[0]:
            a, b = c["u_off"][wdx], c["u_off"][wdx + 1]
            if st[wdx] == 0:
                mine_z.append(zz[a:b])
                mine_i.append(ii[a:b])
            elif st[wdx] not in (5, 6):
                bad += 1
    cks = [u64_checksum(mine_z) , u64_checksum(mine_i)]
    if env.dist is not None:
        t = torch.tensor([cks[0] & 0xFFFFFFFF, cks[0] >> 32, cks[1] & 0xFFFFFFFF, cks[1] >> 32, bad], device=env.dev, dtype=torch.int64)
        env.dist.all_reduce(t)
        lo0, hi0, lo1, hi1, bad = (int(x) for x in t)
        cks = [((hi0 << 32) + lo0) & (2 ** 64 - 1), ((hi1 << 32) + lo1) & (2 ** 64 - 1)]
    out = dict(value=n_imp * args.steps / (dev_ms_max / 1e3), ms_per_step=dev_ms_max / args.steps,
               wall_ms_per_step=wall_ms_max / args.steps, launches=int(launches_all), clocks=clocks, n_imputed=int(n_imp),
               genome=dict(layout_s=layout_s, plan_s=plan_s, fill_synthetic_s=fill_s, shards=infos if env.world == 1 else None,
                           rank0_shard=infos[0], device_ms_per_gpu=(per_gpu / args.steps).tolist(),
                           gram_tops=gram_ops * args.steps / (dev_ms_max / 1e3) / 1e12,
                           solve_tflops=solve_flops * args.steps / (dev_ms_max / 1e3) / 1e12,
                           windows_not_ok=int(bad), result_checksum_u64=dict(z=cks[0], info=cks[1]),
                           note="result_checksum_u64 = sum of the float64 bit patterns of every OK window's z / info mod 2^64: independent of "
                                "how windows are placed on GPUs, so it must be the same number at every N",
                           host_gather="each GPU thread copies its windows' results from pinned staging into the caller's per-chromosome "
                                       "arrays (inside wall_ms_per_step; the D2H itself is inside the device-timed ms_per_step)"))
    # ---- e2e: the same job from pinned HOST rows, upload inside the timed region
    if not args.no_e2e:
        avail = 0
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                avail = int(ln.split()[1]) * 1024
        row5 = api.pack5_row_bytes(sizes)
        need = sum(x["resident_rows"] for x in infos) * row5
        if need > 0.45 * avail / max(1, (env.world if not single else 1)):
            out["e2e"] = dict(value=None, unit=UNIT, h2d_bytes_per_step=int(need), d2h_bytes_per_step=int(n_imp_local * 16),
                              note=f"skipped: {need / 1e9:.0f} GB of pinned host rows would not fit this box's {avail / 1e9:.0f} GB of free RAM")
        else:
            ctx0 = api.Context(devices[0])
            bufs = []
            t0 = time.time()
            for gi in range(n_local):
                for ci, c in enumerate(chroms):
                    for lo, hi in g.resident_ranges(gi, ci):
                        buf = torch.empty((hi - lo, row5), dtype=torch.uint8, pin_memory=True)
                        if g.download_rows(gi, ci, lo, hi - lo, buf.data_ptr(), row5) != 0:   # rows kept expanded: generate again
                            for r0 in range(lo, hi, 262144):
                                r1 = min(hi, r0 + 262144)
                                api.synth_pack5_rows(ctx0, SEED, ci, sizes, r1 - r0, sites=c["sites"][r0:r1], out=buf.numpy()[r0 - lo:r1 - lo])
                        g.set_host_rows(ci, lo, hi - lo, buf.data_ptr(), row5)
                        bufs.append(buf)
            host_fill_s = time.time() - t0
            ctx0.close()
            z2 = [np.zeros_like(a) for a in z]
            i2 = [np.zeros_like(a) for a in info]
            e2e_steps = max(1, min(args.steps, 2))
            g.upload(wait=False)
            g.run(z2, i2, status)      # warm-up, also the equality check below
            eq_up = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(z + info, z2 + i2))
            env.barrier()
            up_ms = np.zeros(n_local)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                g.upload(wait=False)
                g.submit(z2, i2, status)
                _, up = g.wait()
                up_ms += up
            e2e_s = time.perf_counter() - t0
            env.barrier()
            e2e_s_max, = env.reduce([e2e_s], "MAX")
            h2d, = env.reduce([need + sum(len(c["rows_t"]) * 12 + len(c["rows_u"]) * 4 for c in chroms) // max(1, n_parts)], "SUM")
            out["e2e"] = dict(value=n_imp * e2e_steps / e2e_s_max, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(n_imp * 16),
                              steps=e2e_steps, ms_per_step=e2e_s_max / e2e_steps * 1e3, upload_ms_per_gpu=(up_ms / e2e_steps).tolist(),
                              host_rows_fill_s=host_fill_s, equal_to_resident_run=bool(eq_up),
                              note=f"gb_genome_upload (asynchronous) + gb_genome_submit / gb_genome_wait: every step copies each GPU's ternary panel "
                                   f"rows ({row5} B per SNP, pinned HOST memory) to the device again, batches start as their rows land, "
                                   "z / info are written to host arrays; wall clock, max over ranks; results compared bit for bit with the "
                                   "resident run.  One-shot cost of a genome job; a panel that stays resident across traits pays it once")
            del bufs
    g.close()
    return out, chroms


def run_gpu(args):
    import gauss_b200 as gb
    env = Env(args)
    torch = env.torch
    _, sizes, w = synth.flagged_33kg_pgc2()
    workload = "genome" if args.workload == "auto" else args.workload
    ctx = gb.Context(env.local)
    stream = torch.cuda.Stream(env.dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    probes = measure_probes(ctx) if env.rank == 0 else {}
    line = dict(metric=METRIC, unit=UNIT, n_gpus=args.gpus if (args.single_process and env.world == 1) else env.world,
                steps=args.steps, warmup=args.warmup, higher_is_better=True, vs_baseline=None, data="synthetic")

    if workload == "genome":
        try:
            gout, chroms = bench_genome(env, args, sizes, w)
        except Exception as e:  # noqa: BLE001
            if env.world > 1:
                raise
            print(f"[bench] genome workload failed ({type(e).__name__}: {e}); falling back to the chr22 workload", file=sys.stderr)
            workload = "chr22"
    if workload == "genome":
        mode = "one process, one host thread per GPU" if (args.single_process and env.world == 1) else "one process per GPU"
        line.update(value=gout["value"], ms_per_step=gout["ms_per_step"], wall_ms_per_step=gout["wall_ms_per_step"], scaling="strong",
                    config=genome_config(chroms, line["n_gpus"], mode), gpu_launches=gout["launches"], clocks=gout["clocks"],
                    imputed_per_step=gout["n_imputed"], genome=gout["genome"], host_affinity=env.numa)
        if "e2e" in gout:
            line["e2e"] = gout["e2e"]
        if env.rank == 0:
            # kernel-level numbers on one chromosome-22-shaped batch (the round-1 step): per-stage times and rooflines;
            # at N = 1 also its host-buffer legs and the int8 arm
            full = env.world == 1 and not args.quick
            c22 = bench_chr22(env, args, ctx, stream, sizes, w, probes, with_e2e=full and not args.no_e2e, collective=False)
            line["dtype"] = c22["dtype"]
            for k in ("roofline", "roofline_gram", "roofline_solve", "roofline_linv", "roofline_finish", "roofline_chol", "roofline_expand5",
                      "stage_ms", "stage_ms_alone", "stage_ms_serial", "solve", "solver"):
                line[k] = c22[k]
            line["chr22"] = dict(value=c22["value"], ms_per_step=c22["ms_per_step"], imputed_per_step=c22["n_imputed"],
                                 gpu_launches=c22["launches"], config=chr22_config(),
                                 **{k: c22[k] for k in ("e2e", "e2e_cold", "e2e_per_window", "e2e_strings", "pack") if k in c22})
            if full:
                i8 = bench_chr22(env, args, ctx, stream, sizes, w, probes, with_e2e=False, fmt="int8", tag="chr22 int8", collective=False)
                dz = float(np.nanmax(np.abs(i8["z"] - c22["z"])))
                line["int8"] = dict(value=i8["value"], ms_per_step=i8["ms_per_step"], dtype=i8["dtype"], stage_ms=i8["stage_ms"],
                                    roofline_gram=i8["roofline_gram"], max_abs_dz_vs_e2m1=dz,
                                    note="the same chr22 step on a GB_PANEL_INT8 panel (kind::i8, int32 accumulate): the layout the north star names")
    elif workload == "chr22":
        c22 = bench_chr22(env, args, ctx, stream, sizes, w, probes, with_e2e=not args.no_e2e)
        tot_ms, = env.reduce([c22["ms_per_step"]], "MAX")
        n_all, = env.reduce([c22["n_imputed"]], "SUM")
        line.update(value=n_all / (tot_ms / 1e3), ms_per_step=tot_ms, scaling="weak", dtype=c22["dtype"], config=chr22_config(),
                    gpu_launches=c22["launches"], clocks=c22["clocks"], imputed_per_step=c22["n_imputed"], host_affinity=env.numa)
        for k in ("e2e", "e2e_cold", "e2e_per_window", "e2e_strings", "pack", "roofline", "roofline_gram", "roofline_solve", "roofline_linv",
                  "roofline_finish", "roofline_chol", "roofline_expand5", "stage_ms", "stage_ms_alone", "stage_ms_serial", "solve", "solver"):
            if k in c22:
                line[k] = c22[k]
    elif workload in ("ld5000", "dist1kg"):
        line.update(bench_small(env, args, ctx, stream, probes, workload))
    line["probes"] = probes
    if env.rank == 0:
        if env.world == 1 and not args.no_cpu_baseline and workload in ("genome", "chr22"):
            line["cpu_baseline"] = cpu_baseline_one_core(workload, args)
        print(json.dumps(line), flush=True)
    if env.dist is not None:
        env.dist.destroy_process_group()


def bench_small(env, args, ctx, stream, probes, workload):
    """BASELINE config 3 (computeLD, 5,000 SNPs, 33KG-shaped) and config 1 (dist() chr22, 1KG-shaped panel)."""
    import gauss_b200 as gb
    from gauss_b200 import api
    torch = env.torch
    pk = peaks()
    if workload == "ld5000":
        _, sizes, w = synth.flagged_33kg_pgc2()
        n = 5000
        row5 = api.pack5_row_bytes(sizes)
        d5 = torch.empty((n, row5), dtype=torch.uint8, device=env.dev)
        api.synth_pack5_rows_device(ctx, SEED, 0, sizes, n, d5.data_ptr(), row5)
        panel = gb.Panel(ctx, sizes, n, "e2m1")
        panel.append_pack5_device_ptr(d5.data_ptr(), n, row5)
        batch = gb.Batch(panel, [0, n], np.arange(n), None, None, None, w, ld_diag=1.0)
        for _ in range(args.warmup):
            batch.run_stage(0), batch.run_stage(1)
        st = stage_times(env, batch, stream, args.steps, [0, 10, 11])
        N = float(sizes.sum())
        ops = 2.0 * N * n * (n + 1) / 2
        gram_ms, fin_ms = float(st[:, 1].mean()), float(st[:, 2].mean())
        tot = float(st.sum(1).mean())
        ach = ops / (gram_ms / 1e3) / 1e12
        peak = probes.get("mxf4") or 4 * pk["bf16_tflops"]
        # host-in / host-out call
        rows5 = torch.empty((n, row5), dtype=torch.uint8, pin_memory=True)
        rows5.copy_(d5)
        t0 = time.perf_counter()
        p2 = gb.Panel(ctx, sizes, n, "e2m1")
        p2.append_pack5_host(rows5.numpy())
        cm, rc = p2.window_ld(np.arange(n), w)
        e2e_s = time.perf_counter() - t0
        assert rc == 0 and np.array_equal(np.diag(cm), np.ones(n))
        return dict(metric="computeLD SNP pairs/sec", unit="pairs/s", value=n * (n - 1) / 2 / (tot / 1e3), ms_per_step=tot, scaling="weak",
                    dtype="e2m1 x e2m1 -> f32 (exact integer counts, tcgen05 kind::mxf4) + f64 mixture epilogue",
                    config=dict(workload="computeLD (BASELINE config 3): dense 5,000-SNP block, 33KG-shaped panel 21 pops / 32,147 indiv, "
                                         "PGC2 weights; Gram + mixture epilogue only", windows=1, l2="operands 84 MB + 200 MB output > L2"),
                    stage_ms=dict(row_stats=float(st[:, 0].mean()), gram=gram_ms, gram_finish=fin_ms),
                    roofline=dict(bound="tensor", kernel="gram_seg_kernel<kind::mxf4>", achieved=ach, peak=peak, unit="TOP/s", frac=ach / peak,
                                  traffic=None, note=f"algorithmic ops 2 N n (n + 1) / 2 = {ops:.4g} (symmetric half); peak = tensor-pipe probe"),
                    e2e=dict(value=n * (n - 1) / 2 / e2e_s, unit="pairs/s", h2d_bytes_per_step=int(n * row5), d2h_bytes_per_step=int(n * n * 8),
                             note="gb_panel_append_pack5_host + gb_window_ld: ternary host rows in, 5,000 x 5,000 fp64 matrix out (pageable host memory)"),
                    gpu_launches=3 * args.steps, clocks={})
    # dist1kg
    res = {}
    for name, sz in (("EUR", synth.flagged_1kg("EUR")), ("ALL", np.array([n for _, n, _ in synth.POPS_1KG], np.int32))):
        L = chr22_batch_inputs(sz)
        row5 = api.pack5_row_bytes(sz)
        d5 = torch.empty((L["n_all"], row5), dtype=torch.uint8, device=env.dev)
        api.synth_pack5_rows_device(ctx, SEED, 21, sz, L["n_all"], d5.data_ptr(), row5, sites=L["sites"])
        panel = gb.Panel(ctx, sz, L["n_all"], "e2m1")
        panel.append_pack5_device_ptr(d5.data_ptr(), L["n_all"], row5)
        batch = gb.Batch(panel, L["t_off"], L["rows_t"], L["u_off"], L["rows_u"], L["z_t"], None)
        for _ in range(args.warmup):
            batch.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            batch.run()
            z, info, status = batch.fetch()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        n_imp = int(sum(len(x["unmeasured"]) for x in L["windows"] if len(x["measured"]) > 10 and len(x["unmeasured"]) > 10))
        st = stage_times(env, batch, stream, args.steps, [0, 10, 11, 2, 3])
        res[name] = dict(individuals=int(sz.sum()), pops=len(sz), value=n_imp / (ms / 1e3), ms_per_step=ms, windows_ok=int((status == 0).sum()),
                         stage_ms=dict(zip(["row_stats", "gram", "gram_finish", "cholesky", "solve"], [float(x) for x in st.mean(0)])),
                         gram_tops=batch.work()["gram_ops"] / (float(st[:, 1].mean()) / 1e3) / 1e12)
        batch.close(), panel.close()
    return dict(metric="dist imputed SNPs/sec", unit=UNIT, value=res["EUR"]["value"], ms_per_step=res["EUR"]["ms_per_step"], scaling="weak",
                dtype="e2m1 x e2m1 -> f32 (exact integer counts, pooled segment) + f64 CalCor epilogue / solve",
                config=dict(workload="dist() chr22 (BASELINE config 1): 36 x 1 Mb windows at the bundled PGC2 positions, 1KG-shaped panel, "
                                     "study_pop=EUR (503 indiv / 5 pops); `all2504` = every population flagged", windows=36,
                            l2="panel slice 36 MB (EUR) fits L2"),
                all2504=res["ALL"], eur=res["EUR"], gpu_launches=0, clocks={})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gauss_b200", choices=["gauss_b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "genome", "chr22", "ld5000", "dist1kg"])
    ap.add_argument("--genome-mb", default="all", help="comma-separated chromosome lengths in Mb (tuning runs); all = hg19")
    ap.add_argument("--single-process", action="store_true", help="drive all --gpus GPUs from this one process (gb_genome)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel tuning runs: skip the host-buffer legs")
    ap.add_argument("--no-strings", action="store_true", help="skip the gb_run_window_strings leg")
    ap.add_argument("--quick", action="store_true", help="genome workload: skip the chr22 host-buffer legs and the int8 arm")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
