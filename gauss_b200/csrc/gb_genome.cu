// gb_genome.cu -- genome-wide dist()/distmix() from ONE process on 1..8 GPUs (BASELINE config 4).
//
// The reference runs one window per dist()/distmix() call and every call rebuilds all of its state
// (dist.cpp:63-75), so windows are independent and a genome run is a user-level loop over ~2,900 windows.  Here the
// whole window list is cut into contiguous, cost-balanced shards, one per GPU; every GPU has its own host thread,
// context and streams, keeps the panel rows its windows touch (plus the wings of the boundary windows) RESIDENT in
// HBM -- uploaded or generated once -- and writes its windows' (z, info) straight into the caller's arrays: the
// "gather" is a host memcpy.  No collective, no peer traffic.
//
// Residency.  A row is kept in the ternary format of the packed-panel file ("pack5", 1.6 bits per dosage): 10 M SNPs
// of the 33KG shape are 65 GB, so even the whole genome fits one B200.  When the shard is small enough (2+ GPUs) the
// rows are expanded once into the E2M1 operand layout the Gram kernel reads (16.7 KB per SNP) and stay that way;
// otherwise each batch of windows expands the rows it touches into one of two working panels right before it runs
// (an HBM-bound kernel, a few per cent of a batch).  Batches alternate between two compute streams, each with its own
// working panel and workspace, so the latency-bound factorisation chain of one batch runs under the Gram / solve
// kernels of the other.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <thread>

#include <nvtx3/nvToolsExt.h>

#include "gb_batch.cuh"

using namespace gb;

namespace gb {
int launch_synth_pack5(Ctx* ctx, uint8_t* dst, int64_t dst_stride, int64_t n_rows, const int64_t* d_sites,
                       int64_t first_site, int n_pops, const int* d_pop_sizes, const int* d_boff5, int row_bytes,
                       uint64_t seed, int chrom);

double window_cost(double n_t, double n_u, double n_samples, const gb_params& p) {
  if (n_t <= p.min_num_measured_snp || n_u <= p.min_num_unmeasured_snp) return 0.0;
  const double gram = n_samples * (n_u * n_t + 0.5 * n_t * n_t);
  const double solve = n_t * n_t * n_u + n_t * n_t * n_t / 3.0;
  return gram + 80.0 * solve;   // measured on B200 (chr22 step): 7.1e-13 ms per Gram multiply-add, 5.4e-11 ms per solve flop
}

void partition_contiguous(const double* cost, int64_t n, int n_parts, int64_t* cuts) {
  for (int i = 0; i <= n_parts; i++) cuts[i] = n;
  cuts[0] = 0;
  if (n == 0 || n_parts < 1) return;
  double lo = 0.0, hi = 0.0;
  for (int64_t i = 0; i < n; i++) {
    lo = std::max(lo, cost[i]);
    hi += cost[i];
  }
  auto parts_for = [&](double limit, int64_t* out) {
    int parts = 1;
    double acc = 0.0;
    for (int64_t i = 0; i < n; i++) {
      if (acc + cost[i] > limit && acc > 0.0) {
        if (out && parts < n_parts) out[parts] = i;
        parts++;
        acc = 0.0;
      }
      acc += cost[i];
    }
    return parts;
  };
  for (int it = 0; it < 60; it++) {
    const double mid = 0.5 * (lo + hi);
    if (parts_for(mid, nullptr) <= n_parts) hi = mid;
    else lo = mid;
  }
  parts_for(hi * (1.0 + 1e-12), cuts);
  for (int i = 1; i <= n_parts; i++) cuts[i] = std::max(cuts[i], cuts[i - 1]);
  cuts[n_parts] = n;
}
}  // namespace gb

namespace {

struct RowRange {
  int64_t lo = 0, hi = 0;     // chromosome rows [lo, hi)
  int64_t res = 0;            // first row of the range in the shard's resident buffer
  int64_t issued = 0;         // rows [lo, issued) have been uploaded / generated (upload cursor)
};

struct HostPiece {            // ternary HOST rows [lo, hi) of a chromosome
  int64_t lo = 0, hi = 0;
  const uint8_t* ptr = nullptr;
  int64_t stride = 0;
};

struct Chrom {
  int64_t n_rows = 0;
  std::vector<HostPiece> pieces;
  int64_t n_windows = 0;
  std::vector<int64_t> t_off, u_off, rows_t, rows_u, sites;
  std::vector<double> z_t;
};

struct Segment {
  int chrom = 0;
  int64_t w0 = 0, w1 = 0;                 // windows [w0, w1) of the chromosome
  std::vector<RowRange> ranges;           // chromosome rows the windows touch; `res` = position in the working panel
  int64_t n_rows = 0;
  gb_batch* batch = nullptr;
  int64_t out_off = 0;                    // u_off[w0]: position of the results in the chromosome's arrays
  int64_t stage_off = 0;                  // position in the shard's pinned result staging
  size_t status_off = 0;
  cudaEvent_t landed = nullptr;           // the rows this segment touches are resident
  int slot = 0;                           // which compute stream / working panel / arena
  bool expanded = false;                  // its rows stay resident in the E2M1 operand layout (no expansion per run)
};

struct Shard {
  int index = 0;                          // position in the genome's GPU list
  int part = 0;                           // partition index this shard computes
  gb_ctx* ctx = nullptr;
  std::vector<Segment> segs;
  // Residency is decided per batch: as many batches as HBM allows keep their rows EXPANDED (E2M1 operand rows, 16.7 KB
  // per SNP of the 33KG shape, read by TMA directly); the others keep them as ternary rows (6.5 KB) and expand them into
  // a working panel right before they run.  A whole genome on one GPU is ~70 % / 30 %; from two GPUs on all is expanded.
  std::vector<std::vector<RowRange>> resident;     // per chromosome: every row range this shard needs (what a feeder supplies)
  std::vector<std::vector<RowRange>> res_e, res_t; // ... split by form: expanded (res = row in `big`) / ternary (res = row in d_rows5)
  int64_t resident_rows = 0, rows_e = 0, rows_t = 0;
  bool e2m1_resident = false;             // every batch is expanded
  uint8_t* d_rows5 = nullptr;             // ternary residency: [rows_t][row5]
  int64_t* d_sites = nullptr;             // synthetic fill: site index per needed row (indexed like `resident`)
  gb_panel* big = nullptr;                // expanded residency: [rows_e] operand rows
  gb_panel* panels[4] = {nullptr, nullptr, nullptr, nullptr};   // working panels of the ternary batches
  Arena arenas[4];
  cudaStream_t cs[2] = {nullptr, nullptr}, sides[2] = {nullptr, nullptr}, copy = nullptr;
  cudaEvent_t ev_lane[4] = {nullptr, nullptr, nullptr, nullptr}, ev_chain[4] = {nullptr, nullptr, nullptr, nullptr};   // pipelined run: B11 done / chain done, per arena slot
  cudaEvent_t ev_start = nullptr, ev_end = nullptr, ev_tmp = nullptr;
  double* h_z = nullptr;                  // pinned staging [n_u of the shard]
  double* h_info = nullptr;
  int* h_status = nullptr;
  int64_t n_u_total = 0;
  size_t status_total = 0;
  uint8_t* d_chunk[2] = {nullptr, nullptr};   // E2M1 residency: ternary staging chunks of the one-time expansion
  int64_t chunk_rows = 0;
  // work statistics
  double cost = 0, gram_ops = 0, solve_flops = 0;
  int64_t n_windows = 0, n_imputed = 0;
  double last_ms = 0, upload_ms = 0;
  // worker thread
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  int cmd = 0, cmd_seq = 0, done_seq = 0, rc = GB_OK;
  std::string err;
};

enum { CMD_NONE = 0, CMD_PLAN, CMD_UPLOAD, CMD_FILL, CMD_RUN, CMD_EXIT };

}  // namespace

struct gb_genome {
  int n_gpus = 0;
  std::vector<int> devices;
  int n_pops = 0;
  std::vector<int> pop_sizes;
  std::vector<double> pop_wgt;
  bool mix = false;
  gb_params params{};
  int row5 = 0;
  int64_t n_samples = 0;
  std::vector<Chrom> chroms;
  std::vector<Shard*> shards;
  bool planned = false, rows_ready = false;
  int n_parts = 1, first_part = 0;
  int64_t batch_windows = 48;
  int n_streams = 2;
  int lane_depth = 1;          // pipelined run: the solve of a batch is issued this many batches behind its front (GB_GENOME_LANE_DEPTH, 1..3)
  int chain_sms = 32;          // > 0: batches are software-pipelined over one heavy lane and one factorisation lane that keeps this
                               // many SMs (GB_GENOME_CHAIN_SMS; 0: every batch forks its own factorisation, two batches alternate)
  int resident_mode = 0;       // 0 auto, 1 pack5, 2 e2m1
  double expanded_gb = -1.0;   // >= 0: cap on the bytes spent on keeping batches expanded (GB_GENOME_EXPANDED_GB)
  uint64_t synth_seed = 0;
  // run arguments
  double* const* out_z = nullptr;
  double* const* out_info = nullptr;
  int* const* out_status = nullptr;
  bool wait_rows = false;      // the run follows an asynchronous upload: batches wait for their rows
  std::string err;
  std::vector<int64_t> part_cuts;          // global window cuts of the partition [n_parts + 1]
  std::vector<int64_t> chrom_w0;           // first global window id of each chromosome
};

namespace {

#define SH_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      sh->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
      return e__ == cudaErrorMemoryAllocation ? GB_ERR_OOM : GB_ERR_CUDA;                   \
    }                                                                                       \
  } while (0)

// rows of chromosome c touched by windows [w0, w1): the measured rows (with the wings) and the unmeasured rows, as
// merged ranges
std::vector<RowRange> rows_touched(const Chrom& c, int64_t w0, int64_t w1) {
  std::vector<RowRange> r;
  auto add = [&](const std::vector<int64_t>& rows, int64_t a, int64_t b) {
    if (b <= a) return;
    int64_t lo = rows[(size_t)a], hi = rows[(size_t)a];
    for (int64_t i = a; i < b; i++) {
      lo = std::min(lo, rows[(size_t)i]);
      hi = std::max(hi, rows[(size_t)i]);
    }
    RowRange x;
    x.lo = lo;
    x.hi = hi + 1;
    r.push_back(x);
  };
  // per window, so that two far-apart runs (measured block / unmeasured block of the packed layout) stay two ranges
  for (int64_t w = w0; w < w1; w++) {
    add(c.rows_t, c.t_off[(size_t)w], c.t_off[(size_t)w + 1]);
    add(c.rows_u, c.u_off[(size_t)w], c.u_off[(size_t)w + 1]);
  }
  std::sort(r.begin(), r.end(), [](const RowRange& a, const RowRange& b) { return a.lo < b.lo; });
  std::vector<RowRange> m;
  for (const RowRange& x : r) {
    if (!m.empty() && x.lo <= m.back().hi + 256) m.back().hi = std::max(m.back().hi, x.hi);   // small gaps are cheaper kept
    else m.push_back(x);
  }
  return m;
}

int64_t map_row(const std::vector<RowRange>& ranges, int64_t row) {
  for (const RowRange& x : ranges)
    if (row >= x.lo && row < x.hi) return x.res + (row - x.lo);
  return -1;
}

int shard_free(gb_genome* g, Shard* sh) {
  (void)g;
  if (!sh->ctx) return GB_OK;
  cudaSetDevice(sh->ctx->device);
  for (int i = 0; i < 2; i++) {
    if (sh->cs[i]) cudaStreamSynchronize(sh->cs[i]);
    if (sh->sides[i]) cudaStreamSynchronize(sh->sides[i]);
  }
  if (sh->copy) cudaStreamSynchronize(sh->copy);
  cudaStreamSynchronize(sh->ctx->stream);
  for (Segment& s : sh->segs) {
    if (s.batch) {
      batch_free_device(s.batch);
      delete s.batch;
    }
    if (s.landed) cudaEventDestroy(s.landed);
  }
  sh->segs.clear();
  cudaStreamSynchronize(sh->ctx->stream);
  if (sh->big) gb_panel_destroy(sh->big);
  sh->big = nullptr;
  for (int i = 0; i < 4; i++) {
    if (sh->panels[i]) gb_panel_destroy(sh->panels[i]);
    if (sh->arenas[i].base) cudaFree(sh->arenas[i].base);
    sh->panels[i] = nullptr;
    sh->arenas[i] = Arena{};
  }
  for (int i = 0; i < 2; i++) {
    if (sh->d_chunk[i]) cudaFree(sh->d_chunk[i]);
    sh->d_chunk[i] = nullptr;
  }
  if (sh->d_rows5) cudaFree(sh->d_rows5);
  if (sh->d_sites) cudaFree(sh->d_sites);
  if (sh->h_z) cudaFreeHost(sh->h_z);
  if (sh->h_info) cudaFreeHost(sh->h_info);
  if (sh->h_status) cudaFreeHost(sh->h_status);
  sh->d_rows5 = nullptr;
  sh->d_sites = nullptr;
  sh->h_z = sh->h_info = nullptr;
  sh->h_status = nullptr;
  for (int i = 0; i < 2; i++) {
    if (sh->cs[i]) cudaStreamDestroy(sh->cs[i]);
    if (sh->sides[i]) cudaStreamDestroy(sh->sides[i]);
    sh->cs[i] = sh->sides[i] = nullptr;
  }
  if (sh->copy) cudaStreamDestroy(sh->copy);
  sh->copy = nullptr;
  for (cudaEvent_t* e : {&sh->ev_start, &sh->ev_end, &sh->ev_tmp, &sh->ev_lane[0], &sh->ev_lane[1], &sh->ev_lane[2], &sh->ev_lane[3],
                         &sh->ev_chain[0], &sh->ev_chain[1], &sh->ev_chain[2], &sh->ev_chain[3]})
    if (*e) {
      cudaEventDestroy(*e);
      *e = nullptr;
    }
  return GB_OK;
}

// ---- plan: segments, residency, working panels, arenas, batches (runs on the shard's own thread) --------------------
// arenas / working panels a shard rotates its batches over: D + 1 in the pipelined run, else one per compute stream
static int slots_of(const gb_genome* g, const Shard* sh) {
  return (sh->ctx->heavy_sms > 0 && g->n_streams == 2) ? g->lane_depth + 1 : g->n_streams;
}

int shard_plan(gb_genome* g, Shard* sh) {
  Ctx* ctx = sh->ctx;
  SH_CUDA(cudaSetDevice(ctx->device));
  const int64_t lo = g->part_cuts[(size_t)sh->part], hi = g->part_cuts[(size_t)sh->part + 1];
  sh->resident.assign(g->chroms.size(), {});
  sh->segs.clear();
  sh->cost = sh->gram_ops = sh->solve_flops = 0;
  sh->n_windows = hi - lo;
  sh->n_imputed = 0;
  // 1. segments: per chromosome piece, ~batch_windows windows each, cut by cost
  for (size_t c = 0; c < g->chroms.size(); c++) {
    const Chrom& ch = g->chroms[c];
    const int64_t a = std::max<int64_t>(lo - g->chrom_w0[c], 0), b = std::min<int64_t>(hi - g->chrom_w0[c], ch.n_windows);
    if (b <= a) continue;
    const int64_t n = b - a;
    const int n_seg = (int)std::max<int64_t>(1, (n + g->batch_windows - 1) / g->batch_windows);
    std::vector<double> cost((size_t)n);
    for (int64_t w = 0; w < n; w++) {
      const double nt = (double)(ch.t_off[(size_t)(a + w) + 1] - ch.t_off[(size_t)(a + w)]);
      const double nu = (double)(ch.u_off[(size_t)(a + w) + 1] - ch.u_off[(size_t)(a + w)]);
      cost[(size_t)w] = window_cost(nt, nu, (double)g->n_samples, g->params) + 1.0;
      sh->cost += cost[(size_t)w];
      if (cost[(size_t)w] > 1.0) sh->n_imputed += (int64_t)nu;
    }
    std::vector<int64_t> cuts((size_t)n_seg + 1);
    partition_contiguous(cost.data(), n, n_seg, cuts.data());
    for (int s = 0; s < n_seg; s++) {
      if (cuts[(size_t)s + 1] <= cuts[(size_t)s]) continue;
      Segment sg;
      sg.chrom = (int)c;
      sg.w0 = a + cuts[(size_t)s];
      sg.w1 = a + cuts[(size_t)s + 1];
      sg.ranges = rows_touched(ch, sg.w0, sg.w1);
      sg.out_off = ch.u_off[(size_t)sg.w0];
      sh->segs.push_back(std::move(sg));
    }
  }
  // 2. every row range this shard needs, per chromosome (what a feeder has to supply)
  auto merge_ranges = [&](bool want_e, bool want_t, std::vector<std::vector<RowRange>>& out, int64_t& total) {
    out.assign(g->chroms.size(), {});
    total = 0;
    for (size_t c = 0; c < g->chroms.size(); c++) {
      std::vector<RowRange> all;
      for (const Segment& s : sh->segs)
        if (s.chrom == (int)c && ((s.expanded && want_e) || (!s.expanded && want_t))) all.insert(all.end(), s.ranges.begin(), s.ranges.end());
      std::sort(all.begin(), all.end(), [](const RowRange& a, const RowRange& b) { return a.lo < b.lo; });
      std::vector<RowRange>& m = out[c];
      for (const RowRange& x : all) {
        if (!m.empty() && x.lo <= m.back().hi) m.back().hi = std::max(m.back().hi, x.hi);
        else m.push_back(x);
      }
      for (RowRange& x : m) {
        x.res = total;
        x.issued = x.lo;
        total += x.hi - x.lo;
      }
    }
  };
  merge_ranges(true, true, sh->resident, sh->resident_rows);
  int64_t max_seg_rows = 1;
  for (Segment& s : sh->segs) {
    int64_t n = 0;
    for (const RowRange& x : s.ranges) n += x.hi - x.lo;
    s.n_rows = n;
    max_seg_rows = std::max(max_seg_rows, n);
  }
  if (sh->resident_rows > (int64_t)std::numeric_limits<int32_t>::max() - 512) {
    sh->err = "shard holds more than 2^31 panel rows";
    return GB_ERR_UNSUPPORTED;
  }
  // 3. streams, events
  int prio_lo = 0, prio_hi = 0;
  SH_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  for (int i = 0; i < 2; i++) {
    SH_CUDA(cudaStreamCreateWithFlags(&sh->cs[i], cudaStreamNonBlocking));
    // the factorisation lanes: many small dependent launches, scheduled ahead of the heavy lane's CTAs wherever an SM frees up
    SH_CUDA(cudaStreamCreateWithPriority(&sh->sides[i], cudaStreamNonBlocking, prio_hi));
  }
  SH_CUDA(cudaStreamCreateWithFlags(&sh->copy, cudaStreamNonBlocking));
  SH_CUDA(cudaEventCreate(&sh->ev_start));
  SH_CUDA(cudaEventCreate(&sh->ev_end));
  SH_CUDA(cudaEventCreateWithFlags(&sh->ev_tmp, cudaEventDisableTiming));
  for (int i = 0; i < 4; i++) {
    SH_CUDA(cudaEventCreateWithFlags(&sh->ev_lane[i], cudaEventDisableTiming));
    SH_CUDA(cudaEventCreateWithFlags(&sh->ev_chain[i], cudaEventDisableTiming));
  }
  const int n_slots = slots_of(g, sh);
  // 4. residency per batch.  Workspace need of a batch from its window sizes (what batch_arena_bytes will report):
  size_t arena_est = 0;
  for (const Segment& s : sh->segs) {
    const Chrom& ch = g->chroms[(size_t)s.chrom];
    double tt = 0, ut = 0, dinv = 0, ntt = 0, nut = 0;
    for (int64_t w = s.w0; w < s.w1; w++) {
      const double nt = (double)(ch.t_off[(size_t)w + 1] - ch.t_off[(size_t)w]), nu = (double)(ch.u_off[(size_t)w + 1] - ch.u_off[(size_t)w]);
      tt += nt * (nt + 8);
      ut += nt * (nu + 8);
      dinv += (nt / 64 + 1) * 4096;
      ntt += nt;
      nut += nu;
    }
    const double bytes = 8 * (2 * tt + ut + 2 * dinv + 2 * nut) + 12 * nut + 12.0 * g->n_pops * (ntt + nut) + (1 << 20);
    arena_est = std::max(arena_est, (size_t)bytes);
  }
  size_t free_b = 0, total_b = 0;
  SH_CUDA(cudaMemGetInfo(&free_b, &total_b));
  int64_t k_elems = 0;   // E2M1 row: population blocks on 128-column boundaries, two dosages per byte
  for (int p = 0; p < g->n_pops; p++) k_elems += (g->pop_sizes[(size_t)p] + K_BLOCK - 1) / K_BLOCK * K_BLOCK;
  const double row_e = (double)(k_elems / 2 + 8 * g->n_pops), row_t = (double)g->row5;
  {
    // start with everything ternary (+ the two working panels), then keep batches expanded while the budget lasts
    const double base = (double)sh->resident_rows * row_t + (double)n_slots * (double)max_seg_rows * row_e;
    double extra_budget = (double)free_b * 0.93 - (double)n_slots * (double)arena_est * 1.05 - 4e9 - base;
    if (g->expanded_gb >= 0.0) extra_budget = g->expanded_gb * 1e9;   // GB_GENOME_EXPANDED_GB (tuning / tests): cap on the extra bytes
    const bool force_t = g->resident_mode == 1, force_e = g->resident_mode == 2;
    double used = 0.0;
    for (Segment& s : sh->segs) {
      const double extra = (double)s.n_rows * (row_e - row_t);
      s.expanded = force_e || (!force_t && used + extra <= extra_budget);
      if (s.expanded) used += extra;
    }
  }
  merge_ranges(true, false, sh->res_e, sh->rows_e);
  merge_ranges(false, true, sh->res_t, sh->rows_t);
  sh->e2m1_resident = sh->rows_t == 0;
  if (sh->rows_e > 0) {
    int rc = gb_panel_create_fmt(sh->ctx, g->n_pops, g->pop_sizes.data(), sh->rows_e, GB_PANEL_E2M1, &sh->big);
    if (rc) {
      sh->err = "resident E2M1 panel: " + ctx->err;
      return rc;
    }
    sh->big->n_rows = sh->rows_e;
    sh->chunk_rows = std::min<int64_t>(sh->rows_e, 65536);
    for (int i = 0; i < 2; i++) SH_CUDA(cudaMalloc(reinterpret_cast<void**>(&sh->d_chunk[i]), (size_t)sh->chunk_rows * (size_t)g->row5));
  }
  {
    // the working panels double as the holder of the population tables the synthetic generator needs
    int64_t max_t_rows = 1;
    for (const Segment& s : sh->segs)
      if (!s.expanded) max_t_rows = std::max(max_t_rows, s.n_rows);
    const int n_work = sh->rows_t > 0 ? n_slots : 1;
    for (int i = 0; i < n_work; i++) {
      int rc = gb_panel_create_fmt(sh->ctx, g->n_pops, g->pop_sizes.data(), sh->rows_t > 0 ? max_t_rows : 1, GB_PANEL_E2M1, &sh->panels[i]);
      if (rc) {
        sh->err = "working panel: " + ctx->err;
        return rc;
      }
      sh->panels[i]->n_rows = sh->panels[i]->capacity;
    }
    if (sh->rows_t > 0) SH_CUDA(cudaMalloc(reinterpret_cast<void**>(&sh->d_rows5), (size_t)sh->rows_t * (size_t)g->row5));
  }
  // 5. batches: phase 1 for all (sizes), arenas, phase 2
  size_t arena_need = 0;
  sh->n_u_total = 0;
  sh->status_total = 0;
  for (size_t si = 0; si < sh->segs.size(); si++) {
    Segment& s = sh->segs[si];
    const Chrom& ch = g->chroms[(size_t)s.chrom];
    s.slot = (int)(si % (size_t)n_slots);
    gb_panel* panel = s.expanded ? sh->big : sh->panels[s.slot];
    // position of the segment's rows: in the resident expanded panel, or in its working panel
    if (s.expanded) {
      for (RowRange& x : s.ranges) x.res = map_row(sh->res_e[(size_t)s.chrom], x.lo);
    } else {
      int64_t off = 0;
      for (RowRange& x : s.ranges) {
        x.res = off;
        off += x.hi - x.lo;
      }
    }
    const int64_t nw = s.w1 - s.w0;
    std::vector<int64_t> to((size_t)nw + 1), uo((size_t)nw + 1);
    for (int64_t i = 0; i <= nw; i++) {
      to[(size_t)i] = ch.t_off[(size_t)(s.w0 + i)] - ch.t_off[(size_t)s.w0];
      uo[(size_t)i] = ch.u_off[(size_t)(s.w0 + i)] - ch.u_off[(size_t)s.w0];
    }
    std::vector<int64_t> rt((size_t)to[(size_t)nw]), ru((size_t)uo[(size_t)nw]);
    const int64_t t0 = ch.t_off[(size_t)s.w0], u0 = ch.u_off[(size_t)s.w0];
    for (size_t i = 0; i < rt.size(); i++) rt[i] = map_row(s.ranges, ch.rows_t[(size_t)t0 + i]);
    for (size_t i = 0; i < ru.size(); i++) ru[i] = map_row(s.ranges, ch.rows_u[(size_t)u0 + i]);
    s.batch = batch_new(sh->ctx, panel, nw, g->mix ? g->pop_wgt.data() : nullptr, &g->params, false, false,
                        /*defer_flag_check=*/true, 1.0);
    if (!s.batch) return GB_ERR_OOM;
    int rc = batch_plan_host(s.batch, to.data(), rt.data(), uo.data(), ru.data(), ch.z_t.data() + t0,
                             g->mix ? g->pop_wgt.data() : nullptr);
    if (rc) {
      sh->err = ctx->err;
      return rc;
    }
    arena_need = std::max(arena_need, batch_arena_bytes(s.batch));
    s.stage_off = sh->n_u_total;
    sh->n_u_total += uo[(size_t)nw];
    s.status_off = sh->status_total;
    sh->status_total += 2 * (size_t)nw + 3;
    sh->gram_ops += s.batch->work_gram_ops;
    sh->solve_flops += s.batch->work_solve_flops;
    SH_CUDA(cudaEventCreateWithFlags(&s.landed, cudaEventDisableTiming));
  }
  for (int i = 0; i < n_slots; i++) {
    SH_CUDA(cudaMalloc(reinterpret_cast<void**>(&sh->arenas[i].base), std::max<size_t>(arena_need, 256)));
    sh->arenas[i].cap = arena_need;
  }
  for (Segment& s : sh->segs) {
    int rc = batch_plan_device(s.batch, sh->arenas[s.slot], /*sync=*/false);
    if (rc) {
      sh->err = ctx->err;
      return rc;
    }
  }
  SH_CUDA(cudaStreamSynchronize(ctx->stream));
  SH_CUDA(cudaMallocHost(reinterpret_cast<void**>(&sh->h_z), sizeof(double) * (size_t)std::max<int64_t>(sh->n_u_total, 1)));
  SH_CUDA(cudaMallocHost(reinterpret_cast<void**>(&sh->h_info), sizeof(double) * (size_t)std::max<int64_t>(sh->n_u_total, 1)));
  SH_CUDA(cudaMallocHost(reinterpret_cast<void**>(&sh->h_status), sizeof(int) * std::max<size_t>(sh->status_total, 1)));
  return GB_OK;
}

// Rows [a, b) of a chromosome from the caller's HOST pieces to `dst` (device, row5 bytes apart) on the copy stream.
int copy_host_rows(gb_genome* g, Shard* sh, const Chrom& ch, int64_t a, int64_t b, uint8_t* dst) {
  while (a < b) {
    const HostPiece* pc = nullptr;
    for (auto it = ch.pieces.rbegin(); it != ch.pieces.rend(); ++it)   // newest piece first
      if (a >= it->lo && a < it->hi) {
        pc = &*it;
        break;
      }
    if (!pc) {
      sh->ctx->err = "no host rows were given for panel row " + std::to_string(a);
      return GB_ERR_BAD_ARG;
    }
    const int64_t n = std::min(b, pc->hi) - a;
    const uint8_t* src = pc->ptr + (size_t)(a - pc->lo) * (size_t)pc->stride;
    cudaError_t e = pc->stride == g->row5
                        ? cudaMemcpyAsync(dst, src, (size_t)n * (size_t)g->row5, cudaMemcpyHostToDevice, sh->copy)
                        : cudaMemcpy2DAsync(dst, (size_t)g->row5, src, (size_t)pc->stride, (size_t)g->row5, (size_t)n,
                                            cudaMemcpyHostToDevice, sh->copy);
    if (e != cudaSuccess) {
      sh->ctx->err = std::string("host -> device copy of panel rows: ") + cudaGetErrorString(e);
      return GB_ERR_CUDA;
    }
    dst += (size_t)n * (size_t)g->row5;
    a += n;
  }
  return GB_OK;
}

// ---- rows: upload from the host, or synthetic fill, segment by segment --------------------------------------------
// Issues, on the copy stream, whatever is still missing of the rows segment s touches and records s.landed behind it.
// E2M1 residency: the ternary rows pass through two staging chunks and are expanded into the resident panel.
int shard_rows_for_segment(gb_genome* g, Shard* sh, Segment& s, bool synthetic) {
  Ctx* ctx = sh->ctx;
  const Chrom& ch = g->chroms[(size_t)s.chrom];
  cudaStream_t keep = ctx->stream;
  ctx->stream = sh->copy;
  int rc = GB_OK;
  static thread_local int chunk_flip = 0;
  std::vector<RowRange>& have = s.expanded ? sh->res_e[(size_t)s.chrom] : sh->res_t[(size_t)s.chrom];
  for (const RowRange& need : s.ranges) {
    for (RowRange& res : have) {
      if (need.lo < res.lo || need.lo >= res.hi) continue;
      const int64_t a = res.issued, b = std::max(res.issued, need.hi);
      if (b <= a) break;
      const int64_t pos = res.res + (a - res.lo);      // position in the ternary buffer / the expanded panel
      // site index of the rows (synthetic fill): d_sites is laid out like `resident`
      const int64_t spos = ch.sites.empty() ? 0 : map_row(sh->resident[(size_t)s.chrom], a);
      if (!s.expanded) {
        uint8_t* dst = sh->d_rows5 + (size_t)pos * (size_t)g->row5;
        if (synthetic) {
          rc = launch_synth_pack5(ctx, dst, g->row5, b - a, ch.sites.empty() ? nullptr : sh->d_sites + spos, a, g->n_pops,
                                  sh->panels[0]->d_pop_sizes, sh->panels[0]->d_boff5, g->row5, g->synth_seed, s.chrom);
        } else {
          rc = copy_host_rows(g, sh, ch, a, b, dst);
        }
      } else {
        // ternary rows pass through a staging chunk and are expanded into the resident panel once
        for (int64_t r0 = a; r0 < b && !rc; r0 += sh->chunk_rows) {
          const int64_t n = std::min(sh->chunk_rows, b - r0);
          uint8_t* stg = sh->d_chunk[chunk_flip & 1];
          chunk_flip++;
          // (same stream: the expansion that last read this staging chunk is ordered before the copy that refills it)
          if (synthetic) {
            rc = launch_synth_pack5(ctx, stg, g->row5, n, ch.sites.empty() ? nullptr : sh->d_sites + spos + (r0 - a), r0,
                                    g->n_pops, sh->panels[0]->d_pop_sizes, sh->panels[0]->d_boff5, g->row5, g->synth_seed, s.chrom);
          } else {
            rc = copy_host_rows(g, sh, ch, r0, r0 + n, stg);
          }
          if (!rc) rc = launch_expand5(ctx, sh->big, stg, g->row5, pos + (r0 - a), n);
        }
      }
      res.issued = b;
      break;
    }
    if (rc) break;
  }
  ctx->stream = keep;
  if (rc) {
    sh->err = ctx->err.empty() ? "row upload failed" : ctx->err;
    return rc;
  }
  SH_CUDA(cudaEventRecord(s.landed, sh->copy));
  return GB_OK;
}

int shard_rows(gb_genome* g, Shard* sh, bool synthetic) {
  Ctx* ctx = sh->ctx;
  SH_CUDA(cudaSetDevice(ctx->device));
  for (auto* set : {&sh->res_e, &sh->res_t})
    for (auto& v : *set)
      for (RowRange& x : v) x.issued = x.lo;
  if (sh->big) gb_panel_clear(sh->big), sh->big->n_rows = sh->rows_e;
  SH_CUDA(cudaStreamSynchronize(ctx->stream));
  if (synthetic) {
    // site index of every resident row (rows of a chromosome are not in bp order: measured block | unmeasured block)
    bool any = false;
    for (const Chrom& c : g->chroms) any |= !c.sites.empty();
    if (any && !sh->d_sites) {
      SH_CUDA(cudaMalloc(reinterpret_cast<void**>(&sh->d_sites), sizeof(int64_t) * (size_t)std::max<int64_t>(sh->resident_rows, 1)));
      for (size_t c = 0; c < g->chroms.size(); c++)
        for (const RowRange& x : sh->resident[c])
          if (!g->chroms[c].sites.empty())
            SH_CUDA(cudaMemcpyAsync(sh->d_sites + x.res, g->chroms[c].sites.data() + x.lo, sizeof(int64_t) * (size_t)(x.hi - x.lo),
                                    cudaMemcpyHostToDevice, sh->copy));
    }
  }
  SH_CUDA(cudaEventRecord(sh->ev_start, sh->copy));
  for (Segment& s : sh->segs) {
    int rc = shard_rows_for_segment(g, sh, s, synthetic);
    if (rc) return rc;
  }
  SH_CUDA(cudaEventRecord(sh->ev_end, sh->copy));
  return GB_OK;
}

// ---- run: every batch of the shard, alternating between the compute streams ----------------------------------------
int shard_run(gb_genome* g, Shard* sh) {
  Ctx* ctx = sh->ctx;
  SH_CUDA(cudaSetDevice(ctx->device));
  const int n_slots = slots_of(g, sh);
  cudaStream_t main_stream = ctx->stream, main_side = ctx->side_stream;
  cudaEvent_t ev_t0, ev_t1;
  SH_CUDA(cudaEventCreate(&ev_t0));
  SH_CUDA(cudaEventCreate(&ev_t1));
  SH_CUDA(cudaEventRecord(ev_t0, sh->cs[0]));
  for (int i = 1; i < 2; i++) SH_CUDA(cudaStreamWaitEvent(sh->cs[i], ev_t0, 0));
  int rc = GB_OK;
  if (ctx->heavy_sms > 0 && g->n_streams == 2) {
    // Software pipeline over two lanes.  The heavy lane (one stream: the Gram kernels, their finish passes, the solve GEMM;
    // its persistent kernels take heavy_sms CTAs, one per SM) runs, with a lag of D = lane_depth batches,
    //   P_0 Q_0 | P_1 Q_1 | P_2 S_0 Q_2 | P_3 S_1 Q_3 | ...        (P = row statistics + B11 tiles, Q = B21 tiles, S = solve)
    // and the factorisation lane (the other stream, high priority, on the SMs the heavy lane leaves) runs the Cholesky +
    // L^-1 chain of batch i between P_i and S_i: ~100 small dependent launches whose latency is set by the largest window
    // of the batch get D heavy periods to finish (with D = 1 the solve waited ~1 ms per batch for them).
    // Batches rotate over D + 1 arenas; stream order on the heavy lane (S_{i-D} before P_{i+1}) keeps them apart.
    cudaStream_t H = sh->cs[0], Cs = sh->sides[0];
    const int heavy = ctx->heavy_sms;
    const size_t lag = (size_t)(n_slots - 1);
    // diagnostics (timing only, results meaningless): which lane sets the period?
    const char* skip_env = getenv("GB_GENOME_SKIP");
    const bool skip_chain = skip_env && !strcmp(skip_env, "chain"), skip_heavy = skip_env && !strcmp(skip_env, "heavy");
    auto solve_of = [&](Segment& p) -> int {
      SH_CUDA(cudaStreamWaitEvent(H, sh->ev_chain[p.slot], 0));
      int r = skip_heavy ? GB_OK : batch_run_solve(p.batch);
      if (!r) r = batch_fetch_enqueue(p.batch, sh->h_z + p.stage_off, sh->h_info + p.stage_off, sh->h_status + p.status_off);
      return r;
    };
    size_t next_solve = 0;
    for (size_t si = 0; si < sh->segs.size() && !rc; si++) {
      Segment& s = sh->segs[si];
      SH_CUDA(cudaStreamWaitEvent(H, s.landed, 0));
      ctx->stream = H;
      if (!s.expanded) {
        gb_panel* panel = sh->panels[s.slot];
        for (const RowRange& x : s.ranges) {
          const int64_t pos = map_row(sh->res_t[(size_t)s.chrom], x.lo);
          rc = launch_expand5(ctx, panel, sh->d_rows5 + (size_t)pos * (size_t)g->row5, g->row5, x.res, x.hi - x.lo);
          if (rc) break;
        }
      }
      if (!rc) rc = batch_run_front(s.batch, heavy);
      if (rc) break;
      SH_CUDA(cudaEventRecord(sh->ev_lane[s.slot], H));
      SH_CUDA(cudaStreamWaitEvent(Cs, sh->ev_lane[s.slot], 0));
      ctx->stream = Cs;
      if (!skip_chain) rc = batch_run_chain(s.batch);
      ctx->stream = H;
      if (rc) break;
      SH_CUDA(cudaEventRecord(sh->ev_chain[s.slot], Cs));
      if (si >= lag) {
        rc = solve_of(sh->segs[next_solve++]);
        if (rc) break;
      }
      if (!skip_heavy) rc = batch_run_b21(s.batch, heavy);
    }
    ctx->stream = H;
    while (!rc && next_solve < sh->segs.size()) rc = solve_of(sh->segs[next_solve++]);
    if (!rc) {   // the timing join below looks at cs[0] only; the factorisation lane has been joined by the last solve
      SH_CUDA(cudaEventRecord(sh->ev_tmp, Cs));
      SH_CUDA(cudaStreamWaitEvent(H, sh->ev_tmp, 0));
    }
  } else
  for (Segment& s : sh->segs) {
    cudaStream_t cs = sh->cs[s.slot];
    SH_CUDA(cudaStreamWaitEvent(cs, s.landed, 0));   // the rows this batch touches are resident (a completed event costs nothing)
    ctx->stream = cs;
    ctx->side_stream = sh->sides[s.slot];
    if (!s.expanded) {
      gb_panel* panel = sh->panels[s.slot];
      for (const RowRange& x : s.ranges) {
        const int64_t pos = map_row(sh->res_t[(size_t)s.chrom], x.lo);
        rc = launch_expand5(ctx, panel, sh->d_rows5 + (size_t)pos * (size_t)g->row5, g->row5, x.res, x.hi - x.lo);
        if (rc) break;
      }
    }
    if (!rc) rc = gb_batch_run(s.batch);
    if (!rc) rc = batch_fetch_enqueue(s.batch, sh->h_z + s.stage_off, sh->h_info + s.stage_off, sh->h_status + s.status_off);
    if (rc) break;
  }
  ctx->stream = main_stream;
  ctx->side_stream = main_side;
  if (rc) {
    sh->err = ctx->err;
    for (int i = 0; i < 2; i++) cudaStreamSynchronize(sh->cs[i]);
    cudaEventDestroy(ev_t0);
    cudaEventDestroy(ev_t1);
    return rc;
  }
  for (int i = 1; i < 2; i++) {
    SH_CUDA(cudaEventRecord(sh->ev_tmp, sh->cs[i]));
    SH_CUDA(cudaStreamWaitEvent(sh->cs[0], sh->ev_tmp, 0));
  }
  SH_CUDA(cudaEventRecord(ev_t1, sh->cs[0]));
  SH_CUDA(cudaEventSynchronize(ev_t1));
  float ms = 0;
  cudaEventElapsedTime(&ms, ev_t0, ev_t1);
  sh->last_ms = ms;
  cudaEventDestroy(ev_t0);
  cudaEventDestroy(ev_t1);
  // host gather: statuses interpreted, results copied into the caller's per-chromosome arrays
  int worst = GB_OK;
  for (Segment& s : sh->segs) {
    const int64_t n = s.batch->n_u_total;
    double* zs = sh->h_z + s.stage_off;
    double* is = sh->h_info + s.stage_off;
    int* wst = g->out_status && g->out_status[s.chrom] ? g->out_status[s.chrom] + s.w0 : nullptr;
    std::vector<int> tmp;
    if (!wst) {
      tmp.resize((size_t)(s.w1 - s.w0));
      wst = tmp.data();
    }
    rc = batch_fetch_finish(s.batch, sh->h_status + s.status_off, zs, is, wst);
    bool repair = false;
    for (int64_t w = 0; w < s.w1 - s.w0 && !rc; w++) repair |= wst[w] == GB_ERR_NOT_PD || wst[w] == GB_ERR_BREAKDOWN;
    if (repair) {
      // uncertified windows take the eigen-clip path (MakePosDef proper).  With ternary residency the working panel has
      // been reused by later batches meanwhile: expand this batch's rows again first.
      if (!s.expanded)
        for (const RowRange& x : s.ranges) {
          const int64_t pos = map_row(sh->res_t[(size_t)s.chrom], x.lo);
          rc = launch_expand5(ctx, sh->panels[s.slot], sh->d_rows5 + (size_t)pos * (size_t)g->row5, g->row5, x.res, x.hi - x.lo);
          if (rc) break;
        }
      if (!rc) rc = batch_fetch_finish_repair(s.batch, sh->h_status + s.status_off, zs, is, wst);
    }
    if (rc) {
      sh->err = ctx->err;
      return rc;
    }
    for (int64_t w = 0; w < s.w1 - s.w0; w++)
      if (wst[w] != GB_OK && worst == GB_OK) worst = wst[w];
    if (g->out_z && g->out_z[s.chrom]) std::memcpy(g->out_z[s.chrom] + s.out_off, zs, sizeof(double) * (size_t)n);
    if (g->out_info && g->out_info[s.chrom]) std::memcpy(g->out_info[s.chrom] + s.out_off, is, sizeof(double) * (size_t)n);
  }
  return g->out_status ? GB_OK : worst;
}

void shard_thread(gb_genome* g, Shard* sh) {
  int seen = 0;
  for (;;) {
    int cmd;
    {
      std::unique_lock<std::mutex> lk(sh->mu);
      sh->cv.wait(lk, [&] { return sh->cmd_seq != seen; });
      seen = sh->cmd_seq;
      cmd = sh->cmd;
    }
    int rc = GB_OK;
    sh->err.clear();
    switch (cmd) {
      case CMD_PLAN:
        nvtxRangePushA("gb:genome_plan");
        rc = shard_plan(g, sh);
        nvtxRangePop();
        break;
      case CMD_UPLOAD: {
        rc = shard_rows(g, sh, false);
        break;
      }
      case CMD_FILL: rc = shard_rows(g, sh, true); break;
      case CMD_RUN:
        nvtxRangePushA("gb:genome_run");
        rc = shard_run(g, sh);
        nvtxRangePop();
        break;
      default: break;
    }
    {
      std::lock_guard<std::mutex> lk(sh->mu);
      sh->rc = rc;
      sh->done_seq = seen;
    }
    sh->cv.notify_all();
    if (cmd == CMD_EXIT) return;
  }
}

void post(gb_genome* g, int cmd) {
  for (Shard* sh : g->shards) {
    {
      std::lock_guard<std::mutex> lk(sh->mu);
      sh->cmd = cmd;
      sh->cmd_seq++;
    }
    sh->cv.notify_all();
  }
}

int wait_all(gb_genome* g) {
  int rc = GB_OK;
  for (Shard* sh : g->shards) {
    std::unique_lock<std::mutex> lk(sh->mu);
    sh->cv.wait(lk, [&] { return sh->done_seq == sh->cmd_seq; });
    if (sh->rc != GB_OK && rc == GB_OK) {
      rc = sh->rc;
      g->err = "GPU " + std::to_string(sh->ctx ? sh->ctx->device : -1) + ": " + sh->err;
    }
  }
  return rc;
}

}  // namespace

// =====================================================================================================================
extern "C" {

int gb_genome_create(int n_gpus, const int* devices, int n_pops, const int* pop_sizes, const double* pop_wgt,
                     const gb_params* params, gb_genome** out) {
  if (!out || n_gpus < 1 || n_gpus > 64 || n_pops < 1 || n_pops > P_MAX || !pop_sizes) return GB_ERR_BAD_ARG;
  *out = nullptr;
  gb_genome* g = new (std::nothrow) gb_genome();
  if (!g) return GB_ERR_OOM;
  g->n_gpus = n_gpus;
  g->n_pops = n_pops;
  g->pop_sizes.assign(pop_sizes, pop_sizes + n_pops);
  for (int p = 0; p < n_pops; p++) g->n_samples += pop_sizes[p];
  if (pop_wgt) {
    g->pop_wgt.assign(pop_wgt, pop_wgt + n_pops);
    g->mix = true;
  }
  if (params) g->params = *params;
  else gb_params_default(&g->params);
  g->row5 = pack5_layout(n_pops, pop_sizes, nullptr);
  if (g->row5 < 0) {
    delete g;
    return GB_ERR_BAD_ARG;
  }
  if (const char* e = getenv("GB_GENOME_BATCH_WINDOWS")) g->batch_windows = std::max(1, atoi(e));
  if (const char* e = getenv("GB_GENOME_STREAMS")) g->n_streams = atoi(e) == 1 ? 1 : 2;
  if (const char* e = getenv("GB_GENOME_CHAIN_SMS")) g->chain_sms = std::max(0, atoi(e));
  if (g->n_streams < 2) g->chain_sms = 0;
  if (const char* e = getenv("GB_GENOME_LANE_DEPTH")) g->lane_depth = std::min(3, std::max(1, atoi(e)));
  if (const char* e = getenv("GB_GENOME_EXPANDED_GB")) g->expanded_gb = atof(e);
  if (const char* e = getenv("GB_GENOME_RESIDENT")) g->resident_mode = !strcmp(e, "pack5") ? 1 : !strcmp(e, "e2m1") ? 2 : 0;
  for (int i = 0; i < n_gpus; i++) {
    Shard* sh = new Shard();
    sh->index = i;
    g->shards.push_back(sh);
    g->devices.push_back(devices ? devices[i] : i);
    int rc = gb_ctx_create(g->devices.back(), &sh->ctx);
    if (!rc && g->chain_sms > 0 && sh->ctx->sm_count > 2 * g->chain_sms) sh->ctx->heavy_sms = sh->ctx->sm_count - g->chain_sms;
    if (rc) {
      g->err = gb_last_error(nullptr);
      for (Shard* s : g->shards) {
        if (s->ctx) gb_ctx_destroy(s->ctx);
        delete s;
      }
      delete g;
      return rc;
    }
  }
  try {
    for (Shard* sh : g->shards) sh->th = std::thread(shard_thread, g, sh);
  } catch (...) {   // nothing may be thrown across the C-ABI
    gb_genome_destroy(g);   // stops and joins the threads that did start
    return GB_ERR_OOM;
  }
  *out = g;
  return GB_OK;
}

void gb_genome_destroy(gb_genome* g) {
  if (!g) return;
  post(g, CMD_EXIT);
  for (Shard* sh : g->shards)
    if (sh->th.joinable()) sh->th.join();
  for (Shard* sh : g->shards) {
    shard_free(g, sh);
    if (sh->ctx) gb_ctx_destroy(sh->ctx);
    delete sh;
  }
  delete g;
}

const char* gb_genome_last_error(const gb_genome* g) { return g ? g->err.c_str() : "null genome"; }

int gb_genome_add_chromosome(gb_genome* g, int64_t n_rows, const void* host_rows5, int64_t row_stride, int64_t n_windows,
                             const int64_t* t_off, const int64_t* rows_t, const int64_t* u_off, const int64_t* rows_u,
                             const double* z_t, const int64_t* sites) {
  if (!g || n_rows < 0 || n_windows < 0 || !t_off || !u_off || (host_rows5 && row_stride < g->row5)) {
    if (g) g->err = "null or negative argument";
    return GB_ERR_BAD_ARG;
  }
  if (g->planned) {
    g->err = "chromosomes must be added before gb_genome_plan";
    return GB_ERR_BAD_ARG;
  }
  if (t_off[0] != 0 || u_off[0] != 0 || (t_off[n_windows] > 0 && (!rows_t || !z_t)) || (u_off[n_windows] > 0 && !rows_u)) {
    g->err = "bad window lists";
    return GB_ERR_BAD_ARG;
  }
  Chrom c;
  c.n_rows = n_rows;
  if (host_rows5 && n_rows > 0) c.pieces.push_back(HostPiece{0, n_rows, static_cast<const uint8_t*>(host_rows5), row_stride});
  c.n_windows = n_windows;
  c.t_off.assign(t_off, t_off + n_windows + 1);
  c.u_off.assign(u_off, u_off + n_windows + 1);
  c.rows_t.assign(rows_t, rows_t + t_off[n_windows]);
  c.rows_u.assign(rows_u, rows_u + u_off[n_windows]);
  c.z_t.assign(z_t, z_t + t_off[n_windows]);
  if (c.z_t.empty()) c.z_t.push_back(0.0);
  for (int64_t w = 0; w < n_windows; w++)
    if (t_off[w + 1] < t_off[w] || u_off[w + 1] < u_off[w]) {
      g->err = "window offsets must be non-decreasing";
      return GB_ERR_BAD_ARG;
    }
  for (int64_t r : c.rows_t)
    if (r < 0 || r >= n_rows) {
      g->err = "measured row index out of range";
      return GB_ERR_BAD_ARG;
    }
  for (int64_t r : c.rows_u)
    if (r < 0 || r >= n_rows) {
      g->err = "unmeasured row index out of range";
      return GB_ERR_BAD_ARG;
    }
  if (sites) c.sites.assign(sites, sites + n_rows);
  g->chroms.push_back(std::move(c));
  return (int)GB_OK;
}

int gb_genome_plan(gb_genome* g, int n_parts, int first_part) {
  if (!g || n_parts < g->n_gpus || first_part < 0 || first_part + g->n_gpus > n_parts) {
    if (g) g->err = "bad partition arguments";
    return GB_ERR_BAD_ARG;
  }
  if (g->planned) {
    g->err = "already planned";
    return GB_ERR_BAD_ARG;
  }
  g->n_parts = n_parts;
  g->first_part = first_part;
  // global window list in chromosome order, cut into n_parts contiguous runs of ~equal cost
  g->chrom_w0.clear();
  std::vector<double> cost;
  for (const Chrom& c : g->chroms) {
    g->chrom_w0.push_back((int64_t)cost.size());
    for (int64_t w = 0; w < c.n_windows; w++)
      cost.push_back(window_cost((double)(c.t_off[(size_t)w + 1] - c.t_off[(size_t)w]),
                                 (double)(c.u_off[(size_t)w + 1] - c.u_off[(size_t)w]), (double)g->n_samples, g->params) + 1.0);
  }
  g->part_cuts.assign((size_t)n_parts + 1, 0);
  partition_contiguous(cost.data(), (int64_t)cost.size(), n_parts, g->part_cuts.data());
  for (int i = 0; i < g->n_gpus; i++) g->shards[(size_t)i]->part = first_part + i;
  post(g, CMD_PLAN);
  int rc = wait_all(g);
  g->planned = rc == GB_OK;
  return rc;
}

int gb_genome_num_chromosomes(const gb_genome* g) { return g ? (int)g->chroms.size() : -1; }

int gb_genome_shard_info(const gb_genome* g, int gpu, int64_t* first_window, int64_t* n_windows, int64_t* resident_rows,
                         int64_t* n_batches, int64_t* n_imputed, int* e2m1_resident, double* gram_ops, double* solve_flops) {
  if (!g || !g->planned || gpu < 0 || gpu >= g->n_gpus) return GB_ERR_BAD_ARG;
  const Shard* sh = g->shards[(size_t)gpu];
  if (first_window) *first_window = g->part_cuts[(size_t)sh->part];
  if (n_windows) *n_windows = sh->n_windows;
  if (resident_rows) *resident_rows = sh->resident_rows;
  if (n_batches) *n_batches = (int64_t)sh->segs.size();
  if (n_imputed) *n_imputed = sh->n_imputed;
  if (e2m1_resident) {   // per cent of this GPU's panel rows kept in the expanded operand layout (100 = no expansion at run time)
    const int64_t tot = sh->rows_e + sh->rows_t;
    *e2m1_resident = tot ? (int)((100 * sh->rows_e + tot / 2) / tot) : 100;
    if (sh->rows_t > 0 && *e2m1_resident == 100) *e2m1_resident = 99;
  }
  if (gram_ops) *gram_ops = sh->gram_ops;
  if (solve_flops) *solve_flops = sh->solve_flops;
  return GB_OK;
}

static int rows_cmd(gb_genome* g, int cmd, int wait) {
  if (!g || !g->planned) {
    if (g) g->err = "gb_genome_plan has not run";
    return GB_ERR_BAD_ARG;
  }
  post(g, cmd);
  int rc = wait_all(g);   // the worker threads only ENQUEUE copies / kernels; this returns as soon as they have
  if (rc) return rc;
  g->rows_ready = true;
  g->wait_rows = true;    // the next run orders every batch behind the event of its rows
  if (wait) {
    for (Shard* sh : g->shards) {
      cudaSetDevice(sh->ctx->device);
      if (cudaStreamSynchronize(sh->copy) != cudaSuccess) {
        g->err = "row upload failed";
        return GB_ERR_CUDA;
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, sh->ev_start, sh->ev_end);
      sh->upload_ms = ms;
    }
  }
  return GB_OK;
}

int gb_genome_upload(gb_genome* g, int wait) { return rows_cmd(g, CMD_UPLOAD, wait); }

int gb_genome_set_host_rows(gb_genome* g, int chrom, int64_t row_lo, int64_t n_rows, const void* host_rows5, int64_t row_stride) {
  if (!g || chrom < 0 || chrom >= (int)g->chroms.size() || row_lo < 0 || n_rows < 0 || !host_rows5 || row_stride < g->row5 ||
      row_lo + n_rows > g->chroms[(size_t)chrom].n_rows) {
    if (g) g->err = "bad host row piece";
    return GB_ERR_BAD_ARG;
  }
  int rc0 = wait_all(g);
  if (rc0) return rc0;
  Chrom& c = g->chroms[(size_t)chrom];
  // pieces may overlap (two GPUs both need the wing between their shards): the newest piece holding a row is used; a
  // piece that is covered entirely by the new one is dropped
  c.pieces.erase(std::remove_if(c.pieces.begin(), c.pieces.end(),
                                [&](const HostPiece& x) { return x.lo >= row_lo && x.hi <= row_lo + n_rows; }),
                 c.pieces.end());
  if (n_rows > 0) c.pieces.push_back(HostPiece{row_lo, row_lo + n_rows, static_cast<const uint8_t*>(host_rows5), row_stride});
  return GB_OK;
}

// Rows [row_lo, row_lo + n_rows) of chromosome `chrom` back from GPU `gpu`'s resident ternary rows into HOST memory
// (GB_ERR_UNSUPPORTED when that GPU keeps its rows expanded).  Lets a caller that generated a panel on the device keep
// a host copy without generating it twice.
int gb_genome_download_rows(gb_genome* g, int gpu, int chrom, int64_t row_lo, int64_t n_rows, void* host_out, int64_t out_stride) {
  if (!g || !g->planned || !g->rows_ready || gpu < 0 || gpu >= g->n_gpus || chrom < 0 || chrom >= (int)g->chroms.size() ||
      n_rows < 0 || !host_out || out_stride < g->row5) {
    if (g) g->err = "bad download arguments";
    return GB_ERR_BAD_ARG;
  }
  int rc0 = wait_all(g);
  if (rc0) return rc0;
  Shard* sh = g->shards[(size_t)gpu];
  for (const RowRange& x : sh->res_t[(size_t)chrom])
    if (row_lo >= x.lo && row_lo + n_rows <= x.hi) {
      cudaSetDevice(sh->ctx->device);
      cudaStreamSynchronize(sh->copy);
      cudaError_t e = cudaMemcpy2D(host_out, (size_t)out_stride, sh->d_rows5 + (size_t)(x.res + row_lo - x.lo) * (size_t)g->row5,
                                   (size_t)g->row5, (size_t)g->row5, (size_t)n_rows, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) {
        g->err = std::string("device -> host copy of panel rows: ") + cudaGetErrorString(e);
        return GB_ERR_CUDA;
      }
      return GB_OK;
    }
  g->err = "the rows are not resident in ternary form on this GPU (expanded operand rows cannot be read back)";
  return GB_ERR_UNSUPPORTED;
}

int gb_genome_resident_ranges(const gb_genome* g, int gpu, int chrom, int max_ranges, int64_t* lo, int64_t* hi, int* n_ranges) {
  if (!g || !g->planned || gpu < 0 || gpu >= g->n_gpus || chrom < 0 || chrom >= (int)g->chroms.size() || !n_ranges) return GB_ERR_BAD_ARG;
  const std::vector<RowRange>& v = g->shards[(size_t)gpu]->resident[(size_t)chrom];
  *n_ranges = (int)v.size();
  for (int i = 0; i < (int)v.size() && i < max_ranges; i++) {
    if (lo) lo[i] = v[(size_t)i].lo;
    if (hi) hi[i] = v[(size_t)i].hi;
  }
  return GB_OK;
}

int gb_genome_fill_synthetic(gb_genome* g, uint64_t seed, int wait) {
  if (g) g->synth_seed = seed;
  return rows_cmd(g, CMD_FILL, wait);
}

int gb_genome_submit(gb_genome* g, double* const* z_u, double* const* info_u, int* const* window_status) {
  if (!g || !g->planned || !g->rows_ready) {
    if (g) g->err = "plan and upload / fill the genome before running it";
    return GB_ERR_BAD_ARG;
  }
  int rc0 = wait_all(g);   // a previous command must have finished (one command in flight per worker)
  if (rc0) return rc0;
  g->out_z = z_u;
  g->out_info = info_u;
  g->out_status = window_status;
  post(g, CMD_RUN);
  return GB_OK;
}

int gb_genome_wait(gb_genome* g, double* gpu_ms, double* upload_ms) {
  if (!g) return GB_ERR_BAD_ARG;
  int rc = wait_all(g);
  for (size_t i = 0; i < g->shards.size(); i++) {
    Shard* sh = g->shards[i];
    if (gpu_ms) gpu_ms[i] = sh->last_ms;
    if (upload_ms) {
      cudaSetDevice(sh->ctx->device);
      float ms = 0;
      if (cudaEventQuery(sh->ev_end) == cudaSuccess && cudaEventElapsedTime(&ms, sh->ev_start, sh->ev_end) == cudaSuccess) sh->upload_ms = ms;
      upload_ms[i] = sh->upload_ms;
    }
  }
  g->wait_rows = false;   // the rows are resident now: later runs need not wait for them
  return rc;
}

int gb_genome_run(gb_genome* g, double* const* z_u, double* const* info_u, int* const* window_status, double* gpu_ms) {
  int rc = gb_genome_submit(g, z_u, info_u, window_status);
  if (rc) return rc;
  return gb_genome_wait(g, gpu_ms, nullptr);
}

int64_t gb_genome_launch_count(const gb_genome* g) {
  if (!g) return 0;
  int64_t n = 0;
  for (const Shard* sh : g->shards) n += sh->ctx->launches;
  return n;
}

// The cost-balanced contiguous partition the genome driver uses, for callers that place the parts themselves (one
// process per GPU under torchrun: every rank computes the same cuts).  cuts has n_parts + 1 entries.
int gb_partition_windows(int64_t n_windows, const int64_t* n_t, const int64_t* n_u, int64_t n_samples,
                         const gb_params* params, int n_parts, int64_t* cuts, double* cost_out) {
  if (n_windows < 0 || n_parts < 1 || !cuts || (n_windows && (!n_t || !n_u))) return GB_ERR_BAD_ARG;
  gb_params p;
  if (params) p = *params;
  else gb_params_default(&p);
  std::vector<double> cost((size_t)n_windows);
  for (int64_t w = 0; w < n_windows; w++) {
    cost[(size_t)w] = window_cost((double)n_t[w], (double)n_u[w], (double)n_samples, p) + 1.0;
    if (cost_out) cost_out[w] = cost[(size_t)w];
  }
  partition_contiguous(cost.data(), n_windows, n_parts, cuts);
  return GB_OK;
}

}  // extern "C"
