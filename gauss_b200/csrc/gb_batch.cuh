// gb_batch.cuh -- internal definitions shared by gb_api.cu (C-ABI, batch engine) and gb_genome.cu (multi-GPU genome
// driver): the opaque handle types and the two-phase batch planner.  Not part of the C-ABI.
#pragma once
#include <cstddef>
#include <string>
#include <vector>

#include "gb_common.cuh"

struct gb_ctx : public gb::Ctx {};
struct gb_panel : public gb::Panel {};
struct gb_pipe;

namespace gb {

// Device workspace several batches take turns in (the genome driver keeps two of them per GPU: one per compute
// stream).  Everything a batch rewrites on every run -- correlation blocks, per-row statistics, results -- is carved
// from it, so a genome-wide plan of a hundred batches costs the memory of the two largest.
struct Arena {
  uint8_t* base = nullptr;
  size_t cap = 0;
};

}  // namespace gb

struct gb_batch {
  gb::Ctx* ctx = nullptr;
  gb::Panel* panel = nullptr;
  int mode = gb::GRAM_MIX;
  bool ld_mode = false;      // computeLD: T x T only, full symmetric output, no solve
  double ld_diag = 1.0;      // value forced on the diagonal in ld_mode (computeLD.cpp:107: 1.0; gene.cpp:578: 1 + lambda)
  bool counts_mode = false;  // raw per-population counts
  gb_params params{};
  int64_t n_windows = 0;
  std::vector<int64_t> t_off, u_off;
  int64_t n_t_total = 0, n_u_total = 0;
  std::vector<int> plan_status;       // per window: GB_OK or a TOO_FEW_* code (window skipped)
  std::vector<int> active;            // window ids that run
  std::vector<gb::SolveWin> h_wins;   // aligned with `active`
  std::vector<gb::SolveWin> h_wins_all;   // [real windows | PD-certificate copies] as uploaded
  std::vector<gb::GramTile> h_tiles;  // [B11 tiles of all windows | B21 tiles of all windows]
  int n_tiles_tt = 0;                 // length of the B11 part
  int* d_tile_counter = nullptr;      // tile ids of the B21 Gram range, drawn by the main launch and the helper launch
  double* d_y = nullptr;              // qcat only: y = L^-1 Z1 written by the solve kernel
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // factorisation runs beside the B21 part (gb_batch_run)
  int64_t n_gather = 0;
  int n_chol_wins = 0, max_nt = 0, max_nu = 0;
  double work_gram_ops = 0, work_solve_flops = 0, work_panel_bytes = 0;
  long long tt_elems = 0, ut_elems = 0, dinv_elems = 0, counts_elems = 0;
  // host side of the plan, kept between the two planning phases
  std::vector<int32_t> h_rows_t, h_rows_u, h_gather;
  std::vector<double> h_coef, h_wgt, h_zt;
  // device
  int32_t *d_rows_t = nullptr, *d_rows_u = nullptr, *d_gather = nullptr;
  int32_t *d_pool_t = nullptr, *d_pool_u = nullptr;
  double *d_sd_t = nullptr, *d_sd_u = nullptr, *d_rq_t = nullptr;
  int32_t *d_st_sx_t = nullptr, *d_st_sx_u = nullptr;     // [n_pops][n_*_total] per listed row: sum x
  double *d_st_mean_t = nullptr, *d_st_mean_u = nullptr;  // [n_pops][n_*_total] sum x / m
  int* d_skip = nullptr;
  double gneg = 0.0;  // (sum(w)-1)_+ * max(w), +inf when the analytic PD bound does not apply
  double *d_zt = nullptr, *d_zu = nullptr, *d_info = nullptr;
  double *d_tt = nullptr, *d_ut = nullptr, *d_dinv = nullptr;
  double *d_coef = nullptr, *d_wgt = nullptr;
  int32_t* d_counts = nullptr;
  int* d_status = nullptr;
  gb::SolveWin* d_wins = nullptr;
  gb::GramTile* d_tiles = nullptr;
  int8_t* d_scratch = nullptr;
  gb::RowMaps tmaps_scratch;
  int fkind = 0;                      // tensor-core kind of this batch (GramParams::fkind)
  bool defer_flag_check = false;      // pipelined path: the panel is still being packed at plan time
  std::vector<int> h_status;          // fetch staging: [2*n_windows + 2 status words | panel flags]
  gb::GramParams gp{};
  // int8-split solve (ctx->solve_ozaki at plan time)
  bool ozaki = false;
  int oz_kpad = 0, oz_n_tiles = 0;
  long long oz_a_rows = 0, oz_b_rows = 0;
  std::vector<uint8_t> h_oz_wins, h_oz_tiles;
  void *d_oz_wins = nullptr, *d_oz_tiles = nullptr;
  double *d_x = nullptr, *d_oz_y = nullptr;
  int8_t *d_oz_pa = nullptr, *d_oz_pb = nullptr;
  uint8_t* d_oz_nan = nullptr;       // [n_u_total] rows of B21 the digit planes cannot represent (-> NaN results, as the doubles give)
  unsigned long long* d_oz_amax = nullptr;
  int* d_oz_ex = nullptr;
  // memory: `owned` pointers are freed with the batch; buffers carved from `arena` (when set) are not
  std::vector<void*> owned;
  gb::Arena arena;
  size_t arena_used = 0;
  bool clip_mode = false;             // repair batch: every window goes through the eigen-clip (MakePosDef) before the Cholesky
  double *d_eig_G = nullptr, *d_eig_V = nullptr, *d_evals = nullptr;   // clip_mode workspaces (tt_elems / n_t_total doubles)
};

namespace gb {

// Phase 1 (host only): validate, lay out tiles / windows, compute sizes.  No device work.
int batch_plan_host(gb_batch* b, const int64_t* t_off, const int64_t* rows_t, const int64_t* u_off,
                    const int64_t* rows_u, const double* z_t, const double* pop_wgt);
// Bytes of arena the batch needs when its transient buffers do not come from cudaMallocAsync.
size_t batch_arena_bytes(const gb_batch* b);
// Phase 2: device buffers + uploads.  arena.base == nullptr -> every buffer is stream-ordered pool memory.
// sync == false leaves the uploads in flight on ctx->stream (the caller synchronises once for many batches).
int batch_plan_device(gb_batch* b, Arena arena, bool sync);

gb_batch* batch_new(gb_ctx* ctx, gb_panel* panel, int64_t n_windows, const double* pop_wgt, const gb_params* params,
                    bool ld_mode, bool counts_mode, bool defer_flag_check, double ld_diag);
void batch_free_device(gb_batch* b);
int batch_fetch_enqueue(gb_batch* b, double* z_u, double* info_u, int* status_staging);
// The pieces of gb_batch_run for a caller that pipelines batches itself (gb_genome.cu): all on ctx->stream as set by the caller.
int batch_run_front(gb_batch* b, int max_ctas);   // row statistics + the B11 Gram tiles (finished)
int batch_run_chain(gb_batch* b);                 // factorisation (+ explicit L^-1 for the int8-split solve): needs B11 only
int batch_run_b21(gb_batch* b, int max_ctas);     // the B21 Gram tiles (finished)
int batch_run_solve(gb_batch* b);                 // needs all of the above
int batch_fetch_finish(gb_batch* b, const int* st, double* z_u, double* info_u, int* window_status_out);
// The same, followed by the slow path for windows the certificate could not vouch for (GB_ERR_NOT_PD) or whose
// factorisation broke down: B11 is rebuilt, eigendecomposed and clipped on the device like the reference's MakePosDef
// (util.cpp:302-318), factored and solved again; their results replace the first ones and their status becomes GB_OK.
// Synchronises ctx->stream.  Windows that still break down (NaN correlations) keep GB_ERR_BREAKDOWN and NaN results.
int batch_fetch_finish_repair(gb_batch* b, const int* st, double* z_u, double* info_u, int* window_status_out);
int pack5_layout(int n_pops, const int* pop_sizes, std::vector<int>* boff);

// Relative device time of a window (Gram multiply-adds on the tensor cores + the fp64 solve, weighted by the measured
// ratio of the two rates); windows the reference refuses (<= min SNPs, dist.cpp:146) cost nothing.
double window_cost(double n_t, double n_u, double n_samples, const gb_params& p);
// Cut range(n) into n_parts contiguous [lo, hi) runs minimising the largest run cost; cuts[n_parts + 1].
void partition_contiguous(const double* cost, int64_t n, int n_parts, int64_t* cuts);

}  // namespace gb
