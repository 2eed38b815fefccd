// gb_common.cuh -- shared declarations of the gauss_b200 device library (not part of the C-ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/gauss_b200.h"

namespace gb {

// ------------------------------------------------------------------ geometry of the packed panel
constexpr int K_ATOM = 32;    // int8 elements one tcgen05.mma kind::i8 consumes along K
constexpr int K_BLOCK = 128;  // bytes of K per smem stage row == TMA box width == swizzle span
constexpr int P_MAX = 30;     // max flagged populations (33KG, the largest panel the reference documents, has 29; 1KG 26);
                              // bounded by the epilogue's shared-memory statistics tables

// One K-segment = one population block of the packed panel (or the pooled whole row for dist()).
struct Seg {
  int koff;    // first K column of the segment (multiple of 32)
  int natoms;  // number of 32-wide K atoms to multiply (ceil(m/32))
  int m;       // individuals in the segment
  int pad;
};

// ------------------------------------------------------------------ Gram kernel interface
enum GramMode : int {
  GRAM_MIX = 0,     // CalWgtCov -> correlation (distmix / computeLD)
  GRAM_POOLED = 1,  // CalCor pooled Pearson r (dist)
  GRAM_COUNTS = 2   // raw int32 per-segment counts (parity surface)
};

struct GramTile {      // one 128 x 128 output tile
  int a_row0, b_row0;  // first row of the A (M side) / B (N side) operand in its TMA source
  int a_src, b_src;    // 0: panel, 1: gathered scratch
  int a_list0, b_list0;// index of the first A / B row in the batch row lists (stats, sd)
  int a_valid, b_valid;// rows that exist (<= 128)
  int a_is_u;          // 1: A rows come from the unmeasured list (B21 tile), 0: measured (B11 tile)
  int i0, j0;          // local coordinates of the tile inside the window's output matrix
  int ld_out;          // leading dimension of the output matrix (doubles / int32)
  long long out_off;   // element offset of the window's output matrix in the output buffer
  // int8-split solve (B21 tiles only): the finish pass writes the digit planes of B21 instead of its doubles
  int oz_ra;           // rows per digit plane of the window (0: not an int8-split batch)
  int oz_pad;
  long long oz_row0;   // row of digit plane 0 that holds the tile's first A row
};

struct GramParams {
  const GramTile* tiles;  // one descriptor per 128 x 128 output tile
  int n_tiles;            // number of tiles
  int* tile_counter;      // nullptr: tiles dealt round-robin; else a zeroed device counter several launches draw tile ids from
  int mode;
  int n_seg;
  int fkind;             // tensor-core operand kind: 0 = int8 (kind::i8), k > 0 = kind::f8f6f4 format k-1
  int raw_out;           // 1: store the raw weighted Gram sum; gram_finalize_kernel finishes it (E2M1, mixture)
  int mirror;            // 1: also store the transposed entry (full symmetric matrix, computeLD)
  int wide_fold;         // int8 mixture fold: 1 = m*sumxy - sumx*sumy may not fit int32, form it exactly in fp64
  int feed_test;         // diagnostics only (GB_GRAM_FEEDTEST): 1 = every tile loads rows 0.. (an L2-hot, fully shared feed), results meaningless
  long long* dbg;        // diagnostics only (GB_GRAM_TRACE): per-CTA stall counters, 8 per CTA; nullptr in production
  Seg seg[P_MAX];
  double coef[P_MAX];    // w_p * (m_p / (m_p - 1))          (util.cpp:117-118)
  double wgt[P_MAX];     // w_p
  double coefm[P_MAX];   // coef_p * m_p                     (regrouped fold of E2M1 panels)
  double kappa[P_MAX];   // w_p / m_p^2 - coef_p
  double n_pooled;       // total flagged individuals (dist)
  double diag;           // value forced on the diagonal of T x T tiles (1 + lambda, or 1.0)
  // per listed row and population, precomputed by row_prep_kernel (list order, [n_seg][st_ld_*]):
  const int32_t* st_sx_t;   // sum x
  const int32_t* st_sx_u;
  const double* st_mean_t;  // sum x / m   (util.cpp:119)
  const double* st_mean_u;
  long long st_ld_t, st_ld_u;
  const double* sd_t;    // per listed row: mix: sqrt(cov_ii); pooled: sqrt(N*sxx - sx^2)
  const double* sd_u;
  const int32_t* pool_t; // pooled sum x per listed row (dist)
  const int32_t* pool_u;
  double* out_tt;        // B11 buffers (column-major, lower triangle incl. whole diagonal tiles)
  double* out_ut;        // B21^T buffers ([n_t][ld_u], u contiguous)
  int32_t* out_counts;   // GRAM_COUNTS: [n_seg][n_a][n_b]
  long long counts_seg_stride;
  int8_t* oz_pa;         // int8-split solve: digit planes of B21 ([plane][u][k], k contiguous), written by the finish pass
  int oz_kpad;           // bytes per plane row
  uint8_t* oz_nan;       // [n_u_total] set when a row of B21 holds a value the digits cannot carry (NaN: a monomorphic SNP)
};

// ------------------------------------------------------------------ context / panel
struct Ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t side_stream = nullptr;   // lazily created: the factorisation beside the B21 Gram tiles (gb_batch_run)
  cudaStream_t aux_stream = nullptr;    // lazily created: the rows of L^-1 beside the factorisation's own steps (launch_cholesky)
  cudaEvent_t ev_aux = nullptr;
  char* h_stage = nullptr;              // lazily grown pinned staging of gb_panel_append_strings (two chunks of rows)
  size_t h_stage_cap = 0;
  cudaEvent_t ev_stage[2] = {nullptr, nullptr};
  void* win_panel = nullptr;            // working panel of the per-window string entry points (a gb_panel, kept between calls)
  int seg_order = 2;                    // processing order of the populations in the regrouped Gram fold (GB_SEG_ORDER)
  int chol_sms = 64;                    // SMs the B21 Gram launch leaves to it (GB_CHOL_SMS; 0 = run the stages one after another)
  cudaStream_t chrom_sides[2] = {nullptr, nullptr};     // ... and the side stream each of them forks its factorisation onto
  cudaStream_t chrom_streams[2] = {nullptr, nullptr};   // lazily created: alternating batches of the chromosome driver
  cudaStream_t copy_stream = nullptr;   // lazily created: host->device copies of the chromosome driver
  std::string err;
  int64_t launches = 0;
  // lazily grown device scratch shared by the single-window entry points
  void* fn_encode_tiled = nullptr;  // cuTensorMapEncodeTiled via cudaGetDriverEntryPoint
  int panel_format = GB_PANEL_E2M1; // format gb_panel_create uses (GB_PANEL_FORMAT=int8|e2m1 overrides)
  int e2m1_mxf4 = 1;                // E2M1 panels: 1 = kind::mxf4 (packed nibbles, K = 64), 0 = kind::f8f6f4 (GB_GRAM_KIND)
  int heavy_sms = 0;                // > 0: CTAs the persistent kernels of the int8-split solve may use (the genome driver keeps the
                                    // rest of the SMs for the factorisation chain of the next batch); 0 = sm_count
  int solve_ozaki = 1;              // 1: the solve's n_t^2 n_u term runs as an int8-split GEMM on tcgen05 (GB_SOLVE=fp64 opts out)
};

// TMA descriptors of one row-major packed-row matrix {k_elems, n_rows} with boxes of 128 K columns x
// {128, 64, 32, 16} rows (the kernels use the 128-row box).
enum MapFormat : int { MAP_INT8 = 0, MAP_E2M1_EXPAND = 1, MAP_E2M1_PACKED = 2 };
struct RowMaps {
  static constexpr int N = 4;
  CUtensorMap m[N];
};

struct Panel {
  Ctx* ctx = nullptr;
  int n_pops = 0;
  std::vector<int> pop_sizes;
  std::vector<int> koff;       // per pop first K column (padded to 32)
  int n_samples = 0;           // sum pop_sizes
  int format = GB_PANEL_E2M1;  // operand encoding of the packed rows (gb_panel_format)
  int seg_align = K_ATOM;      // population blocks start on multiples of this many K columns
  int k_elems = 0;             // K columns per packed row (multiple of 128)
  int k_stride = 0;            // bytes per packed row: k_elems (int8) or k_elems / 2 (E2M1 nibbles)
  int* d_flags = nullptr;      // [1] bit 0: a dosage outside the format's exact set was packed; bit 1: a byte outside {0,1,2}
  int64_t capacity = 0;
  int64_t n_rows = 0;
  int8_t* d_rows = nullptr;    // [capacity][k_stride bytes]
  int32_t* d_sx = nullptr;     // [n_pops][capacity]
  int32_t* d_sxx = nullptr;    // [n_pops][capacity]
  int* d_pop_sizes = nullptr;  // [n_pops]
  int* d_koff = nullptr;       // [n_pops]
  int* d_boff5 = nullptr;      // [n_pops] byte offset of each population block in a pack5 host row
  int pack5_row_bytes = 0;
  RowMaps tmaps;               // int8 rows, or E2M1 rows expanded to bytes by the TMA unit (kind::f8f6f4)
  RowMaps tmaps_packed;        // E2M1 rows kept nibble-packed in shared memory (kind::mxf4)
};

#define GB_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) {                                                                  \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                          \
      return e__ == cudaErrorMemoryAllocation ? GB_ERR_OOM : GB_ERR_CUDA;                      \
    }                                                                                          \
  } while (0)

// gb_pack.cu
int launch_pack(Ctx* ctx, Panel* panel, const void* dev_src, int64_t src_stride, int is_ascii,
                int64_t row0, int64_t n_rows);
int launch_expand2(Ctx* ctx, Panel* panel, const void* dev_src, int64_t src_stride, int64_t row0, int64_t n_rows);
int launch_expand5(Ctx* ctx, Panel* panel, const void* dev_src, int64_t src_stride, int64_t row0, int64_t n_rows);
int launch_gather_rows(Ctx* ctx, const Panel* panel, const int32_t* d_rows, int64_t n, int8_t* dst);
int launch_row_prep(Ctx* ctx, const Panel* panel, const int32_t* d_rows, int64_t n, int mode,
                    const double* d_coef, const double* d_wgt, double* d_sd, int32_t* d_pool, double* d_rq,
                    int32_t* d_st_sx, double* d_st_mean);

// gb_gram.cu
int make_row_tensor_maps(Ctx* ctx, RowMaps* out, const void* base, int64_t n_rows, int64_t k_elems,
                         int64_t k_stride_bytes, int format);
int launch_gram_finalize(Ctx* ctx, const GramParams& prm, int n_descriptors);
int launch_zmix_pairs(Ctx* ctx, const Panel* panel, const int32_t* d_counts, int n, const int32_t* d_rows,
                      const double* d_z, double* d_out);
int launch_gram(Ctx* ctx, const RowMaps& panel, const RowMaps& scratch, const GramParams& prm, int max_ctas);

// gb_solve.cu
struct SolveWin {        // per-window solve descriptor
  int n_t, n_u;
  int ld_t, ld_u;        // leading dims of B11 (column-major, lower) and B21^T (row-major n_t x ld_u)
  long long off_tt;      // element offset of B11 / L in the TT buffer
  long long off_ut;      // element offset of B21^T / W in the UT buffer
  long long off_t;       // offset into per-measured arrays (z_t, y)
  long long off_u;       // offset into per-unmeasured arrays (z_u, info_u)
  long long off_dinv;    // element offset of this window's inverted 64x64 diagonal blocks
  int flags;             // bit 0: publish inv(L_kk) (real factorisation; clear for the PD-certificate copy)
  int pad;
};
// X = L^-1 row block by row block (int8-split solve): buffers of the REAL windows (no certificate copies)
struct LinvArgs {
  const SolveWin* d_wins;
  int n_real;
  const double* d_tt;
  const double* d_dinv;
  double* d_x;
  const double* d_zt;
  double* d_y;
  unsigned long long* d_amax;
};
int launch_cholesky(Ctx* ctx, const SolveWin* d_wins, int n_wins, int max_nt, double* d_tt, double* d_dinv,
                    int* d_status, const int* d_skip, const LinvArgs* linv = nullptr);
int launch_linv_rows(Ctx* ctx, const LinvArgs& a, int max_nt);
int launch_pd_bound(Ctx* ctx, const SolveWin* d_wins, int n_real, const double* d_rq_t, double lambda,
                    double gneg, double min_abs_eig, int* d_skip);
int launch_trsm_finalize(Ctx* ctx, const SolveWin* d_wins, int n_wins, int max_nt, int max_nu, const double* d_tt,
                         const double* d_dinv, double* d_ut, const double* d_zt, double* d_zu, double* d_info,
                         double* d_y_out, int tri = 0, unsigned long long* d_amax = nullptr);
int launch_qcat_patch(Ctx* ctx, const SolveWin* d_wins, double* d_ut, int n_u, int core_first, int n_core, double diag);
int launch_qcat_finalize(Ctx* ctx, const SolveWin* d_wins, const double* d_ut, const double* d_y, int n_tested,
                         int num_eig, double* d_qt, double* d_qchisq);
int launch_eig_jacobi(Ctx* ctx, const SolveWin* d_wins, int n_wins, int max_nt, double* d_tt, double* d_G, double* d_V,
                      double* d_evals, double min_abs_eig, int clip, int* d_n_clipped);
int launch_copy_shift(Ctx* ctx, const SolveWin* d_wins, int n_wins, const double* d_src, double* d_dst,
                      double shift, const int* d_skip);

// gb_ozaki.cu: the solve's n_t^2 n_u term as an exact int8-split GEMM on tcgen05 (kind::i8)
constexpr int OZ_NDIG = 6;    // signed 8-bit digit planes per operand
constexpr int OZ_QBITS = 8 * OZ_NDIG - 2;   // fixed-point bits below the scale: v = 2^e q 2^-46, |q| < 2^46 (1.95 2^46 for B21)
constexpr unsigned long long OZ_BIAS = 0x0000808080808080ull;   // sum_p 128 * 256^p: q + BIAS has the unsigned digits d_p + 128
size_t ozaki_win_bytes();
size_t ozaki_tile_bytes();
void ozaki_plan(const SolveWin* wins, int n_wins, int kpad, int n_ctas, void* ow_out, std::vector<uint8_t>* tiles_out, long long* a_rows,
                long long* b_rows);
int launch_ozaki_solve(Ctx* ctx, const SolveWin* d_wins, const void* d_ow, const void* h_ow, int n_wins, const void* d_tiles,
                       int n_tiles, int kpad, const double* d_x, const double* d_ut, int slice_b21, int8_t* d_planes_a,
                       long long a_rows, int8_t* d_planes_b, long long b_rows, unsigned long long* d_amax, int* d_ex,
                       const double* d_y, uint8_t* d_nan, double* d_zu, double* d_info);
void ozaki_tile_rows(const void* h_ow, int win, long long* a_row0, int* ra);

// gb_gene.cu
int launch_jepeg_genes(Ctx* ctx, const void* d_genes, int n_genes, const double* d_tt, const double* d_z, const double* d_info,
                       const double* d_cw, double min_abs_eig, double cor_cutoff, int denorm, double* d_out);
size_t jepeg_gene_desc_bytes();
void jepeg_gene_desc_fill(void* dst, long long off_tt, long long snp0, int ld, int n);

}  // namespace gb
