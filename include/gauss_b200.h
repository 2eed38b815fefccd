/*
 * gauss_b200.h -- C-ABI of the B200-native window hot path of GAUSS.
 *
 * This is the drop-in boundary.  The reference has no plugin API; the seam
 * these entry points replace is the body of
 *     void run_dist   (std::vector<Snp*>&, Arguments&)   reference src/dist.cpp:129-227
 *     void run_distmix(std::vector<Snp*>&, Arguments&)   reference src/distmix.cpp:138-253
 *     the LD block of computeLD()                         reference src/computeLD.cpp:95-116
 * and the arithmetic they call in src/util.cpp:49-70 (CalCor), 103-124 (CalWgtCov),
 * 262-264 (MpMatMat), 298-300 (InvMat), 302-318 (MakePosDef).  The Rcpp entry points
 * (src/RcppExports.cpp:31-47,65-82,85-102) keep their signatures; INTEGRATION.md shows the
 * binding a maintainer adds.
 *
 * Conventions: every function returns an int status (GB_OK == 0), never throws, never exits,
 * never calls back into R.  The caller owns all host buffers; the library owns device memory
 * behind opaque handles.  One gb_ctx drives one GPU; calls on one ctx are not thread-safe,
 * calls on different ctxs are independent (one host thread or process per GPU).
 * There is NO CPU fallback: without a CUDA device every compute entry point fails with
 * GB_ERR_NO_DEVICE / GB_ERR_CUDA.
 */
#ifndef GAUSS_B200_H
#define GAUSS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GB_VERSION 100

#if defined(__GNUC__)
#define GB_API __attribute__((visibility("default")))
#else
#define GB_API
#endif

enum gb_status {
  GB_OK = 0,
  GB_ERR_BAD_ARG = 1,
  GB_ERR_CUDA = 2,
  GB_ERR_NO_DEVICE = 3,
  GB_ERR_OOM = 4,
  /* "Not enough number of SNPs loaded - DIST[MIX] not performed": dist.cpp:146-151, distmix.cpp:154-160 */
  GB_ERR_TOO_FEW_MEASURED = 5,
  GB_ERR_TOO_FEW_UNMEASURED = 6,
  /* B11 is not certified to satisfy lambda_min >= min_abs_eig, i.e. the reference's MakePosDef (util.cpp:302-318) may
   * have modified it.  The imputation entry points (window / batch / chromosome / pipe / genome) do not return this any
   * more: such windows are eigendecomposed and clipped on the device like MakePosDef and come back GB_OK. */
  GB_ERR_NOT_PD = 7,
  GB_ERR_UNSUPPORTED = 8,
  /* The Cholesky factorisation of B11 itself met a non-positive pivot (lambda = 0 with duplicated SNPs, negative
   * weights, NaN correlations of a monomorphic SNP ...): no result exists for the window and its z / info are NaN.
   * The reference's MakePosDef would have clipped the spectrum instead (util.cpp:302-318). */
  GB_ERR_BREAKDOWN = 9
};

typedef struct gb_ctx gb_ctx;     /* one per GPU */
typedef struct gb_panel gb_panel; /* HBM-resident packed reference panel */
typedef struct gb_batch gb_batch; /* a planned set of windows on one panel */
typedef struct gb_pipe gb_pipe;   /* asynchronous per-window pipeline on host buffers */

/* Hidden arguments of the reference (struct Arguments, gauss.h:44-62; defaults gauss.cpp:18-35). */
typedef struct gb_params {
  double lambda;              /* 0.1   ridge added as B11 diagonal := 1 + lambda (dist.cpp:172) */
  double min_abs_eig;         /* 1e-5  MakePosDef threshold (dist.cpp:181) */
  int min_num_measured_snp;   /* 10    window rejected if n_t <= this */
  int min_num_unmeasured_snp; /* 10    window rejected if n_u <= this */
  int check_pd;               /* 1: certify lambda_min(B11) > min_abs_eig (analytic bound, else a shifted Cholesky) and send
                                 uncertified windows through the eigen-clip path; 0: only detect factorisation breakdown */
  int reserved;
} gb_params;

GB_API void gb_params_default(gb_params *p);
GB_API int gb_version(void);
GB_API const char *gb_status_string(int status);

/* ---- context ------------------------------------------------------------------------------ */
GB_API int gb_ctx_create(int device, gb_ctx **out);
GB_API void gb_ctx_destroy(gb_ctx *ctx);
/* Run all work of this ctx on a caller-owned cudaStream_t (NULL restores the ctx's own stream). */
GB_API int gb_ctx_set_stream(gb_ctx *ctx, void *cuda_stream);
GB_API int gb_ctx_synchronize(gb_ctx *ctx);
/* Last error text of this ctx (or of ctx creation when ctx == NULL). */
GB_API const char *gb_last_error(const gb_ctx *ctx);
/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
GB_API int64_t gb_ctx_launch_count(const gb_ctx *ctx);

/* ---- panel packing (new step hanging off ReadGenotype, gauss.cpp:720-785) ------------------ */
/* pop_sizes[p] = individuals of the p-th FLAGGED population, in panel order -- exactly the
 * strings Snp::genotype_vec_ holds (snp.h:109).  Rows are SNPs. */
GB_API int gb_panel_create(gb_ctx *ctx, int n_pops, const int *pop_sizes, int64_t capacity_rows,
                    gb_panel **out);
/* Operand encoding of the packed rows.  Both feed the same tcgen05 Gram kernel and give the same
 * exact integer counts; they differ in bytes per dosage.
 *   GB_PANEL_INT8  one signed byte per dosage (kind::i8, int32 accumulate): exact for ANY byte the
 *                  reference's (c - '0') arithmetic can produce.
 *   GB_PANEL_E2M1  one 4-bit E2M1 float per dosage (tcgen05.mma kind::mxf4 with unit block scales, fp32
 *                  accumulate; the DEFAULT): half the HBM footprint and twice the dosages per TMA row; exact for
 *                  dosages in {0, 1, 2, 3, 4, 6}
 *                  (every product and every per-population sum is an integer below 2^24).  A row holding any other value makes the next
 *                  batch/window call fail with GB_ERR_UNSUPPORTED -- repack as GB_PANEL_INT8.
 * gb_panel_create uses GB_PANEL_E2M1 (real panels hold only '0','1','2') unless the environment
 * variable GB_PANEL_FORMAT=int8 is set; gb_run_window_strings repacks as int8 by itself. */
enum gb_panel_format { GB_PANEL_INT8 = 0, GB_PANEL_E2M1 = 1 };
GB_API int gb_panel_create_fmt(gb_ctx *ctx, int n_pops, const int *pop_sizes, int64_t capacity_rows,
                        int format, gb_panel **out);
GB_API int gb_panel_format_of(const gb_panel *panel);
GB_API void gb_panel_destroy(gb_panel *panel);
GB_API int gb_panel_clear(gb_panel *panel); /* forget all rows, keep the allocation */
GB_API int64_t gb_panel_num_rows(const gb_panel *panel);
GB_API int64_t gb_panel_num_samples(const gb_panel *panel);
/* Append SNP rows given as the reference stores them: n_rows * n_pops C strings, string
 * (r, p) holding EXACTLY pop_sizes[p] chars '0'/'1'/'2' (any 7-bit char c packs as c - '0', like the
 * reference's arithmetic) and a terminating NUL -- std::string::c_str() of Snp::genotype_vec_; a shorter or longer
 * string is GB_ERR_BAD_ARG (the reference's CalCor walks x[i].length() characters, util.cpp:55).  HOST memory. */
GB_API int gb_panel_append_strings(gb_panel *panel, int64_t n_rows, const char *const *pop_strings);
/* Same rows as one flat HOST buffer: row r at rows + r*row_stride, populations concatenated
 * (sum(pop_sizes) bytes).  is_ascii != 0: bytes are chars; 0: bytes are int8 dosages. */
GB_API int gb_panel_append_host(gb_panel *panel, int64_t n_rows, const void *rows, int64_t row_stride,
                         int is_ascii);
/* Same, but the flat buffer already lives in DEVICE memory of this ctx's GPU. */
GB_API int gb_panel_append_device(gb_panel *panel, int64_t n_rows, const void *dev_rows,
                           int64_t row_stride, int is_ascii);

/* ---- raw integer statistics (bit-exact parity surface) -------------------------------------- */
/* Per-population Gram counts S^p[a][b] = sum_k x_a,k * x_b,k over population p, plus (optional)
 * per-row sums: out_sxy [n_pops][n_a][n_b], out_sx / out_sxx [n_pops][n_a] (int32, HOST). */
GB_API int gb_gram_counts(gb_ctx *ctx, gb_panel *panel, int64_t n_a, const int64_t *rows_a, int64_t n_b,
                   const int64_t *rows_b, int32_t *out_sxy, int32_t *out_sx, int32_t *out_sxx);

/* ---- one window, host in / host out (what run_dist / run_distmix / computeLD call) ---------- */
/* rows_t: panel rows of the measured SNPs (type 1, whole extended window, bp order);
 * rows_u: panel rows of the unmeasured SNPs (type 0 inside [start_bp,end_bp]);
 * z_t: their input Z-scores.  Outputs, per unmeasured SNP: z_u = normalised imputed z
 * (dist.cpp:195-201), info_u = |b21 B11^-1 b12|. */
GB_API int gb_window_dist(gb_ctx *ctx, gb_panel *panel, int64_t n_t, const int64_t *rows_t, int64_t n_u,
                   const int64_t *rows_u, const double *z_t, const gb_params *params, double *z_u,
                   double *info_u);
/* pop_wgt: weights of the flagged populations in panel order (Arguments::pop_wgt_vec). */
GB_API int gb_window_distmix(gb_ctx *ctx, gb_panel *panel, int64_t n_t, const int64_t *rows_t,
                      int64_t n_u, const int64_t *rows_u, const double *z_t, const double *pop_wgt,
                      const gb_params *params, double *z_u, double *info_u);
/* computeLD block: cormat is n x n (symmetric; column-major == row-major), diagonal exactly 1.0. */
GB_API int gb_window_ld(gb_ctx *ctx, gb_panel *panel, int64_t n, const int64_t *rows,
                 const double *pop_wgt, double *cormat);
/* Per-gene LD blocks of jepeg() / jepegmix() (CorG, gene.cpp:300-316 and 569-587) for many genes in one batch:
 * gene g owns rows[g_off[g] .. g_off[g+1]) (>= 1 SNP); its n_g x n_g correlation matrix, `diag` forced on the
 * diagonal (1 + lambda there, 1.0 in computeLD), is written at out + sum_{h<g} n_h^2.  pop_wgt == NULL selects the
 * pooled CalCor of jepeg(). */
GB_API int gb_genes_ld(gb_ctx *ctx, gb_panel *panel, int64_t n_genes, const int64_t *g_off, const int64_t *rows,
                const double *pop_wgt, double diag, double *out);
/* jepeg() / jepegmix() (BASELINE config 5): the statistics of Gene::CalJepegPval / Gene::CalJepegmixPval (gene.cpp:288-547,
 * 553-822) for every gene of a run (jepegmix.cpp:115-139) in one call.  Genes as in gb_genes_ld; per SNP (aligned with
 * rows): z, info (Snp::GetInfo: 1 for a measured SNP) and categ_wgt[6] -- the annotation weight of the SNP in category
 * PFS, TFB, STR, TAR, CIS, TRN (gene.cpp:28-44), NaN where the SNP has no entry in that category (Snp::categ_map_).
 * lambda, min_abs_eig, categ_cor_cutoff, denorm_norm_w: Arguments defaults 0.1, 1e-5, 0.8, 3 (gauss.cpp:18-35).
 * out: 16 doubles per gene --
 *   [0] chisq  [1] df  [2] jepeg_pval  [3] number of available categories  [4] top category (0..5, -1 = ".")
 *   [5] top category p-value  [6] index of the top SNP within the gene  [7] reserved
 *   [8..13] p-value of each category (NaN = not available)  [14] bit mask of removed categories  [15] reserved
 * with the reference's defaults (chisq -1, df 0, p-values -1) when every category was removed.  pop_wgt == NULL selects
 * jepeg() (pooled CalCor).  The top SNP's p-value, 2 pnorm(|z|), is left to the caller like every other pnorm5 of the
 * output stage. */
GB_API int gb_genes_jepeg(gb_ctx *ctx, gb_panel *panel, int64_t n_genes, const int64_t *g_off, const int64_t *rows,
                   const double *pop_wgt, const double *z, const double *info, const double *categ_wgt, double lambda,
                   double min_abs_eig, double categ_cor_cutoff, int denorm_norm_w, double *out);
/* Parity/debug surface: the correlation blocks the solve consumes.  pop_wgt == NULL selects the
 * pooled Pearson r of dist().  B11 is n_t x n_t symmetric with diagonal 1+lambda; B21 is
 * n_u x n_t row-major. */
GB_API int gb_window_cor(gb_ctx *ctx, gb_panel *panel, int64_t n_t, const int64_t *rows_t, int64_t n_u,
                  const int64_t *rows_u, const double *pop_wgt, const gb_params *params,
                  double *B11, double *B21);

/* ---- prep_zmix5 pair loop (zmix.cpp:151-170) ------------------------------------------------------ */
/* For every pair i < j of the n listed SNPs (row-major over i, then j) one row of `out`, COLUMN-major
 * [n(n-1)/2][1 + n_pops] like the reference's NumericMatrix: column 0 = z[i]*z[j], column 1 + p = the Pearson r of
 * population p (CalCor(std::string&, std::string&), util.cpp:153-169; NaN where a SNP is monomorphic in p, as
 * in the reference).  Every population of the panel is used (prep_zmix5 flags all, zmix.cpp:142-144). */
GB_API int gb_zmix_pair_cor(gb_ctx *ctx, gb_panel *panel, int64_t n, const int64_t *rows, const double *z,
                     double *out);

/* ---- qcat() / qcatmix() window (run_qcat qcat.cpp:133-238, run_qcatmix qcatmix.cpp:140-269) ------------ */
/* Tests every SNP of the prediction window: measured SNPs rows_t[core_first .. core_first + n_core) (the
 * reference's [num_measured_headwing, + num_measured_pred) range of the extended window) and the unmeasured
 * SNPs rows_u.  pop_wgt == NULL -> qcat (pooled CalCor), else qcatmix (CalWgtCov).  Outputs: qcat_t and
 * qcat_chisq per tested SNP (SetQcatT / SetQcatChisq) and *num_eig (SetQcatM; CountPC util.cpp:355-388).
 * GB_ERR_TOO_FEW_MEASURED as qcat.cpp:157; qcatmix additionally GB_ERR_TOO_FEW_UNMEASURED (qcatmix.cpp:168-169);
 * GB_ERR_NOT_PD when no eigenvalue bound above eig_cutoff (default 0.01, gauss.cpp:22) could be certified. */
GB_API int gb_window_qcat(gb_ctx *ctx, gb_panel *panel, int64_t n_t, const int64_t *rows_t, const double *z_t,
                   int64_t core_first, int64_t n_core, int64_t n_u, const int64_t *rows_u,
                   const double *pop_wgt, const gb_params *params, double eig_cutoff, int *num_eig,
                   double *t_m, double *chisq_m, double *t_u, double *chisq_u);

/* ---- many windows on a resident panel (genome driver / benchmark path) ---------------------- */
/* Window w uses rows_t[t_off[w] .. t_off[w+1]) and rows_u[u_off[w] .. u_off[w+1]); z_t is
 * aligned with rows_t.  pop_wgt == NULL selects dist().  Planning uploads the descriptors once. */
GB_API int gb_batch_create(gb_ctx *ctx, gb_panel *panel, int64_t n_windows, const int64_t *t_off,
                    const int64_t *rows_t, const int64_t *u_off, const int64_t *rows_u,
                    const double *z_t, const double *pop_wgt, const gb_params *params,
                    gb_batch **out);
/* computeLD() blocks as a resident batch: window w = the n_w x n_w correlation matrix of rows_t[t_off[w] .. t_off[w+1])
 * (computeLD.cpp:95-116, diagonal forced to `diag`); only stages 0 and 1 of gb_batch_run_stage apply.  For callers that
 * keep the matrices on the device or time the Gram + mixture epilogue alone (BASELINE config 3); gb_window_ld is the
 * host-in / host-out form. */
GB_API int gb_batch_create_ld(gb_ctx *ctx, gb_panel *panel, int64_t n_windows, const int64_t *t_off,
                       const int64_t *rows_t, const double *pop_wgt, double diag, gb_batch **out);
GB_API void gb_batch_destroy(gb_batch *batch);
/* Enqueue every kernel of the batch on the ctx stream (asynchronous). */
GB_API int gb_batch_run(gb_batch *batch);
/* Copy results to HOST (aligned with rows_u) and per-window status codes; synchronises. */
GB_API int gb_batch_fetch(gb_batch *batch, double *z_u, double *info_u, int *window_status);
/* Algorithmic work of the batch: int8 Gram ops, solve fp64 flops, panel bytes. */
GB_API int gb_batch_work(const gb_batch *batch, double *gram_ops, double *solve_flops,
                  double *panel_bytes);
/* Enqueue only one stage (profiling / roofline timing): 0 = row statistics, 1 = Gram + finish pass,
 * 2 = Cholesky (+ the explicit L^-1 of the int8-split solve), 3 = solve + finalise (int8-split GEMM on tcgen05, or the
 * fp64 triangular solve with GB_SOLVE=fp64); 10 / 11 = the two halves of stage 1 (tensor-core kernel / finish pass),
 * 20 / 21 = the two halves of stage 2 (factorisation / L^-1). */
GB_API int gb_batch_run_stage(gb_batch *batch, int stage);

/* ---- 2-bit host rows ("pack2") and the chromosome driver ---------------------------------------- */
/* The packed-panel step the north star hangs off ReadGenotype (gauss.cpp:720-785): a dosage in {0,1,2} is
 * stored in two bits, 4 per byte (low bits = lower individual), population blocks on 128-dosage boundaries
 * and zero padded -- the column positions of the E2M1 device row, so the GPU side is a pure bit expansion
 * and PCIe carries a quarter of the bytes of a char row.  This is also the layout a cached on-disk packed
 * panel would use. */
GB_API int64_t gb_pack2_row_bytes(int n_pops, const int *pop_sizes);
/* HOST-side packer (CPU threads; formatting only, no statistics): rows as in gb_panel_append_host ->
 * pack2 rows at out + r*out_stride.  GB_ERR_UNSUPPORTED if a dosage outside {0,1,2} is met (keep such a
 * panel as char/int8 rows). */
GB_API int gb_pack2_rows_host(int n_pops, const int *pop_sizes, int64_t n_rows, const void *rows,
                       int64_t row_stride, int is_ascii, void *out, int64_t out_stride);
/* Append pack2 rows (HOST memory) to a GB_PANEL_E2M1 panel: copy + expand2_rows_kernel. */
GB_API int gb_panel_append_pack2_host(gb_panel *panel, int64_t n_rows, const void *rows2, int64_t row_stride);
/* One chromosome (any bp-sorted run of windows) of dist()/distmix() from pack2 HOST rows to HOST results:
 * `panel` (E2M1, capacity >= n_rows) is cleared and refilled; the rows travel in n_groups chunks on a copy
 * stream while the windows -- cut into n_groups contiguous cost-balanced batches -- run as soon as the rows
 * they touch have landed.  Arguments as gb_batch_create; window_status (optional) receives one status per
 * window; returns the first non-OK window status otherwise.  Synchronous. */
GB_API int gb_chrom_run_pack2(gb_ctx *ctx, gb_panel *panel, int64_t n_rows, const void *host_rows2,
                       int64_t row_stride, int64_t n_windows, const int64_t *t_off, const int64_t *rows_t,
                       const int64_t *u_off, const int64_t *rows_u, const double *z_t, const double *pop_wgt,
                       const gb_params *params, int n_groups, double *z_u, double *info_u, int *window_status);

/* ---- ternary host rows ("pack5") --------------------------------------------------------------------- */
/* Five dosages per byte (byte = d0 + 3 d1 + 9 d2 + 27 d3 + 81 d4): 1.6 bits per dosage against an information content of
 * log2(3) = 1.58, i.e. another fifth off the PCIe bytes of pack2.  Population p starts at a 4-byte boundary; the row is
 * padded to 16 bytes (gb_pack5_row_bytes).  Same contract as the pack2 functions above. */
GB_API int64_t gb_pack5_row_bytes(int n_pops, const int *pop_sizes);
GB_API int gb_pack5_rows_host(int n_pops, const int *pop_sizes, int64_t n_rows, const void *rows,
                       int64_t row_stride, int is_ascii, void *out, int64_t out_stride);
GB_API int gb_panel_append_pack5_host(gb_panel *panel, int64_t n_rows, const void *rows5, int64_t row_stride);
/* ... and when the ternary rows already sit in DEVICE memory of the panel's GPU (expansion kernel only). */
GB_API int gb_panel_append_pack5_device(gb_panel *panel, int64_t n_rows, const void *dev_rows5, int64_t row_stride);
GB_API int gb_chrom_run_pack5(gb_ctx *ctx, gb_panel *panel, int64_t n_rows, const void *host_rows5,
                       int64_t row_stride, int64_t n_windows, const int64_t *t_off, const int64_t *rows_t,
                       const int64_t *u_off, const int64_t *rows_u, const double *z_t, const double *pop_wgt,
                       const gb_params *params, int n_groups, double *z_u, double *info_u, int *window_status);

/* ---- genome-wide run from ONE process on 1..8 GPUs (BASELINE config 4) --------- */
/* The reference's unit of work is one dist()/distmix() call per window, each rebuilding all of its state
 * (dist.cpp:63-75); a genome run is a user-level R loop over ~2,900 such calls (the gb_init / gb_windows_submit /
 * gb_windows_wait triple SURVEY.md section 8b sketches).  A gb_genome takes the whole window list at once: it is cut into
 * contiguous, cost-balanced shards, one per GPU; every GPU gets its own host thread, context and streams, keeps the
 * panel rows its windows touch (plus the wings of its boundary windows) resident in HBM -- uploaded ONCE -- and writes
 * the (z, info) of its windows straight into the caller's per-chromosome arrays.  No collective is used: none is needed.
 * An R caller behind .Call (RcppExports.cpp:85-102) drives all GPUs of the box through these blocking / submit-wait
 * calls; nothing here calls back into R.
 *
 *   gb_genome_create          devices == NULL -> GPUs 0 .. n_gpus-1; pop_wgt == NULL -> dist(), else distmix()
 *   gb_genome_add_chromosome  one bp-sorted run of windows over one row space: arguments as gb_chrom_run_pack5 (host_rows5 =
 *                             ternary HOST rows, may be NULL when the rows will be generated by gb_genome_fill_synthetic;
 *                             `sites` = optional site index of every row for that generator).  The window lists are
 *                             copied; the host rows are NOT (they must stay valid until the upload has finished).
 *   gb_genome_plan            partitions ALL windows (chromosome order) into n_parts runs of ~equal cost and plans parts
 *                             first_part .. first_part + n_gpus - 1 on this genome's GPUs (n_parts > n_gpus: the other parts
 *                             belong to other processes, e.g. one rank per GPU under torchrun; every rank computes the
 *                             same cuts).  Batches of ~48 windows, workspaces, residency mode (ternary rows, or E2M1 operand
 *                             rows when the shard fits) are decided here.
 *   gb_genome_upload          host -> device copy of every GPU's row ranges (wait == 0: asynchronous; a run submitted next
 *                             starts each batch as soon as the rows it touches have landed)
 *   gb_genome_submit / _wait  all batches of all GPUs; z_u[c] / info_u[c] / window_status[c] are the arrays of chromosome c
 *                             (aligned with its rows_u / windows; entries may be NULL).  gpu_ms / upload_ms (optional,
 *                             [n_gpus]) receive the device-timed duration of the run / of the upload per GPU.
 *                             With window_status == NULL the first non-OK window status is returned. */
typedef struct gb_genome gb_genome;
GB_API int gb_genome_create(int n_gpus, const int *devices, int n_pops, const int *pop_sizes, const double *pop_wgt,
                     const gb_params *params, gb_genome **out);
GB_API void gb_genome_destroy(gb_genome *g);
GB_API const char *gb_genome_last_error(const gb_genome *g);
GB_API int gb_genome_add_chromosome(gb_genome *g, int64_t n_rows, const void *host_rows5, int64_t row_stride,
                             int64_t n_windows, const int64_t *t_off, const int64_t *rows_t, const int64_t *u_off,
                             const int64_t *rows_u, const double *z_t, const int64_t *sites);
GB_API int gb_genome_plan(gb_genome *g, int n_parts, int first_part);
GB_API int gb_genome_num_chromosomes(const gb_genome *g);
/* What GPU `gpu` of this genome was given (any pointer may be NULL).  e2m1_resident: per cent of its panel rows kept
 * resident in the expanded operand layout (100: no expansion at run time; a whole 33KG genome on one GPU: ~70). */
GB_API int gb_genome_shard_info(const gb_genome *g, int gpu, int64_t *first_window, int64_t *n_windows,
                         int64_t *resident_rows, int64_t *n_batches, int64_t *n_imputed, int *e2m1_resident,
                         double *gram_ops, double *solve_flops);
/* The panel rows GPU `gpu` keeps resident of chromosome `chrom` (merged [lo, hi) ranges: the measured rows of its windows
 * with their wings, and the unmeasured rows), so that a feeder only has to read those; and the way to hand them over
 * piece by piece instead of as one buffer per chromosome (pieces may overlap -- two GPUs both keep the wing between
 * their shards; the newest piece holding a row is the one read). */
GB_API int gb_genome_resident_ranges(const gb_genome *g, int gpu, int chrom, int max_ranges, int64_t *lo, int64_t *hi,
                              int *n_ranges);
GB_API int gb_genome_set_host_rows(gb_genome *g, int chrom, int64_t row_lo, int64_t n_rows, const void *host_rows5,
                            int64_t row_stride);
/* Resident ternary rows of one GPU back into HOST memory (GB_ERR_UNSUPPORTED when that GPU keeps them expanded). */
GB_API int gb_genome_download_rows(gb_genome *g, int gpu, int chrom, int64_t row_lo, int64_t n_rows, void *host_out,
                            int64_t out_stride);
GB_API int gb_genome_upload(gb_genome *g, int wait);
GB_API int gb_genome_submit(gb_genome *g, double *const *z_u, double *const *info_u, int *const *window_status);
GB_API int gb_genome_wait(gb_genome *g, double *gpu_ms, double *upload_ms);
GB_API int gb_genome_run(gb_genome *g, double *const *z_u, double *const *info_u, int *const *window_status,
                  double *gpu_ms);
GB_API int64_t gb_genome_launch_count(const gb_genome *g);
/* The partition gb_genome_plan uses, for callers that place the parts themselves: window w costs
 * N (n_u n_t + n_t^2 / 2) Gram multiply-adds + 80 (n_t^2 n_u + n_t^3 / 3) (the measured ratio of the two rates on B200);
 * windows the reference refuses (dist.cpp:146) cost nothing.  cuts has n_parts + 1 entries; cost_out (optional) n_windows.
 * Pure host code. */
GB_API int gb_partition_windows(int64_t n_windows, const int64_t *n_t, const int64_t *n_u, int64_t n_samples,
                         const gb_params *params, int n_parts, int64_t *cuts, double *cost_out);

/* ---- synthetic panel rows, generated on the device (benchmark / test DATA, not a reference path) --------- */
/* The 33KG panel is not distributable (docs/articles/ref_33KG.md:7) and BASELINE config 4 is ~10 M SNPs x 32,953
 * individuals, so SURVEY.md section 7 asks for on-device generation.  Rows come out in the ternary host format
 * (gb_pack5_row_bytes); a dosage is a pure function of (seed, chrom, site, population, individual), so every shard
 * regenerates the rows any other would.  site of row r = sites[r], or first_site + r when sites == NULL.
 * out: HOST memory (out_is_device == 0) or DEVICE memory of ctx's GPU. */
GB_API int gb_synth_pack5_rows(gb_ctx *ctx, uint64_t seed, int chrom, int64_t n_rows, const int64_t *sites,
                        int64_t first_site, int n_pops, const int *pop_sizes, void *out, int64_t out_stride,
                        int out_is_device);
/* Fill every GPU's resident rows with that generator instead of uploading them (chromosome index = `chrom`). */
GB_API int gb_genome_fill_synthetic(gb_genome *g, uint64_t seed, int wait);

/* ---- native converter of the reference's panel data file ------------------------------- */
/* Reads the reference's BGZF data file once (gauss.cpp:572-585: per SNP one text line of n_pops genotype strings over ALL
 * populations followed by n_pops allele frequencies; blocks as bgzf.c:486-536 reads them), inflating blocks and parsing
 * lines on n_threads host threads (0 = all), and writes a ".gbpack" file: per SNP a ternary row (gb_pack5_row_bytes
 * over all populations), its n_pops allele frequencies as doubles (MakeSnpVecMix's filter, gauss.cpp:631-693) and the
 * BGZF virtual offset of its line (the `fpos` column of the index file, bgzf.h:108) -- layout in gb_packfile.cu;
 * gauss_b200/packfile.py maps it.  Replaces the per-call bgzf_seek + istringstream parse of ReadGenotype
 * (gauss.cpp:720-785).  Pure host code.  err (optional) receives a message on failure. */
GB_API int gb_packfile_convert(const char *geno_path, int n_pops, const int *pop_sizes, const char *out_path,
                        int n_threads, int64_t *n_rows, double *text_bytes, double *seconds, char *err, int err_cap);

/* ---- pipe-peak probes (roofline denominators measured on the bench box at bench time) ---- */
/* MEASURED_PEAKS.json has HBM GB/s and dense bf16 TF/s only.  which: 0 = tcgen05.mma kind::i8 128x128x32 (TOP/s),
 * 1 = tcgen05.mma kind::mxf4 128x128x64 (TOP/s), 2 = fp64 mma.sync m8n8k4 (TFLOP/s), 3 = device copy (GB/s, read +
 * write).  Each keeps that one pipe busy on every SM with nothing else going on (no TMA, no epilogue); best of `reps`
 * event-timed launches. */
GB_API int gb_probe_peak(gb_ctx *ctx, int which, int reps, double *value);

/* ---- pipelined single windows, host in / host out ---------------------------------------------- */
/* What a genome loop over dist()/distmix() calls (dist.cpp:63-75 runs one window per call): the
 * host->device copy of window w+1 overlaps the kernels of window w.  `depth` device slots of
 * max_rows_per_window rows each; format < 0 selects the ctx default.  Host row buffers, z_u and
 * info_u must stay valid until the ticket is waited for, and should be page-locked (pinned) for the
 * copies to be asynchronous.  A ticket must be waited for within `depth` further submissions.
 * gb_pipe_wait returns the window's status (GB_OK, GB_ERR_TOO_FEW_*, GB_ERR_NOT_PD, ...) in
 * *window_status exactly as gb_window_dist/distmix would. */
GB_API int gb_pipe_create(gb_ctx *ctx, int n_pops, const int *pop_sizes, int64_t max_rows_per_window,
                   int depth, int format, gb_pipe **out);
GB_API void gb_pipe_destroy(gb_pipe *pipe);
/* host_rows_t / host_rows_u: n_t measured and n_u unmeasured SNP rows (sum(pop_sizes) bytes each,
 * row_stride apart; is_ascii as in gb_panel_append_host).  pop_wgt == NULL selects dist(). */
GB_API int gb_pipe_submit(gb_pipe *pipe, int64_t n_t, const void *host_rows_t, int64_t n_u,
                   const void *host_rows_u, int64_t row_stride, int is_ascii, const double *z_t,
                   const double *pop_wgt, const gb_params *params, double *z_u, double *info_u,
                   int64_t *ticket);
GB_API int gb_pipe_wait(gb_pipe *pipe, int64_t ticket, int *window_status);

/* ---- host-side mirror of the reference seam --------------------------------------------------- */
/* run_dist / run_distmix on a bp-sorted snp_vec given as parallel arrays: type (0/1/2), bp, z,
 * info and, per SNP, n_pops genotype strings (may be NULL for type 2).  Splits measured /
 * unmeasured exactly as dist.cpp:132-141, enforces the thresholds (dist.cpp:146), packs the
 * strings, runs the window on the GPU and writes z/info of the imputed SNPs back in place
 * (SetZ/SetInfo, dist.cpp:200-202).  pop_wgt == NULL -> run_dist.
 * The context keeps a pinned staging buffer and a working panel between calls (a genome is ~2,900 calls of one
 * shape) and gathers the strings on a few host threads; like every call on a gb_ctx it is not re-entrant. */
GB_API int gb_run_window_strings(gb_ctx *ctx, int64_t n_snps, const int *type, const long long *bp,
                          double *z, double *info, const char *const *pop_strings, int n_pops,
                          const int *pop_sizes, const double *pop_wgt, long long start_bp,
                          long long end_bp, const gb_params *params, int *n_measured,
                          int *n_unmeasured);

/* run_qcat / run_qcatmix (qcat.cpp:134-262, qcatmix.cpp:145-286) on the same parallel arrays: splits the SNPs as
 * qcat.cpp:139-152 (head wing, tested measured, tested unmeasured), packs the strings, runs gb_window_qcat and writes
 * qcat_m / qcat_t / qcat_chisq of the tested SNPs in place (SetQcatM / SetQcatT / SetQcatChisq); other entries are left
 * untouched.  pop_wgt == NULL -> run_qcat. */
GB_API int gb_run_qcat_strings(gb_ctx *ctx, int64_t n_snps, const int *type, const long long *bp, const double *z,
                        const char *const *pop_strings, int n_pops, const int *pop_sizes,
                        const double *pop_wgt, long long start_bp, long long end_bp, const gb_params *params,
                        double eig_cutoff, double *qcat_m, double *qcat_t, double *qcat_chisq, int *n_measured,
                        int *n_unmeasured);

#ifdef __cplusplus
}
#endif
#endif /* GAUSS_B200_H */
