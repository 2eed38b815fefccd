/*
 * gauss_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the statsleelab/gauss window hot path, used only as
 * the checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under gauss_b200/ may include, link or call it.
 *
 * Parity status: the reference ships no tests or golden vectors,
 * so this oracle is pinned against the reference's OWN functions compiled from
 * /root/reference by oracle/build_ref.sh into oracle/_ref/ (see oracle/README.md)
 * and against tests/golden/ fixtures minted from that build.
 *
 * Genotype representation: the reference keeps one std::string of '0'/'1'/'2'
 * chars per flagged population on every Snp (src/snp.h:109).  Here a SNP is a
 * flat char row of n_samples = sum(m[p]) chars; population p occupies
 * [off[p], off[p]+m[p]).  Same bytes, same order.
 */
#ifndef GAUSS_ORACLE_H
#define GAUSS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GO_OK 0
#define GO_ERR_TOO_FEW_SNPS 1 /* "Not enough number of SNPs loaded" dist.cpp:146-151 */

/* pooled Pearson r, follows src/util.cpp:49-70 */
double go_cal_cor(const char *x, const char *y, const int *m, int n_pops);
/* ancestry-weighted covariance, follows src/util.cpp:103-124 */
double go_cal_wgt_cov(const char *x, const char *y, const int *m, int n_pops, const double *w);

/* brute-force integer statistics: per population sum(x*y), sum(x), sum(x*x) */
void go_gram_counts(const char *geno_a, int64_t n_a, const char *geno_b, int64_t n_b,
                    const int *m, int n_pops, int32_t *sxy /* [P][n_a][n_b] */,
                    int32_t *sx_a /* [P][n_a] or NULL */, int32_t *sxx_a /* [P][n_a] or NULL */);

/* Eigen-dependent primitives, restated (src/util.cpp:262-264, 298-300, 302-318).
 * Matrices are column-major n x n doubles like Eigen::MatrixXd. */
int go_sym_eig(const double *A, int n, double *evals, double *evecs); /* 0 = converged */
int go_make_pos_def(double *A, int n, double min_abs_eig);  /* returns 1 if A was modified */
void go_inv_full_piv_lu(double *inv, const double *A, int n);

typedef struct {
  long long start_bp, end_bp; /* prediction (core) window, inclusive */
  double lambda;              /* 0.1  gauss.cpp:20 */
  double min_abs_eig;         /* 1e-5 gauss.cpp:21 */
  int min_num_measured_snp;   /* 10   gauss.cpp:27 */
  int min_num_unmeasured_snp; /* 10   gauss.cpp:28 */
} go_args;

void go_args_default(go_args *a);

/* run_dist (src/dist.cpp:129-227) / run_distmix (src/distmix.cpp:138-253).
 * Inputs are the bp-sorted snp_vec: type (0 unmeasured, 1 measured, 2 no-ref),
 * bp, z, genotype rows.  z[] and info[] are updated in place for type-0 SNPs
 * inside [start_bp,end_bp] exactly as SetZ/SetInfo do.  w == NULL selects
 * run_dist (CalCor), else run_distmix (CalWgtCov).
 * Optional dumps (may be NULL): B11 (n_t x n_t col-major, after MakePosDef),
 * B21 (n_u x n_t row-major). */
int go_run_window(const int *type, const long long *bp, double *z, double *info,
                  const char *geno, int64_t n_snps, const int *m, int n_pops,
                  const double *w, const go_args *args, int *n_measured, int *n_unmeasured,
                  double *B11_out, double *B21_out);

/* run_qcat (src/qcat.cpp:133-238, w == NULL) / run_qcatmix (src/qcatmix.cpp:140-269): the QCAT test of every SNP
 * of the prediction window.  qcat_m / qcat_t / qcat_chisq are indexed like snp_vec (SetQcatM / SetQcatT /
 * SetQcatChisq); entries of SNPs that are not tested are left untouched.  CountPC: util.cpp:355-388. */
int go_count_pc(const double *A, int n, double eig_cutoff);
int go_run_qcat(const int *type, const long long *bp, const double *z, const char *geno, int64_t n_snps,
                const int *m, int n_pops, const double *w, const go_args *args, double eig_cutoff,
                double *qcat_m, double *qcat_t, double *qcat_chisq);

/* per-population Pearson r (src/util.cpp:153-169) and the pair loop of prep_zmix5 (src/zmix.cpp:151-170):
 * out is column-major [n(n-1)/2][1 + n_pops], column 0 = z_i z_j, pairs i < j in row-major order. */
double go_cal_cor_pop(const char *x, const char *y, int n);
void go_zmix_pairs(const char *geno, int64_t n, const int *m, int n_pops, const double *z, double *out);

/* computeLD kernel (src/computeLD.cpp:95-116): correlation among n SNPs,
 * diagonal exactly 1.0, col-major n x n. */
void go_compute_ld(const char *geno, int64_t n, const int *m, int n_pops, const double *w,
                   double *cormat);

/* timing helper for bench: number of (sample x SNP-pair) products evaluated by
 * the last go_run_window / go_compute_ld call on this thread. */
double go_last_sample_pairs(void);

#ifdef __cplusplus
}
#endif
#endif
