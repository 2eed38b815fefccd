/*
 * gauss_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See gauss_oracle.h for scope and parity status.  Every function cites the
 * reference file:line (relative to /root/reference) whose behaviour it follows.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off oracle/gauss_oracle.c -o oracle/libgauss_oracle.so -lm
 * (-ffp-contract=off: R's default x86-64 build has no FMA contraction, and the
 *  bit-level comparisons in tests/ rely on plain IEEE mul/add ordering.)
 */
#include "gauss_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static __thread double g_sample_pairs = 0.0;
double go_last_sample_pairs(void) { return g_sample_pairs; }

/* ---- src/util.cpp:49-70 : CalCor(vector<string>&, vector<string>&) -------- */
double go_cal_cor(const char *x, const char *y, const int *m, int n_pops) {
  int num_samples = 0;
  double sumx = 0, sumy = 0, sumxsq = 0, sumysq = 0, sumxy = 0;
  const char *px = x, *py = y;
  for (int p = 0; p < n_pops; p++) {
    int len = m[p];
    for (int j = 0; j < len; j++) {
      double xv = (double)(px[j] - '0');
      double yv = (double)(py[j] - '0');
      sumx += xv;
      sumy += yv;
      sumxsq += xv * xv;
      sumysq += yv * yv;
      sumxy += xv * yv;
    }
    px += len;
    py += len;
    num_samples += len;
  }
  g_sample_pairs += num_samples;
  double numer = num_samples * sumxy - sumx * sumy;
  double denor = sqrt(num_samples * sumxsq - sumx * sumx) * sqrt(num_samples * sumysq - sumy * sumy);
  return numer / denor;
}

/* ---- src/util.cpp:103-124 : CalWgtCov ------------------------------------- */
double go_cal_wgt_cov(const char *x, const char *y, const int *m, int n_pops, const double *w) {
  double wsumcov = 0, wsum_mi_mj = 0, wsum_mi = 0, wsum_mj = 0;
  const char *px = x, *py = y;
  for (int p = 0; p < n_pops; p++) {
    int len = m[p];
    double sumx = 0, sumy = 0, sumxy = 0;
    double wgt = w[p];
    for (int j = 0; j < len; j++) {
      double xv = (double)(px[j] - '0');
      double yv = (double)(py[j] - '0');
      sumx += xv;
      sumy += yv;
      sumxy += xv * yv;
    }
    double factor = ((double)len) / (len - 1);
    wsumcov += wgt * factor * (len * sumxy - sumx * sumy);
    wsum_mi_mj += wgt * (sumx / len) * (sumy / len);
    wsum_mi += wgt * (sumx / len);
    wsum_mj += wgt * (sumy / len);
    px += len;
    py += len;
    g_sample_pairs += len;
  }
  return (wsumcov + wsum_mi_mj - wsum_mi * wsum_mj);
}

/* ---- brute-force integer statistics (Appendix B of SURVEY.md) ------------- */
void go_gram_counts(const char *geno_a, int64_t n_a, const char *geno_b, int64_t n_b, const int *m,
                    int n_pops, int32_t *sxy, int32_t *sx_a, int32_t *sxx_a) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  int64_t off = 0;
  for (int p = 0; p < n_pops; p++) {
    for (int64_t i = 0; i < n_a; i++) {
      const char *xa = geno_a + i * N + off;
      if (sx_a || sxx_a) {
        int32_t s = 0, q = 0;
        for (int k = 0; k < m[p]; k++) {
          int v = xa[k] - '0';
          s += v;
          q += v * v;
        }
        if (sx_a) sx_a[(int64_t)p * n_a + i] = s;
        if (sxx_a) sxx_a[(int64_t)p * n_a + i] = q;
      }
      if (!sxy) continue;
      for (int64_t j = 0; j < n_b; j++) {
        const char *xb = geno_b + j * N + off;
        int32_t acc = 0;
        for (int k = 0; k < m[p]; k++) acc += (xa[k] - '0') * (xb[k] - '0');
        sxy[((int64_t)p * n_a + i) * n_b + j] = acc;
      }
    }
    off += m[p];
  }
}


/* ---- symmetric eigen-decomposition ----------------------------------------
 * Stands in for Eigen::SelfAdjointEigenSolver<MatrixXd> (src/util.cpp:304), a
 * third-party dependency absent from /root/reference (RcppEigen, unpinned in
 * DESCRIPTION:12-16; CRAN 0.3.4.x bundles Eigen 3.4.0).  Eigen's published
 * algorithm is Householder reduction to tridiagonal form followed by implicit
 * symmetric QR iterations; restated here in the classic EISPACK tred2/tql2
 * formulation.  Eigenvalues ascending, eigenvectors column-major, like Eigen. */
#define VV(i, j) V[(size_t)(i) * n + (j)] /* row-major work matrix */

static void householder_tridiag(double *V, int n, double *d, double *e) {
  for (int j = 0; j < n; j++) d[j] = VV(n - 1, j);
  for (int i = n - 1; i > 0; i--) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; k++) scale += fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; j++) {
        d[j] = VV(i - 1, j);
        VV(i, j) = 0.0;
        VV(j, i) = 0.0;
      }
    } else {
      for (int k = 0; k < i; k++) {
        d[k] /= scale;
        h += d[k] * d[k];
      }
      double f = d[i - 1];
      double g = sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h -= f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; j++) e[j] = 0.0;
      for (int j = 0; j < i; j++) {
        f = d[j];
        VV(j, i) = f;
        g = e[j] + VV(j, j) * f;
        for (int k = j + 1; k <= i - 1; k++) {
          g += VV(k, j) * d[k];
          e[k] += VV(k, j) * f;
        }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; j++) {
        e[j] /= h;
        f += e[j] * d[j];
      }
      double hh = f / (h + h);
      for (int j = 0; j < i; j++) e[j] -= hh * d[j];
      for (int j = 0; j < i; j++) {
        f = d[j];
        g = e[j];
        for (int k = j; k <= i - 1; k++) VV(k, j) -= (f * e[k] + g * d[k]);
        d[j] = VV(i - 1, j);
        VV(i, j) = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; i++) {
    VV(n - 1, i) = VV(i, i);
    VV(i, i) = 1.0;
    double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; k++) d[k] = VV(k, i + 1) / h;
      for (int j = 0; j <= i; j++) {
        double g = 0.0;
        for (int k = 0; k <= i; k++) g += VV(k, i + 1) * VV(k, j);
        for (int k = 0; k <= i; k++) VV(k, j) -= g * d[k];
      }
    }
    for (int k = 0; k <= i; k++) VV(k, i + 1) = 0.0;
  }
  for (int j = 0; j < n; j++) {
    d[j] = VV(n - 1, j);
    VV(n - 1, j) = 0.0;
  }
  VV(n - 1, n - 1) = 1.0;
  e[0] = 0.0;
}
#undef VV

/* Q holds eigenvector i as the contiguous row Q[i*n .. i*n+n) (== column-major
 * eigenvector matrix), so the plane rotations below stream through memory. */
static int tridiag_ql(double *Q, int n, double *d, double *e) {
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = 2.220446049250313e-16;
  for (int l = 0; l < n; l++) {
    double t = fabs(d[l]) + fabs(e[l]);
    if (t > tst1) tst1 = t;
    int mm = l;
    while (mm < n - 1 && fabs(e[mm]) > eps * tst1) mm++;
    if (mm > l) {
      int iter = 0;
      do {
        if (++iter > 100) return 1;
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = hypot(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f += h;
        p = d[mm];
        double c = 1.0, c2 = 1.0, c3 = 1.0;
        double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = mm - 1; i >= l; i--) {
          c3 = c2;
          c2 = c;
          s2 = s;
          g = c * e[i];
          h = c * p;
          r = hypot(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          double *qi = Q + (size_t)i * n, *qi1 = Q + (size_t)(i + 1) * n;
          for (int k = 0; k < n; k++) {
            double hv = qi1[k];
            qi1[k] = s * qi[k] + c * hv;
            qi[k] = c * qi[k] - s * hv;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (fabs(e[l]) > eps * tst1);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
  for (int i = 0; i < n - 1; i++) { /* ascending */
    int k = i;
    double p = d[i];
    for (int j = i + 1; j < n; j++)
      if (d[j] < p) {
        k = j;
        p = d[j];
      }
    if (k != i) {
      d[k] = d[i];
      d[i] = p;
      double *qi = Q + (size_t)i * n, *qk = Q + (size_t)k * n;
      for (int j = 0; j < n; j++) {
        double tv = qi[j];
        qi[j] = qk[j];
        qk[j] = tv;
      }
    }
  }
  return 0;
}

int go_sym_eig(const double *A, int n, double *evals, double *evecs) {
  if (n == 1) {
    evals[0] = A[0];
    if (evecs) evecs[0] = 1.0;
    return 0;
  }
  double *e = (double *)malloc(sizeof(double) * (size_t)n);
  double *V = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *Q = evecs ? evecs : (double *)malloc(sizeof(double) * (size_t)n * n);
  memcpy(V, A, sizeof(double) * (size_t)n * n); /* symmetric: layout-agnostic */
  householder_tridiag(V, n, evals, e);
  for (int i = 0; i < n; i++) /* Q[i][k] = V[k][i] */
    for (int k = 0; k < n; k++) Q[(size_t)i * n + k] = V[(size_t)k * n + i];
  int rc = tridiag_ql(Q, n, evals, e);
  free(e);
  free(V);
  if (!evecs) free(Q);
  return rc;
}

/* ---- src/util.cpp:302-318 : MakePosDef ------------------------------------ */
int go_make_pos_def(double *A, int n, double min_abs_eig) {
  double *ev = (double *)malloc(sizeof(double) * (size_t)n);
  double *Q = (double *)malloc(sizeof(double) * (size_t)n * n);
  int modified = 0;
  if (go_sym_eig(A, n, ev, Q) != 0) goto done; /* solver.info() != Success -> return */
  double mn = ev[0];
  for (int i = 1; i < n; i++)
    if (ev[i] < mn) mn = ev[i];
  if (mn < min_abs_eig) {
    for (int i = 0; i < n; i++)
      if (ev[i] < min_abs_eig) ev[i] = min_abs_eig;
    /* m1 = V * diag(ev) * V^T ; Q row i == eigenvector i */
    memset(A, 0, sizeof(double) * (size_t)n * n);
    for (int k = 0; k < n; k++) {
      const double *q = Q + (size_t)k * n;
      double lam = ev[k];
      for (int j = 0; j < n; j++) {
        double s = lam * q[j];
        double *col = A + (size_t)j * n;
        for (int i = 0; i < n; i++) col[i] += q[i] * s;
      }
    }
    modified = 1;
  }
done:
  free(ev);
  free(Q);
  return modified;
}

/* ---- src/util.cpp:298-300 : InvMat = m1.fullPivLu().inverse() --------------
 * Eigen::FullPivLU is absent here (see above); its published algorithm is
 * Gaussian elimination with complete pivoting, P A Q = L U, and the inverse is
 * obtained by solving against the identity.  Column-major throughout. */
void go_inv_full_piv_lu(double *inv, const double *A, int n) {
  size_t nn = (size_t)n * n;
  double *LU = (double *)malloc(sizeof(double) * nn);
  int *rperm = (int *)malloc(sizeof(int) * (size_t)n); /* row transpositions */
  int *cperm = (int *)malloc(sizeof(int) * (size_t)n); /* col transpositions */
  memcpy(LU, A, sizeof(double) * nn);
#define M(i, j) LU[(size_t)(j) * n + (i)]
  for (int k = 0; k < n; k++) {
    int pr = k, pc = k;
    double best = -1.0;
    for (int j = k; j < n; j++)
      for (int i = k; i < n; i++) {
        double a = fabs(M(i, j));
        if (a > best) {
          best = a;
          pr = i;
          pc = j;
        }
      }
    rperm[k] = pr;
    cperm[k] = pc;
    if (best == 0.0) { /* singular: remaining transpositions are identity */
      for (int t = k + 1; t < n; t++) rperm[t] = t, cperm[t] = t;
      break;
    }
    if (pr != k)
      for (int j = 0; j < n; j++) {
        double t = M(k, j);
        M(k, j) = M(pr, j);
        M(pr, j) = t;
      }
    if (pc != k)
      for (int i = 0; i < n; i++) {
        double t = M(i, k);
        M(i, k) = M(i, pc);
        M(i, pc) = t;
      }
    double piv = M(k, k);
    for (int i = k + 1; i < n; i++) M(i, k) /= piv;
    for (int j = k + 1; j < n; j++) {
      double ukj = M(k, j);
      if (ukj == 0.0) continue;
      for (int i = k + 1; i < n; i++) M(i, j) -= M(i, k) * ukj;
    }
  }
  /* inverse = Q * U^-1 * L^-1 * P : solve for each column of P*I */
  double *col = (double *)malloc(sizeof(double) * (size_t)n);
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < n; i++) col[i] = (i == c) ? 1.0 : 0.0;
    for (int k = 0; k < n; k++) /* apply P */
      if (rperm[k] != k) {
        double t = col[k];
        col[k] = col[rperm[k]];
        col[rperm[k]] = t;
      }
    for (int k = 0; k < n; k++) { /* L y = b (unit lower) */
      double yk = col[k];
      if (yk == 0.0) continue;
      for (int i = k + 1; i < n; i++) col[i] -= M(i, k) * yk;
    }
    for (int k = n - 1; k >= 0; k--) { /* U x = y */
      col[k] /= M(k, k);
      double xk = col[k];
      if (xk == 0.0) continue;
      for (int i = 0; i < k; i++) col[i] -= M(i, k) * xk;
    }
    for (int k = n - 1; k >= 0; k--) /* apply Q */
      if (cperm[k] != k) {
        double t = col[k];
        col[k] = col[cperm[k]];
        col[cperm[k]] = t;
      }
    memcpy(inv + (size_t)c * n, col, sizeof(double) * (size_t)n);
  }
#undef M
  free(col);
  free(LU);
  free(rperm);
  free(cperm);
}

/* ---- defaults: src/gauss.cpp:18-35 ---------------------------------------- */
void go_args_default(go_args *a) {
  a->start_bp = 0;
  a->end_bp = 0;
  a->lambda = 0.1;
  a->min_abs_eig = 1e-5;
  a->min_num_measured_snp = 10;
  a->min_num_unmeasured_snp = 10;
}

/* ---- src/dist.cpp:129-210 (w == NULL) / src/distmix.cpp:138-236 ------------ */
int go_run_window(const int *type, const long long *bp, double *z, double *info, const char *geno,
                  int64_t n_snps, const int *m, int n_pops, const double *w, const go_args *args,
                  int *n_measured, int *n_unmeasured, double *B11_out, double *B21_out) {
  g_sample_pairs = 0.0;
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  int64_t *meas = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_snps + 1));
  int64_t *unme = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_snps + 1));
  int nt = 0, nu = 0;
  /* dist.cpp:132-141: type 0 inside the prediction window -> unmeasured;
   * type 1 anywhere in the extended window -> measured; type 2 ignored */
  for (int64_t i = 0; i < n_snps; i++) {
    if (type[i] == 0 && bp[i] >= args->start_bp && bp[i] <= args->end_bp)
      unme[nu++] = i;
    else if (type[i] == 1)
      meas[nt++] = i;
  }
  if (n_measured) *n_measured = nt;
  if (n_unmeasured) *n_unmeasured = nu;
  if (nt <= args->min_num_measured_snp || nu <= args->min_num_unmeasured_snp) { /* dist.cpp:146 */
    free(meas);
    free(unme);
    return GO_ERR_TOO_FEW_SNPS;
  }
  size_t ntt = (size_t)nt * nt;
  double *B11 = (double *)calloc(ntt, sizeof(double));
  double *B11Inv = (double *)malloc(sizeof(double) * ntt);
  double *Z1 = (double *)malloc(sizeof(double) * (size_t)nt);
  double *b21 = (double *)malloc(sizeof(double) * (size_t)nt);
  double *t21 = (double *)malloc(sizeof(double) * (size_t)nt);
  double *sd = NULL;
  for (int i = 0; i < nt; i++) Z1[i] = z[meas[i]];
  if (w) { /* distmix.cpp:180-187 */
    sd = (double *)malloc(sizeof(double) * (size_t)(nt + nu));
    for (int i = 0; i < nt; i++) {
      const char *g = geno + meas[i] * N;
      sd[i] = sqrt(go_cal_wgt_cov(g, g, m, n_pops, w));
    }
    for (int i = 0; i < nu; i++) {
      const char *g = geno + unme[i] * N;
      sd[nt + i] = sqrt(go_cal_wgt_cov(g, g, m, n_pops, w));
    }
  }
  for (int i = 0; i < nt; i++) { /* dist.cpp:171-179, distmix.cpp:190-200 */
    B11[(size_t)i * nt + i] = 1.0 + args->lambda;
    const char *gi = geno + meas[i] * N;
    for (int j = i + 1; j < nt; j++) {
      const char *gj = geno + meas[j] * N;
      double v;
      if (w) {
        double cov = go_cal_wgt_cov(gi, gj, m, n_pops, w);
        v = cov / (sd[i] * sd[j]);
      } else {
        v = go_cal_cor(gi, gj, m, n_pops);
      }
      B11[(size_t)j * nt + i] = v;
      B11[(size_t)i * nt + j] = v;
    }
  }
  go_make_pos_def(B11, nt, args->min_abs_eig); /* dist.cpp:181 */
  go_inv_full_piv_lu(B11Inv, B11, nt);         /* dist.cpp:182 */
  if (B11_out) memcpy(B11_out, B11, sizeof(double) * ntt);
  for (int u = 0; u < nu; u++) { /* dist.cpp:187-202, distmix.cpp:209-228 */
    const char *gu = geno + unme[u] * N;
    for (int j = 0; j < nt; j++) {
      const char *gj = geno + meas[j] * N;
      if (w) {
        double cov = go_cal_wgt_cov(gu, gj, m, n_pops, w);
        b21[j] = cov / (sd[nt + u] * sd[j]);
      } else {
        b21[j] = go_cal_cor(gu, gj, m, n_pops);
      }
    }
    if (B21_out) memcpy(B21_out + (size_t)u * nt, b21, sizeof(double) * (size_t)nt);
    /* b21B11Inv = b21 * B11Inv (1 x nt) */
    for (int j = 0; j < nt; j++) {
      const double *col = B11Inv + (size_t)j * nt;
      double acc = 0.0;
      for (int k = 0; k < nt; k++) acc += b21[k] * col[k];
      t21[j] = acc;
    }
    double zz = 0.0, val = 0.0;
    for (int k = 0; k < nt; k++) zz += t21[k] * Z1[k];
    for (int k = 0; k < nt; k++) val += t21[k] * b21[k];
    double inf = fabs(val);
    z[unme[u]] = zz / sqrt(inf);
    info[unme[u]] = inf;
  }
  free(meas);
  free(unme);
  free(B11);
  free(B11Inv);
  free(Z1);
  free(b21);
  free(t21);
  free(sd);
  return GO_OK;
}

/* ---- src/qcat.cpp:133-238 (w == NULL) / src/qcatmix.cpp:140-269 ------------------------------
 * Tests every SNP of the prediction window: L = chol(B11), LInvZ1 = L^-1 Z1, and for a SNP with
 * correlation row b (a row of B11 for measured SNPs, of B21 for unmeasured ones) r = Pearson
 * correlation of LInvZ1 and L^-1 b over the measured SNPs, qcat_t = sqrt(num_eig - 3) r,
 * qcat_chisq = (num_eig - 3) r^2, num_eig = CountPC(B11, eig_cutoff) (util.cpp:355-388).
 * Outputs are indexed like snp_vec; untouched entries keep their input value. */
static double vec_cor(const double *x, const double *y, int n) { /* util.cpp:193-202 */
  double mx = 0.0, my = 0.0;
  for (int i = 0; i < n; i++) mx += x[i];
  for (int i = 0; i < n; i++) my += y[i];
  mx /= n;
  my /= n;
  double sxx = 0.0, syy = 0.0, sxy = 0.0;
  for (int i = 0; i < n; i++) sxx += (x[i] - mx) * (x[i] - mx);
  for (int i = 0; i < n; i++) syy += (y[i] - my) * (y[i] - my);
  for (int i = 0; i < n; i++) sxy += (x[i] - mx) * (y[i] - my);
  return sxy / sqrt(sxx * syy);
}

int go_count_pc(const double *A, int n, double eig_cutoff) { /* util.cpp:355-388 */
  double *ev = (double *)malloc(sizeof(double) * (size_t)n);
  int num = n;
  if (go_sym_eig(A, n, ev, NULL) != 0) {
    free(ev);
    return num;
  }
  if (ev[0] < eig_cutoff)
    for (int i = 0; i < n; i++)
      if (ev[i] < eig_cutoff) num--;
  free(ev);
  return num;
}

int go_run_qcat(const int *type, const long long *bp, const double *z, const char *geno, int64_t n_snps,
                const int *m, int n_pops, const double *w, const go_args *args, double eig_cutoff,
                double *qcat_m, double *qcat_t, double *qcat_chisq) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  int64_t *meas = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_snps + 1));
  int64_t *unme = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_snps + 1));
  int nt = 0, nu = 0, headwing = 0, npred = 0;
  for (int64_t i = 0; i < n_snps; i++) { /* qcat.cpp:139-152 */
    if (type[i] == 0 && bp[i] >= args->start_bp && bp[i] <= args->end_bp) {
      unme[nu++] = i;
    } else if (type[i] == 1) {
      meas[nt++] = i;
      if (bp[i] < args->start_bp) headwing++;
      else if (bp[i] >= args->start_bp && bp[i] <= args->end_bp) npred++;
    }
  }
  /* run_qcat checks the measured count only (qcat.cpp:157); run_qcatmix also the unmeasured one (qcatmix.cpp:168-169) */
  if (nt <= args->min_num_measured_snp || (w && nu <= args->min_num_unmeasured_snp)) {
    free(meas);
    free(unme);
    return GO_ERR_TOO_FEW_SNPS;
  }
  size_t ntt = (size_t)nt * nt;
  double *B11 = (double *)calloc(ntt, sizeof(double));
  double *L = (double *)calloc(ntt, sizeof(double));
  double *LInv = (double *)malloc(sizeof(double) * ntt);
  double *y = (double *)malloc(sizeof(double) * (size_t)nt);
  double *b = (double *)malloc(sizeof(double) * (size_t)nt);
  double *lb = (double *)malloc(sizeof(double) * (size_t)nt);
  double *sd = NULL;
  if (w) { /* qcatmix.cpp:196-205 */
    sd = (double *)malloc(sizeof(double) * (size_t)(nt + nu));
    for (int i = 0; i < nt; i++) sd[i] = sqrt(go_cal_wgt_cov(geno + meas[i] * N, geno + meas[i] * N, m, n_pops, w));
    for (int i = 0; i < nu; i++) sd[nt + i] = sqrt(go_cal_wgt_cov(geno + unme[i] * N, geno + unme[i] * N, m, n_pops, w));
  }
  for (int i = 0; i < nt; i++) { /* qcat.cpp:185-193, qcatmix.cpp:208-219 */
    B11[(size_t)i * nt + i] = 1.0 + args->lambda;
    for (int j = i + 1; j < nt; j++) {
      double v = w ? go_cal_wgt_cov(geno + meas[i] * N, geno + meas[j] * N, m, n_pops, w) / (sd[i] * sd[j])
                   : go_cal_cor(geno + meas[i] * N, geno + meas[j] * N, m, n_pops);
      B11[(size_t)j * nt + i] = v;
      B11[(size_t)i * nt + j] = v;
    }
  }
  const int num_eig = go_count_pc(B11, nt, eig_cutoff); /* qcat.cpp:203 */
  /* CholeskyMat (util.cpp:271-274, Eigen::LLT): lower factor, column by column */
  for (int j = 0; j < nt; j++) {
    double d = B11[(size_t)j * nt + j];
    for (int k = 0; k < j; k++) d -= L[(size_t)k * nt + j] * L[(size_t)k * nt + j];
    d = sqrt(d);
    L[(size_t)j * nt + j] = d;
    for (int i = j + 1; i < nt; i++) {
      double v = B11[(size_t)j * nt + i];
      for (int k = 0; k < j; k++) v -= L[(size_t)k * nt + i] * L[(size_t)k * nt + j];
      L[(size_t)j * nt + i] = v / d;
    }
  }
  go_inv_full_piv_lu(LInv, L, nt); /* qcat.cpp:207 */
  for (int i = 0; i < nt; i++) {   /* LInvZ1 = LInv * Z1, qcat.cpp:208 */
    double acc = 0.0;
    for (int k = 0; k < nt; k++) acc += LInv[(size_t)k * nt + i] * z[meas[k]];
    y[i] = acc;
  }
  for (int t = 0; t < npred + nu; t++) { /* qcat.cpp:221-250 */
    const int is_m = t < npred;
    const int64_t snp = is_m ? meas[headwing + t] : unme[t - npred];
    if (is_m) {
      for (int j = 0; j < nt; j++) b[j] = B11[(size_t)j * nt + headwing + t];
    } else {
      for (int j = 0; j < nt; j++)
        b[j] = w ? go_cal_wgt_cov(geno + snp * N, geno + meas[j] * N, m, n_pops, w) / (sd[nt + (t - npred)] * sd[j])
                 : go_cal_cor(geno + snp * N, geno + meas[j] * N, m, n_pops);
    }
    for (int i = 0; i < nt; i++) {
      double acc = 0.0;
      for (int k = 0; k < nt; k++) acc += LInv[(size_t)k * nt + i] * b[k];
      lb[i] = acc;
    }
    const double r = vec_cor(y, lb, nt);
    qcat_m[snp] = num_eig;
    qcat_t[snp] = sqrt((double)(num_eig - 3)) * r;
    qcat_chisq[snp] = (num_eig - 3) * r * r;
  }
  free(meas); free(unme); free(B11); free(L); free(LInv); free(y); free(b); free(lb); free(sd);
  return GO_OK;
}

/* ---- src/util.cpp:153-169 : CalCor(std::string&, std::string&), one population ---------------- */
double go_cal_cor_pop(const char *x, const char *y, int n) {
  double xi = 0, yi = 0, sumx = 0, sumy = 0, sumxsq = 0, sumysq = 0, sumxy = 0;
  for (int i = 0; i < n; i++) {
    xi = (double)(x[i] - '0');
    yi = (double)(y[i] - '0');
    sumx += xi;
    sumy += yi;
    sumxsq += xi * xi;
    sumysq += yi * yi;
    sumxy += xi * yi;
  }
  double numer = n * sumxy - sumx * sumy;
  double denor = sqrt((n)*sumxsq - sumx * sumx) * sqrt((n)*sumysq - sumy * sumy);
  return numer / denor;
}

/* ---- src/zmix.cpp:151-170 : the pair loop of prep_zmix5 -----------------------------------------
 * One output row per SNP pair i < j (row-major over i, then j): column 0 = z_i z_j, column 1 + k =
 * per-population Pearson r of population k.  out is COLUMN-major [n(n-1)/2][1 + n_pops] like the
 * Rcpp::NumericMatrix the reference fills. */
void go_zmix_pairs(const char *geno, int64_t n, const int *m, int n_pops, const double *z, double *out) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  const int64_t total = n * (n - 1) / 2;
  int64_t row = 0;
  for (int64_t i = 0; i < n; i++)
    for (int64_t j = i + 1; j < n; j++) {
      out[row] = z[i] * z[j];
      int64_t off = 0;
      for (int k = 0; k < n_pops; k++) {
        out[(int64_t)(k + 1) * total + row] = go_cal_cor_pop(geno + i * N + off, geno + j * N + off, m[k]);
        off += m[k];
      }
      row++;
    }
}

/* ---- src/computeLD.cpp:95-116 ---------------------------------------------- */
void go_compute_ld(const char *geno, int64_t n, const int *m, int n_pops, const double *w,
                   double *cormat) {
  g_sample_pairs = 0.0;
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  double *sd = (double *)malloc(sizeof(double) * (size_t)n);
  for (int64_t i = 0; i < n; i++) {
    const char *g = geno + i * N;
    sd[i] = sqrt(go_cal_wgt_cov(g, g, m, n_pops, w));
  }
  for (int64_t i = 0; i < n; i++) {
    cormat[(size_t)i * n + i] = 1.0;
    const char *gi = geno + i * N;
    for (int64_t j = i + 1; j < n; j++) {
      double cov = go_cal_wgt_cov(gi, geno + j * N, m, n_pops, w);
      double cor = cov / (sd[i] * sd[j]);
      cormat[(size_t)j * n + i] = cor;
      cormat[(size_t)i * n + j] = cor;
    }
  }
  free(sd);
}
