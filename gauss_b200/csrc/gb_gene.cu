// gb_gene.cu -- jepeg() / jepegmix() per-gene statistics (BASELINE config 5).
//
// Replaces Gene::CalJepegPval / Gene::CalJepegmixPval (gene.cpp:288-547, 553-822) for ALL genes of a run at once: the
// per-gene LD blocks CorG (1 + lambda on the diagonal, gene.cpp:300-316 / 569-587) come out of the same tensor-core
// Gram path as B11 (one batch, gb_genes_ld), and one thread per gene then does the <= 6-category algebra the reference
// does with Eigen: W (weights x sqrt(info), gene.cpp:859-877), CovU = W CorG W^T, U = W Z, the category p-values,
// removal of collinear (|CorU| > categ_cor_cutoff) and low-variance (CovU_ii < (W W^T)_ii / denorm_norm_w) categories,
// MakePosDef + InvMat on the <= 6 x 6 CovX (util.cpp:298-318), chi-square = X^T CovX^-1 X and its upper tail
// (R::pchisq) -- in the reference's operation order.  Thousands of genes with 1-10 SNPs each: one launch.
#include <cstring>

#include "gb_batch.cuh"

namespace gb {

namespace {

constexpr int KC = 6;   // categories: PFS, TFB, STR, TAR, CIS, TRN (gene.cpp:28-44)

// cyclic Jacobi on a symmetric n x n (n <= 6) matrix: eigenvalues in ev, eigenvectors in the COLUMNS of V
__device__ void jacobi_eig(double (&A)[KC][KC], int n, double (&ev)[KC], double (&V)[KC][KC]) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0.0;
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        if (fabs(A[p][q]) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; i++) ev[i] = A[i][i];
}

// inverse by Gauss-Jordan elimination with complete pivoting (FullPivLU::inverse, util.cpp:298-300)
__device__ void inv_full_pivot(const double (&A)[KC][KC], int n, double (&inv)[KC][KC]) {
  double M[KC][2 * KC];
  int colperm[KC];
  for (int i = 0; i < n; i++) {
    colperm[i] = i;
    for (int j = 0; j < n; j++) {
      M[i][j] = A[i][j];
      M[i][n + j] = i == j ? 1.0 : 0.0;
    }
  }
  for (int k = 0; k < n; k++) {
    int pr = k, pc = k;
    double best = -1.0;
    for (int i = k; i < n; i++)
      for (int j = k; j < n; j++)
        if (fabs(M[i][j]) > best) {
          best = fabs(M[i][j]);
          pr = i;
          pc = j;
        }
    if (pr != k)
      for (int j = 0; j < 2 * n; j++) {
        const double t = M[k][j];
        M[k][j] = M[pr][j];
        M[pr][j] = t;
      }
    if (pc != k) {
      for (int i = 0; i < n; i++) {
        const double t = M[i][k];
        M[i][k] = M[i][pc];
        M[i][pc] = t;
      }
      const int t = colperm[k];
      colperm[k] = colperm[pc];
      colperm[pc] = t;
    }
    const double piv = M[k][k];
    for (int j = 0; j < 2 * n; j++) M[k][j] /= piv;
    for (int i = 0; i < n; i++) {
      if (i == k) continue;
      const double f = M[i][k];
      if (f == 0.0) continue;
      for (int j = 0; j < 2 * n; j++) M[i][j] -= f * M[k][j];
    }
  }
  // columns were permuted: row k of the right half belongs to unknown colperm[k]
  for (int k = 0; k < n; k++)
    for (int j = 0; j < n; j++) inv[colperm[k]][j] = M[k][n + j];
}

// upper tail of the chi-square distribution for 1 <= df <= 6 (R::pchisq(x, df, lower = 0, log = 0)): closed forms of
// the regularised upper incomplete gamma function Q(df / 2, x / 2)
__device__ double chisq_upper_tail(double x, int df) {
  if (!(x > 0.0)) return 1.0;
  const double h = 0.5 * x, e = exp(-h);
  if ((df & 1) == 0) {
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < df / 2; k++) {
      term *= h / k;
      sum += term;
    }
    return e * sum;
  }
  double q = erfc(sqrt(h));
  double term = sqrt(2.0 * x / 3.14159265358979323846) * e;   // 2 sqrt(h / pi) e^-h
  for (int k = 1; k <= df / 2; k++) {
    q += term;
    term *= x / (2.0 * k + 1.0);
  }
  return q;
}

struct GeneDesc {
  long long off_tt;   // CorG block: column-major n x ld in the batch's B11 buffer (full symmetric, diagonal 1 + lambda)
  long long snp0;     // first SNP of the gene in the per-SNP arrays
  int ld, n;
};

__global__ void __launch_bounds__(64)
jepeg_gene_kernel(const GeneDesc* __restrict__ genes, int n_genes, const double* __restrict__ tt,
                  const double* __restrict__ z, const double* __restrict__ info, const double* __restrict__ cw,
                  double min_abs_eig, double cor_cutoff, int denorm, double* __restrict__ out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_genes) return;
  const GeneDesc d = genes[g];
  const double* C = tt + d.off_tt;
  const double* zz = z + d.snp0;
  const double* inf = info + d.snp0;
  const double* wv = cw + d.snp0 * KC;
  double* o = out + (long long)g * 16;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  o[0] = -1.0, o[1] = 0.0, o[2] = -1.0, o[4] = -1.0, o[5] = -1.0, o[6] = -1.0, o[7] = -1.0;   // Gene::Gene defaults (gene.cpp:50-62)
  for (int c = 0; c < KC; c++) o[8 + c] = nan;
  // categories with at least one annotated SNP (Gene::RunJepegmix, gene.cpp:187-262)
  int idx[KC], k = 0;
  for (int c = 0; c < KC; c++) {
    int cnt = 0;
    for (int i = 0; i < d.n; i++) cnt += wv[i * KC + c] == wv[i * KC + c];
    if (cnt) idx[k++] = c;
  }
  o[3] = k;
  o[14] = 0.0;
  {  // top SNP: first maximum of |z| (Gene::GetTopSNP, gene.cpp:894-904)
    int top = 0;
    for (int i = 0; i < d.n; i++)
      if (fabs(zz[top]) < fabs(zz[i])) top = i;
    o[6] = top;
  }
  if (k == 0) return;
  auto W = [&](int a, int i) {   // GetW, gene.cpp:859-877: weight (0 when the SNP lacks the category) x sqrt(info)
    const double v = wv[i * KC + idx[a]];
    return __dmul_rn(v == v ? v : 0.0, sqrt(inf[i]));
  };
  double CovU[KC][KC], U[KC], WWd[KC];
  for (int a = 0; a < k; a++) {
    double s = 0.0, u = 0.0;
    for (int i = 0; i < d.n; i++) {
      const double w = W(a, i);
      s = __dadd_rn(s, __dmul_rn(w, w));            // (W W^T)_aa
      u = __dadd_rn(u, __dmul_rn(w, zz[i]));        // U = W Z
    }
    WWd[a] = s;
    U[a] = u;
    for (int b = 0; b < k; b++) CovU[a][b] = 0.0;
    for (int j = 0; j < d.n; j++) {                 // (W CorG)_aj, then CovU = (W CorG) W^T  (MpMatMat twice)
      double t = 0.0;
      for (int i = 0; i < d.n; i++) t = __dadd_rn(t, __dmul_rn(W(a, i), C[(long long)j * d.ld + i]));
      for (int b = 0; b < k; b++) CovU[a][b] = __dadd_rn(CovU[a][b], __dmul_rn(t, W(b, j)));
    }
  }
  bool rmv[KC];
  double pv[KC];
  for (int a = 0; a < k; a++) {
    rmv[a] = false;
    const double u = U[a] / sqrt(CovU[a][a]);
    pv[a] = erfc(fabs(u) * 0.70710678118654752440);   // 2 * pnorm5(|u|, 0, 1, lower = 0)
    o[8 + idx[a]] = pv[a];
  }
  for (int j = k - 1; j > 0; j--)                    // collinear categories (CnvrtCovToCor + gene.cpp:667-675)
    for (int i = 0; i < j; i++) {
      const double cor = CovU[i][j] / (sqrt(CovU[i][i]) * sqrt(CovU[j][j]));
      if (fabs(cor) > cor_cutoff) {
        rmv[j] = true;
        break;
      }
    }
  for (int a = 0; a < k; a++)                        // low-variance categories (gene.cpp:685-691)
    if (CovU[a][a] < WWd[a] / denorm) rmv[a] = true;
  int df = 0, mask = 0;
  for (int a = 0; a < k; a++) {
    df += !rmv[a];
    mask |= rmv[a] ? 1 << idx[a] : 0;
  }
  o[14] = mask;
  o[1] = df;
  if (df == 0) return;
  double X[KC], CovX[KC][KC];
  {
    int r = 0;
    for (int a = 0; a < k; a++) {
      if (rmv[a]) continue;
      X[r] = U[a];
      int c = 0;
      for (int b = 0; b < k; b++)
        if (!rmv[b]) CovX[r][c++] = CovU[a][b];
      r++;
    }
  }
  {  // MakePosDef(CovX, min_abs_eig) (util.cpp:302-318): clip the spectrum from below and rebuild only if needed
    double A[KC][KC], ev[KC], V[KC][KC];
    for (int i = 0; i < df; i++)
      for (int j = 0; j < df; j++) A[i][j] = i >= j ? CovX[i][j] : CovX[j][i];   // Eigen reads the lower triangle
    jacobi_eig(A, df, ev, V);
    double mn = ev[0];
    for (int i = 1; i < df; i++) mn = fmin(mn, ev[i]);
    if (mn < min_abs_eig) {
      for (int i = 0; i < df; i++)
        for (int j = 0; j < df; j++) {
          double s = 0.0;
          for (int q = 0; q < df; q++) s += V[i][q] * fmax(ev[q], min_abs_eig) * V[j][q];
          CovX[i][j] = s;
        }
    }
  }
  double Inv[KC][KC];
  inv_full_pivot(CovX, df, Inv);
  double chisq = 0.0;
  for (int j = 0; j < df; j++) {                     // (X^T CovX^-1) X
    double t = 0.0;
    for (int i = 0; i < df; i++) t = __dadd_rn(t, __dmul_rn(X[i], Inv[i][j]));
    chisq = __dadd_rn(chisq, __dmul_rn(t, X[j]));
  }
  o[0] = chisq;
  o[2] = chisq_upper_tail(chisq, df);
  int top = 0;                                       // Gene::GetTopCateg (gene.cpp:880-891)
  for (int a = 0; a < k; a++)
    if ((pv[top] > pv[a]) & !rmv[a]) top = a;
  o[4] = idx[top];
  o[5] = pv[top];
}

}  // namespace

int launch_jepeg_genes(Ctx* ctx, const void* d_genes, int n_genes, const double* d_tt, const double* d_z, const double* d_info,
                       const double* d_cw, double min_abs_eig, double cor_cutoff, int denorm, double* d_out) {
  if (n_genes <= 0) return GB_OK;
  jepeg_gene_kernel<<<(unsigned)((n_genes + 63) / 64), 64, 0, ctx->stream>>>(static_cast<const GeneDesc*>(d_genes), n_genes, d_tt,
                                                                            d_z, d_info, d_cw, min_abs_eig, cor_cutoff, denorm, d_out);
  GB_CUDA(cudaGetLastError());
  ctx->launches++;
  return GB_OK;
}

size_t jepeg_gene_desc_bytes() { return sizeof(GeneDesc); }
void jepeg_gene_desc_fill(void* dst, long long off_tt, long long snp0, int ld, int n) {
  GeneDesc d;
  d.off_tt = off_tt;
  d.snp0 = snp0;
  d.ld = ld;
  d.n = n;
  std::memcpy(dst, &d, sizeof(d));
}

}  // namespace gb
