// ref_glue.cpp -- builds oracle/_ref/libgauss_ref.so (test infrastructure, NOT product code).
//
// Wraps the REFERENCE'S OWN hot-path code, compiled from /root/reference/src where it lies,
// behind the same C API as oracle/gauss_oracle.c so tests can diff the two:
//   * CalCor / CalWgtCov            src/util.cpp:49-70, 103-124   (extracted verbatim at build time)
//   * run_dist / run_distmix        src/dist.cpp:129-227, src/distmix.cpp:138-253 (ditto)
//   * computeLD kernel block        src/computeLD.cpp:95-116      (ditto)
//   * Arguments::Arguments defaults src/gauss.cpp (compiled unmodified, with bgzf.c: see ref_files.cpp)
//   * class Snp                     src/snp.{h,cpp}               (compiled unmodified)
// The extracted fragments are written by oracle/build_ref.sh into oracle/_ref/gen/*.inc
// (git-ignored; reference sources are never copied into the repository history).
// Eigen is not installed here, so the three Eigen algorithms the reference calls are
// provided below from the restatements in gauss_oracle.c; everything else is the
// reference's code path.
#include <deque>
#include <map>
#include <string>
#include <vector>
#include <cstring>

#include <RcppEigen.h>
#include "snp.h"
#include "gauss.h"
#include "util.h"

extern "C" {
#include "gauss_oracle_internal.h"
}

using namespace Rcpp;

// ---- Eigen-backed helpers of src/util.cpp, re-provided on the oracle's restatements ----
void MpMatMat(Eigen::MatrixXd& result, const Eigen::Ref<const Eigen::MatrixXd>& m1,
              const Eigen::Ref<const Eigen::MatrixXd>& m2) {  // util.cpp:262-264
  Eigen::MatrixXd out(m1.rows(), m2.cols());
  for (long j = 0; j < m2.cols(); j++)
    for (long i = 0; i < m1.rows(); i++) {
      double acc = 0.0;
      for (long k = 0; k < m1.cols(); k++) acc += m1(i, k) * m2(k, j);
      out(i, j) = acc;
    }
  result = out;
}
void InvMat(Eigen::MatrixXd& inverse, const Eigen::Ref<const Eigen::MatrixXd>& m1) {  // util.cpp:298-300
  inverse.resize(m1.rows(), m1.cols());
  gor_inv_full_piv_lu(inverse.data(), m1.data(), (int)m1.rows());
}
void MakePosDef(Eigen::MatrixXd& m1, double min_abs_eig) {  // util.cpp:302-318
  gor_make_pos_def(m1.data(), (int)m1.rows(), min_abs_eig);
}
void CholeskyMat(Eigen::MatrixXd& result, const Eigen::Ref<const Eigen::MatrixXd>& m) {  // util.cpp:271-274 (Eigen::LLT)
  const long n = m.rows();
  result = Eigen::MatrixXd::Zero(n, n);
  for (long j = 0; j < n; j++) {
    double d = m(j, j);
    for (long k = 0; k < j; k++) d -= result(j, k) * result(j, k);
    d = std::sqrt(d);
    result(j, j) = d;
    for (long i = j + 1; i < n; i++) {
      double v = m(i, j);
      for (long k = 0; k < j; k++) v -= result(i, k) * result(j, k);
      result(i, j) = v / d;
    }
  }
}
int CountPC(const Eigen::Ref<const Eigen::MatrixXd>& m1, double eig_cutoff) {  // util.cpp:355-388
  return gor_count_pc(m1.data(), (int)m1.rows(), eig_cutoff);
}
double CalCor(const Eigen::Ref<const Eigen::VectorXd>& x, const Eigen::Ref<const Eigen::VectorXd>& y) {  // util.cpp:193-202
  const long n = x.size();
  double mx = 0.0, my = 0.0;
  for (long i = 0; i < n; i++) mx += x(i);
  for (long i = 0; i < n; i++) my += y(i);
  mx /= n;
  my /= n;
  double sxx = 0.0, syy = 0.0, sxy = 0.0;
  for (long i = 0; i < n; i++) sxx += (x(i) - mx) * (x(i) - mx);
  for (long i = 0; i < n; i++) syy += (y(i) - my) * (y(i) - my);
  for (long i = 0; i < n; i++) sxy += (x(i) - mx) * (y(i) - my);
  return sxy / std::sqrt(sxx * syy);
}
double CalVar(const Eigen::Ref<const Eigen::VectorXd>& x) {  // util.cpp:214-219
  const long n = x.size();
  double mx = 0.0;
  for (long i = 0; i < n; i++) mx += x(i);
  mx /= n;
  double s = 0.0;
  for (long i = 0; i < n; i++) s += (x(i) - mx) * (x(i) - mx);
  return s / (double)(n - 1);
}
void LoadProgressBar(int) {}  // util.cpp:449-461 prints a text bar; silent here
#include "gen/util_284_296.inc"   // CnvrtCovToCor (Eigen-free body, extracted verbatim)

// ---- the reference's own code, extracted at build time -----------------------------------
#include "gen/util_49_70.inc"
#include "gen/util_103_124.inc"
#include "gen/util_153_169.inc"   // CalCor(std::string&, std::string&)
// (Arguments::Arguments, gauss.cpp:18-35, now comes from the reference's own gauss.cpp compiled whole: build_ref.sh)
void run_dist(std::vector<Snp*>& snp_vec, Arguments& args);
void run_distmix(std::vector<Snp*>& snp_vec, Arguments& args);
#include "gen/dist_129_227.inc"
#include "gen/distmix_138_253.inc"
void run_qcat(std::vector<Snp*>& snp_vec, Arguments& args);
void run_qcatmix(std::vector<Snp*>& snp_vec, Arguments& args);
#include "gen/qcat_134_262.inc"
#include "gen/qcatmix_145_286.inc"

static void ref_ld_block(std::vector<Snp*>& snp_vec_measured, Arguments& args, double* out) {
  int num_measured = snp_vec_measured.size();
#include "gen/computeLD_95_116.inc"
  std::memcpy(out, Cor_Mat.data(), sizeof(double) * (size_t)num_measured * num_measured);
}

// CorG of one gene as Gene::CalJepegmixPval (gene.cpp:569-586) / Gene::CalJepegPval (gene.cpp:305-314) build it
static void ref_gene_corg(std::vector<Snp*>& gene_snp_vec, std::vector<double>& pop_wgt_vec_, double lambda_, bool mix,
                          double* out) {
  Eigen::MatrixXd CorG = Eigen::MatrixXd::Zero(gene_snp_vec.size(), gene_snp_vec.size());
  if (mix) {
#include "gen/gene_569_586.inc"
  } else {
#include "gen/gene_305_314.inc"
  }
  std::memcpy(out, CorG.data(), sizeof(double) * gene_snp_vec.size() * gene_snp_vec.size());
}

// ---- same C surface as gauss_oracle.h ------------------------------------------------------
static __thread double g_pairs = 0.0;

static std::vector<std::string> split_pops(const char* row, const int* m, int n_pops) {
  std::vector<std::string> v;
  const char* p = row;
  for (int k = 0; k < n_pops; k++) {
    v.emplace_back(p, p + m[k]);
    p += m[k];
  }
  return v;
}

extern "C" {

double go_last_sample_pairs(void) { return g_pairs; }

// CorG block of one gene (n SNP rows), column-major n x n; w == NULL -> jepeg (pooled CalCor)
void go_gene_corg(const char* geno, int64_t n, const int* m, int n_pops, const double* w, double lambda, double* out) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  std::vector<Snp> store((size_t)n);
  std::vector<Snp*> vec;
  for (int64_t i = 0; i < n; i++) {
    auto gv = split_pops(geno + i * N, m, n_pops);
    store[(size_t)i].SetGenotypeVec(gv);
    vec.push_back(&store[(size_t)i]);
  }
  std::vector<double> wv;
  if (w) wv.assign(w, w + n_pops);
  ref_gene_corg(vec, wv, lambda, w != nullptr, out);
}

// One gene through the reference's own Gene::RunJepegmix / Gene::RunJepeg (gene.cpp compiled UNMODIFIED: category
// counting 187-280, CorG, W, CovU = W CorG W^T, collinear / low-variance category removal, MakePosDef + InvMat on the
// <= 6 x 6 CovX, chi-square and R::pchisq -- gene.cpp:288-547, 553-822).  categ_wgt: [n][6], NaN = the SNP is not
// annotated in that category.  out: chisq, df, jepeg_pval, top_categ (0..5 or -1), top_categ_pval, top_snp, top_snp_pval.
#include "gene.h"
void go_gene_jepeg(const char* geno, int64_t n, const int* m, int n_pops, const double* w, const double* z,
                   const double* info, const double* categ_wgt, double lambda, double min_abs_eig, double categ_cor_cutoff,
                   int denorm_norm_w, double* out) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  Arguments args;
  args.lambda = lambda;
  args.min_abs_eig = min_abs_eig;
  args.categ_cor_cutoff = categ_cor_cutoff;
  args.denorm_norm_w = denorm_norm_w;
  if (w) args.pop_wgt_vec.assign(w, w + n_pops);
  std::vector<Snp> store((size_t)n);
  std::vector<Snp*> vec;
  for (int64_t i = 0; i < n; i++) {
    Snp& s = store[(size_t)i];
    auto gv = split_pops(geno + i * N, m, n_pops);
    s.SetGenotypeVec(gv);
    s.SetZ(z[i]);
    s.SetInfo(info[i]);
    s.SetType(1);
    s.SetRsid("snp" + std::to_string(i));
    s.SetGeneid("G");
    for (int c = 0; c < 6; c++)
      if (categ_wgt[i * 6 + c] == categ_wgt[i * 6 + c]) s.SetCateg(c, categ_wgt[i * 6 + c]);
    vec.push_back(&s);
  }
  Gene gene(args);
  if (w) gene.RunJepegmix(vec);
  else gene.RunJepeg(vec);
  out[0] = gene.GetChisq();
  out[1] = gene.GetDf();
  out[2] = gene.GetJepegPval();
  static const char* names[6] = {"PFS", "TFB", "STR", "TAR", "CIS", "TRN"};
  out[3] = -1;
  for (int c = 0; c < 6; c++)
    if (gene.GetTopCategName() == names[c]) out[3] = c;
  out[4] = gene.GetTopCategPval();
  out[5] = -1;
  for (int64_t i = 0; i < n; i++)
    if (gene.GetTopSnpId() == store[(size_t)i].GetRsid()) out[5] = (double)i;
  out[6] = gene.GetTopSnpPval();
}

// run_qcat / run_qcatmix: the reference's own bodies around the restated Eigen algorithms above
int go_run_qcat(const int* type, const long long* bp, const double* z, const char* geno, int64_t n_snps, const int* m,
                int n_pops, const double* w, const go_args* a, double eig_cutoff, double* qcat_m, double* qcat_t,
                double* qcat_chisq) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  Arguments args;
  args.chr = 0;
  args.start_bp = a->start_bp;
  args.end_bp = a->end_bp;
  args.lambda = a->lambda;
  args.eig_cutoff = eig_cutoff;
  args.min_num_measured_snp = a->min_num_measured_snp;
  args.min_num_unmeasured_snp = a->min_num_unmeasured_snp;
  args.num_samples = (int)N;
  if (w) args.pop_wgt_vec.assign(w, w + n_pops);
  std::vector<Snp> store((size_t)n_snps);
  std::vector<Snp*> snp_vec;
  for (int64_t i = 0; i < n_snps; i++) {
    Snp& s = store[(size_t)i];
    s.SetBp(bp[i]);
    s.SetType(type[i]);
    s.SetZ(z[i]);
    s.SetQcatM(-1);
    auto gv = split_pops(geno + i * N, m, n_pops);
    s.SetGenotypeVec(gv);
    snp_vec.push_back(&s);
  }
  try {
    if (w) run_qcatmix(snp_vec, args);
    else run_qcat(snp_vec, args);
  } catch (const std::runtime_error&) {
    return GO_ERR_TOO_FEW_SNPS;
  }
  for (int64_t i = 0; i < n_snps; i++) {
    if (store[(size_t)i].GetQcatM() < 0) continue;   // not tested
    qcat_m[i] = store[(size_t)i].GetQcatM();
    qcat_t[i] = store[(size_t)i].GetQcatT();
    qcat_chisq[i] = store[(size_t)i].GetQcatChisq();
  }
  return GO_OK;
}

// the pair loop of prep_zmix5 (zmix.cpp:151-170) around the reference's own CalCor(std::string&, std::string&)
void go_zmix_pairs(const char* geno, int64_t n, const int* m, int n_pops, const double* z, double* out) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  std::vector<std::vector<std::string>> g;
  for (int64_t i = 0; i < n; i++) g.push_back(split_pops(geno + i * N, m, n_pops));
  const int64_t total = n * (n - 1) / 2;
  int64_t row = 0;
  for (int64_t i = 0; i < n; i++)
    for (int64_t j = i + 1; j < n; j++) {
      out[row] = z[i] * z[j];
      for (int k = 0; k < n_pops; k++) out[(int64_t)(k + 1) * total + row] = CalCor(g[i][k], g[j][k]);
      row++;
    }
}

double go_cal_cor(const char* x, const char* y, const int* m, int n_pops) {
  auto vx = split_pops(x, m, n_pops), vy = split_pops(y, m, n_pops);
  return CalCor(vx, vy);
}

double go_cal_wgt_cov(const char* x, const char* y, const int* m, int n_pops, const double* w) {
  auto vx = split_pops(x, m, n_pops), vy = split_pops(y, m, n_pops);
  std::vector<double> wv(w, w + n_pops);
  return CalWgtCov(vx, vy, wv);
}

int go_run_window(const int* type, const long long* bp, double* z, double* info, const char* geno,
                  int64_t n_snps, const int* m, int n_pops, const double* w, const go_args* a,
                  int* n_measured, int* n_unmeasured, double* B11_out, double* B21_out) {
  (void)B11_out; (void)B21_out;  // the reference exposes no intermediate dumps
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  Arguments args;
  args.chr = 0;
  args.start_bp = a->start_bp;
  args.end_bp = a->end_bp;
  args.lambda = a->lambda;
  args.min_abs_eig = a->min_abs_eig;
  args.min_num_measured_snp = a->min_num_measured_snp;
  args.min_num_unmeasured_snp = a->min_num_unmeasured_snp;
  args.num_samples = (int)N;
  if (w) args.pop_wgt_vec.assign(w, w + n_pops);
  std::vector<Snp> store((size_t)n_snps);
  std::vector<Snp*> snp_vec;
  int nt = 0, nu = 0;
  for (int64_t i = 0; i < n_snps; i++) {
    Snp& s = store[(size_t)i];
    s.SetBp(bp[i]);
    s.SetType(type[i]);
    s.SetZ(z[i]);
    s.SetInfo(info[i]);
    auto gv = split_pops(geno + i * N, m, n_pops);
    s.SetGenotypeVec(gv);
    snp_vec.push_back(&s);
    if (type[i] == 1) nt++;
    if (type[i] == 0 && bp[i] >= a->start_bp && bp[i] <= a->end_bp) nu++;
  }
  if (n_measured) *n_measured = nt;
  if (n_unmeasured) *n_unmeasured = nu;
  g_pairs = (double)N * ((double)nt * (nt - 1) / 2 + (double)nu * nt + (w ? nt + nu : 0));
  try {
    if (w) run_distmix(snp_vec, args);
    else run_dist(snp_vec, args);
  } catch (const std::runtime_error&) {
    return GO_ERR_TOO_FEW_SNPS;
  }
  for (int64_t i = 0; i < n_snps; i++) {
    z[i] = store[(size_t)i].GetZ();
    info[i] = store[(size_t)i].GetInfo();
  }
  return GO_OK;
}

void go_compute_ld(const char* geno, int64_t n, const int* m, int n_pops, const double* w, double* cormat) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  Arguments args;
  args.pop_wgt_vec.assign(w, w + n_pops);
  std::vector<Snp> store((size_t)n);
  std::vector<Snp*> v;
  for (int64_t i = 0; i < n; i++) {
    auto gv = split_pops(geno + i * N, m, n_pops);
    store[(size_t)i].SetGenotypeVec(gv);
    store[(size_t)i].SetType(1);
    v.push_back(&store[(size_t)i]);
  }
  g_pairs = (double)N * ((double)n * (n - 1) / 2 + n);
  ref_ld_block(v, args, cormat);
}
}  // extern "C"
