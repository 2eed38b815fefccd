// run_window_patched.cpp -- the Rcpp-side patch of INTEGRATION.md section 2, COMPILED (test infrastructure).
//
// These are the bodies a maintainer drops into src/distmix.cpp / src/dist.cpp in place of lines 165-236 / 156-210: the
// reference's own signature `void run_distmix(std::vector<Snp*>&, Arguments&)` over the reference's own Snp and
// Arguments classes (src/snp.h, src/gauss.h, compiled unmodified), calling the C-ABI of libgauss_b200.so.  R and Rcpp
// are not installed here, so the file is compiled against the same 48-line Rcpp.h stand-in the reference's code is
// compiled with in oracle/_ref, and linked with the product library into oracle/_ref/libgauss_patched.so by
// oracle/build_ref.sh.  tests/test_rcpp_patch.py feeds identical std::vector<Snp*> to this and to the UNPATCHED
// run_distmix / run_dist of oracle/_ref and compares GetZ() / GetInfo() and the thrown messages.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include <Rcpp.h>
#include "snp.h"
#include "gauss.h"
#include "gauss_b200.h"

// ---- begin patch --------------------------------------------------------------------------
static gb_ctx* gauss_b200_ctx() {            // one context per R session
  static gb_ctx* ctx = nullptr;
  if (!ctx && gb_ctx_create(0, &ctx) != GB_OK)
    Rcpp::stop(std::string("gauss_b200: ") + gb_last_error(nullptr));   // no CPU fallback
  return ctx;
}

static void gauss_b200_run_window(std::vector<Snp*>& snp_vec, Arguments& args, bool mix, const char* tool) {
  const int64_t n = (int64_t)snp_vec.size();
  std::vector<int> pop_sizes;                                          // flagged populations, panel order
  for (size_t k = 0; k < args.ref_pop_size_vec.size(); k++)
    if (args.pop_flag_vec[k]) pop_sizes.push_back(args.ref_pop_size_vec[k]);
  const int P = (int)pop_sizes.size();
  std::vector<int> type((size_t)n);
  std::vector<long long> bp((size_t)n);
  std::vector<double> z((size_t)n), info((size_t)n);
  std::vector<const char*> geno((size_t)n * P, nullptr);
  for (int64_t i = 0; i < n; i++) {
    Snp* s = snp_vec[(size_t)i];
    type[(size_t)i] = s->GetType();
    bp[(size_t)i] = s->GetBp();
    z[(size_t)i] = s->GetZ();
    info[(size_t)i] = s->GetInfo();
    std::vector<std::string>& g = s->GetGenotypeVec();                 // snp.h:68,109 (filled by ReadGenotype)
    for (int k = 0; k < P && k < (int)g.size(); k++) geno[(size_t)(i * P + k)] = g[(size_t)k].c_str();
  }
  gb_params prm;
  gb_params_default(&prm);
  prm.lambda = args.lambda;                                            // gauss.cpp:18-35
  prm.min_abs_eig = args.min_abs_eig;
  prm.min_num_measured_snp = args.min_num_measured_snp;
  prm.min_num_unmeasured_snp = args.min_num_unmeasured_snp;
  int n_t = 0, n_u = 0;
  const int rc = gb_run_window_strings(gauss_b200_ctx(), n, type.data(), bp.data(), z.data(), info.data(), geno.data(), P,
                                       pop_sizes.data(), mix ? args.pop_wgt_vec.data() : nullptr, args.start_bp, args.end_bp,
                                       &prm, &n_t, &n_u);
  if (rc == GB_ERR_TOO_FEW_MEASURED || rc == GB_ERR_TOO_FEW_UNMEASURED) {   // dist.cpp:146-151, distmix.cpp:154-160
    Rcpp::Rcout << std::endl << "Number of measured SNPs: " << n_t << std::endl
                << "Number of unmeasured SNPs: " << n_u << std::endl;
    Rcpp::stop(std::string("Not enough number of SNPs loaded - ") + tool + " not performed");
  }
  if (rc == GB_ERR_NOT_PD)
    Rcpp::warning("gauss_b200: B11 is not certified positive definite above min_abs_eig; the reference's MakePosDef "
                  "would have modified it (util.cpp:302-318)");
  else if (rc != GB_OK)
    Rcpp::stop(std::string("gauss_b200: ") + gb_status_string(rc) + ": " + gb_last_error(gauss_b200_ctx()));
  for (int64_t i = 0; i < n; i++)                                      // SetZ / SetInfo, dist.cpp:200-202
    if (type[(size_t)i] == 0 && bp[(size_t)i] >= args.start_bp && bp[(size_t)i] <= args.end_bp) {
      snp_vec[(size_t)i]->SetZ(z[(size_t)i]);
      snp_vec[(size_t)i]->SetInfo(info[(size_t)i]);
    }
  Rcpp::Rcout << "Number of measured SNPs: " << n_t << std::endl;      // dist.cpp:214-218
  Rcpp::Rcout << "Number of imputed SNPs: " << n_u << std::endl;
}

void run_distmix(std::vector<Snp*>& snp_vec, Arguments& args) { gauss_b200_run_window(snp_vec, args, true, "DISTMIX"); }
void run_dist(std::vector<Snp*>& snp_vec, Arguments& args) { gauss_b200_run_window(snp_vec, args, false, "DIST"); }
// ---- end patch -------------------------------------------------------------------------------------------------------

// ---- test harness: the same C surface as go_run_window of oracle/ref_glue.cpp -----------------------------------------
extern "C" int go_run_window_patched(const int* type, const long long* bp, double* z, double* info, const char* geno,
                                     int64_t n_snps, const int* m, int n_pops, const double* w, long long start_bp,
                                     long long end_bp, double lambda, double min_abs_eig, int min_measured, int min_unmeasured,
                                     char* err_out, int err_cap) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  Arguments args;
  args.chr = 0;
  args.start_bp = start_bp;
  args.end_bp = end_bp;
  args.lambda = lambda;
  args.min_abs_eig = min_abs_eig;
  args.min_num_measured_snp = min_measured;
  args.min_num_unmeasured_snp = min_unmeasured;
  args.num_samples = (int)N;
  args.num_pops = n_pops;
  for (int p = 0; p < n_pops; p++) {       // every population of this synthetic panel is flagged
    args.ref_pop_size_vec.push_back(m[p]);
    args.pop_flag_vec.push_back(1);
  }
  if (w) args.pop_wgt_vec.assign(w, w + n_pops);
  std::vector<Snp> store((size_t)n_snps);
  std::vector<Snp*> snp_vec;
  for (int64_t i = 0; i < n_snps; i++) {
    Snp& s = store[(size_t)i];
    s.SetBp(bp[i]);
    s.SetType(type[i]);
    s.SetZ(z[i]);
    s.SetInfo(info[i]);
    std::vector<std::string> gv;
    const char* p = geno + i * N;
    for (int k = 0; k < n_pops; k++) {
      gv.emplace_back(p, p + m[k]);
      p += m[k];
    }
    s.SetGenotypeVec(gv);
    snp_vec.push_back(&s);
  }
  try {
    if (w) run_distmix(snp_vec, args);
    else run_dist(snp_vec, args);
  } catch (const std::runtime_error& e) {
    if (err_out && err_cap > 0) snprintf(err_out, (size_t)err_cap, "%s", e.what());
    return 1;
  }
  for (int64_t i = 0; i < n_snps; i++) {
    z[i] = store[(size_t)i].GetZ();
    info[i] = store[(size_t)i].GetInfo();
  }
  return 0;
}
