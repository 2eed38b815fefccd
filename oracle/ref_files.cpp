// ref_files.cpp -- file-level half of oracle/_ref/libgauss_ref.so (test infrastructure, NOT product code).
//
// The reference's own I/O path, compiled UNMODIFIED from /root/reference/src where it lies:
//   bgzf.c  (BGZF reader / writer)  and  gauss.cpp  (Arguments, ReadInputZ 121-190, ReadReferenceIndex 293-399,
//   MakeSnpVecMix 631-693, ReadGenotype 720-785, read_ref_desc 951-993, init_pop_flag_wgt_vec 1093-1117),
// plus BgzfGetLine / FlipGenotypeVec (util.cpp:474-507, extracted at build time).  Two entry points:
//   go_write_bgzf_panel  writes a reference index + data file pair through the reference's bgzf_write, with the real
//                        bgzf_tell virtual offsets in the index's fpos column (SURVEY.md section 4 iv)
//   go_file_distmix      the body of distmix() (distmix.cpp:41-114) on files: read_ref_desc -> init_pop_flag_wgt_vec ->
//                        ReadInputZ -> ReadReferenceIndex -> MakeSnpVecMix -> ReadGenotype -> run_distmix -> output rows
#include <algorithm>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include <RcppEigen.h>
#include "snp.h"
#include "gauss.h"
#include "util.h"
extern "C" {
#include "bgzf.h"
}

#include "gen/util_474_507.inc"   // FlipGenotypeVec, BgzfGetLine

void run_distmix(std::vector<Snp*>& snp_vec, Arguments& args);   // ref_glue.cpp (the reference's own body)

extern "C" {

// geno: [n_snps][N] chars over ALL populations (pop order = m[]), af: [n_snps][n_pops] as they should be printed
// (6 decimals).  Writes data_path (BGZF) and index_path (BGZF) exactly in the reference's formats
// (gauss.cpp:324-330, 572-585).  fpos_out (optional) receives the virtual offset of every data line.
int go_write_bgzf_panel(const char* data_path, const char* index_path, int64_t n_snps, const char* const* rsid,
                        const int* chr, const long long* bp, const char* const* a1, const char* const* a2,
                        const char* geno, const int* m, int n_pops, const double* af, long long* fpos_out) {
  int64_t N = 0;
  for (int p = 0; p < n_pops; p++) N += m[p];
  BGZF* fd = bgzf_open(data_path, "w");
  BGZF* fi = bgzf_open(index_path, "w");
  if (!fd || !fi) return -1;
  std::string line;
  char num[64];
  for (int64_t i = 0; i < n_snps; i++) {
    const long long fpos = bgzf_tell(fd);
    if (fpos_out) fpos_out[i] = fpos;
    line.clear();
    const char* g = geno + i * N;
    for (int p = 0; p < n_pops; p++) {
      line.append(g, (size_t)m[p]);
      line.push_back(' ');
      g += m[p];
    }
    for (int p = 0; p < n_pops; p++) {
      snprintf(num, sizeof(num), "%.6f", af[i * n_pops + p]);
      line.append(num);
      line.push_back(p + 1 < n_pops ? ' ' : '\n');
    }
    if (bgzf_write(fd, line.data(), (int)line.size()) != (int)line.size()) return -2;
    // a SNP's af1ref column: mean of the per-population frequencies (informational, unused by distmix)
    double s = 0;
    for (int p = 0; p < n_pops; p++) s += af[i * n_pops + p];
    line = std::string(rsid[i]) + " " + std::to_string(chr[i]) + " " + std::to_string(bp[i]) + " " + a1[i] + " " + a2[i] + " ";
    snprintf(num, sizeof(num), "%.6f", s / n_pops);
    line += num;
    line += " " + std::to_string(fpos) + "\n";
    if (bgzf_write(fi, line.data(), (int)line.size()) != (int)line.size()) return -2;
  }
  if (bgzf_close(fd) != 0 || bgzf_close(fi) != 0) return -3;
  return 0;
}

// distmix() on files.  Output rows (bp in [start_bp, end_bp], distmix.cpp:100-114) into caller arrays of capacity `cap`;
// strings (rsid, a1, a2) are written newline-separated into str_out.  Returns the number of rows, or a negative code:
// -1 too few SNPs (Rcpp::stop in run_distmix), -2 any other Rcpp::stop, -3 capacity.
int go_file_distmix(const char* input_file, const char* index_file, const char* data_file, const char* pop_desc_file,
                    int chr, long long start_bp, long long end_bp, long long wing, const char* const* wgt_pops,
                    const double* wgt_vals, int n_wgt, double af1_cutoff, int cap, long long* bp_out, double* af1mix_out,
                    double* z_out, double* info_out, int* type_out, char* str_out, int str_cap, char* err_out, int err_cap,
                    int* n_measured, int* n_all) {
  Arguments args;
  args.chr = chr;
  args.start_bp = start_bp;
  args.end_bp = end_bp;
  args.wing_size = wing;
  for (int i = 0; i < n_wgt; i++) {
    std::string pop = wgt_pops[i];
    std::transform(pop.begin(), pop.end(), pop.begin(), ::toupper);   // distmix.cpp:52
    args.pop_wgt_map[pop] = wgt_vals[i];
  }
  args.input_file = input_file;
  args.reference_index_file = index_file;
  args.reference_data_file = data_file;
  args.reference_pop_desc_file = pop_desc_file;
  args.af1_cutoff = af1_cutoff;
  std::map<MapKey, Snp*, LessThanMapKey> snp_map;
  std::vector<Snp*> snp_vec;
  int rc = 0;
  try {
    read_ref_desc(args);
    init_pop_flag_wgt_vec(args);
    ReadInputZ(snp_map, args, false);
    ReadReferenceIndex(snp_map, args);
    MakeSnpVecMix(snp_vec, snp_map, args);
    ReadGenotype(snp_vec, args);
    if (n_all) *n_all = (int)snp_vec.size();
    if (n_measured) {
      *n_measured = 0;
      for (Snp* s : snp_vec) *n_measured += s->GetType() == 1;
    }
    run_distmix(snp_vec, args);
    FreeGenotype(snp_vec);
    std::string strs;
    for (Snp* s : snp_vec) {
      const int bp = s->GetBp();                    // (the reference narrows to int here, distmix.cpp:101)
      if (bp >= start_bp && bp <= end_bp) {
        if (rc >= cap) {
          rc = -3;
          break;
        }
        bp_out[rc] = s->GetBp();
        af1mix_out[rc] = s->GetAf1Mix();
        z_out[rc] = s->GetZ();
        info_out[rc] = s->GetInfo();
        type_out[rc] = s->GetType();
        strs += s->GetRsid() + " " + s->GetA1() + " " + s->GetA2() + "\n";
        rc++;
      }
    }
    if (rc >= 0) {
      if ((int)strs.size() + 1 > str_cap) rc = -3;
      else std::memcpy(str_out, strs.c_str(), strs.size() + 1);
    }
  } catch (const std::runtime_error& e) {
    if (err_out && err_cap > 0) snprintf(err_out, (size_t)err_cap, "%s", e.what());
    rc = std::strstr(e.what(), "Not enough") ? -1 : -2;
  }
  for (auto it = snp_map.begin(); it != snp_map.end(); ++it) delete it->second;
  return rc;
}

}  // extern "C"
