#!/usr/bin/env bash
# Builds oracle/_ref/libgauss_ref.so from the reference's OWN sources where they lie
# (/root/reference/src).  Test infrastructure only.  Outputs go to oracle/_ref/ (git-ignored,
# NOT gpurun-ignored, so the prebuilt .so travels to the GPU box).  No reference source is
# copied into the tracked tree: the hot-path functions are extracted by line range into
# oracle/_ref/gen/*.inc at build time.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GAUSS_REFERENCE_DIR:-/root/reference}"
SRC="$REF/src"
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then
  echo "build_ref: $SRC not present; keeping any prebuilt $OUT/libgauss_ref.so" >&2
  exit 0
fi
mkdir -p "$OUT/gen"
cut_lines() { sed -n "$2,$3p" "$SRC/$1" > "$OUT/gen/$4"; }
cut_lines util.cpp 49 70 util_49_70.inc            # CalCor(vector<string>, vector<string>)
cut_lines util.cpp 103 124 util_103_124.inc        # CalWgtCov
cut_lines util.cpp 153 169 util_153_169.inc        # CalCor(std::string&, std::string&)  (zmix)
cut_lines util.cpp 284 296 util_284_296.inc        # CnvrtCovToCor
cut_lines util.cpp 474 507 util_474_507.inc        # FlipGenotypeVec, BgzfGetLine (needed by gauss.cpp's readers)
cut_lines dist.cpp 129 227 dist_129_227.inc        # run_dist
cut_lines distmix.cpp 138 253 distmix_138_253.inc  # run_distmix
cut_lines computeLD.cpp 95 116 computeLD_95_116.inc
cut_lines qcat.cpp 134 262 qcat_134_262.inc          # run_qcat
cut_lines gene.cpp 569 586 gene_569_586.inc          # CorG of Gene::CalJepegmixPval (mixture)
cut_lines gene.cpp 305 315 gene_305_314.inc          # CorG of Gene::CalJepegPval (pooled CalCor)
cut_lines qcatmix.cpp 145 286 qcatmix_145_286.inc    # run_qcatmix
# guard: the extraction must start/end on the expected function boundaries
grep -q '^double CalCor(std::vector<std::string>& x, std::vector<std::string>& y){' "$OUT/gen/util_49_70.inc"
grep -q '^double CalWgtCov(' "$OUT/gen/util_103_124.inc"
grep -q '^double CalCor(std::string& x, std::string& y){' "$OUT/gen/util_153_169.inc"
grep -q '^void run_dist(' "$OUT/gen/dist_129_227.inc"
grep -q '^void run_distmix(' "$OUT/gen/distmix_138_253.inc"
grep -q '^void run_qcat(' "$OUT/gen/qcat_134_262.inc"
grep -q 'Eigen::VectorXd SNP_STD_VEC' "$OUT/gen/gene_569_586.inc"
grep -q 'CorG(i, i) = 1.0 + lambda_;' "$OUT/gen/gene_305_314.inc"
grep -q '^void run_qcatmix(' "$OUT/gen/qcatmix_145_286.inc"
grep -q '^void CnvrtCovToCor(' "$OUT/gen/util_284_296.inc"
grep -q '^void FlipGenotypeVec(' "$OUT/gen/util_474_507.inc"
grep -q '^int BgzfGetLine(BGZF\* fp, std::string& line){' "$OUT/gen/util_474_507.inc"
CXXFLAGS="-O2 -fPIC -ffp-contract=off -w -I$HERE/ref_shim -I$SRC -I$HERE -I$OUT"
gcc -O2 -fPIC -ffp-contract=off -c "$HERE/gauss_oracle.c" -o "$OUT/gauss_oracle_int.o" \
    -Dgo_make_pos_def=gor_make_pos_def -Dgo_inv_full_piv_lu=gor_inv_full_piv_lu \
    -Dgo_cal_cor=gor_cal_cor -Dgo_cal_wgt_cov=gor_cal_wgt_cov -Dgo_run_window=gor_run_window \
    -Dgo_compute_ld=gor_compute_ld -Dgo_last_sample_pairs=gor_last_sample_pairs \
    -Dgo_gram_counts=gor_gram_counts -Dgo_zmix_pairs=gor_zmix_pairs -Dgo_cal_cor_pop=gor_cal_cor_pop \
    -Dgo_run_qcat=gor_run_qcat -Dgo_count_pc=gor_count_pc -Dgo_sym_eig=gor_sym_eig -Dgo_args_default=gor_args_default
g++ $CXXFLAGS -c "$SRC/snp.cpp" -o "$OUT/snp.o"
# the reference's I/O half, unmodified: BGZF reader/writer and gauss.cpp (Arguments, Read*, MakeSnpVec*, ReadGenotype, ...)
gcc -O2 -fPIC -w -I"$SRC" -c "$SRC/bgzf.c" -o "$OUT/bgzf.o"
g++ $CXXFLAGS -c "$SRC/gauss.cpp" -o "$OUT/gauss.o"
g++ $CXXFLAGS -c "$SRC/gene.cpp" -o "$OUT/gene.o"               # jepeg / jepegmix per-gene code, unmodified
g++ $CXXFLAGS -c "$HERE/ref_glue.cpp" -o "$OUT/ref_glue.o"
g++ $CXXFLAGS -c "$HERE/ref_files.cpp" -o "$OUT/ref_files.o"
printf '#include <RcppEigen.h>\n#include "util.h"\n#include "gen/util_474_507.inc"\n' > "$OUT/gen/util_io.cpp"
g++ $CXXFLAGS -c "$OUT/gen/util_io.cpp" -o "$OUT/util_io.o"     # BgzfGetLine / FlipGenotypeVec alone, for the patched library
g++ -shared -o "$OUT/libgauss_ref.so" "$OUT/ref_glue.o" "$OUT/ref_files.o" "$OUT/gene.o" "$OUT/gauss.o" "$OUT/bgzf.o" "$OUT/snp.o" \
    "$OUT/gauss_oracle_int.o" -lz -lm
echo "build_ref: wrote $OUT/libgauss_ref.so"
# the Rcpp-side patch of INTEGRATION.md, compiled over the reference's own Snp / Arguments and linked with the product
LIBGB="$HERE/../gauss_b200/lib/libgauss_b200.so"
if [ -f "$LIBGB" ]; then
  g++ $CXXFLAGS -I"$HERE/../include" -c "$HERE/patch/run_window_patched.cpp" -o "$OUT/run_window_patched.o"
  g++ -shared -o "$OUT/libgauss_patched.so" "$OUT/run_window_patched.o" "$OUT/gauss.o" "$OUT/bgzf.o" "$OUT/snp.o" "$OUT/util_io.o" \
      -L"$HERE/../gauss_b200/lib" -lgauss_b200 -Wl,-rpath,'$ORIGIN/../../gauss_b200/lib' -lz -lm
  echo "build_ref: wrote $OUT/libgauss_patched.so"
fi
