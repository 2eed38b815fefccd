// Minimal stand-in for <Rcpp.h> (Rcpp is not installed in this image), used ONLY by
// oracle/build_ref.sh to compile the reference's own sources into oracle/_ref/.
// Test infrastructure, not product code.
#ifndef GB_REF_SHIM_RCPP_H
#define GB_REF_SHIM_RCPP_H
#include <cmath>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "Rmath.h"

namespace Rcpp {
struct NullStream : std::ostream {
  struct NullBuf : std::streambuf {
    int overflow(int c) override { return c; }
  } buf;
  NullStream() : std::ostream(&buf) {}
};
inline std::ostream& rcout_instance() {
  static NullStream s;  // the reference prints progress bars; keep the oracle quiet
  return s;
}
#define Rcout rcout_instance()
[[noreturn]] inline void stop(const std::string& msg) { throw std::runtime_error(msg); }
inline std::string& last_warning() {
  static std::string w;
  return w;
}
inline void warning(const std::string& msg) { last_warning() = msg; }

// just enough of NumericVector / NumericMatrix for src/computeLD.cpp:95-116
class NumericVector {
 public:
  void push_back(double v) { d_.push_back(v); }
  double& operator()(size_t i) { return d_[i]; }
  size_t size() const { return d_.size(); }
 private:
  std::vector<double> d_;
};
class NumericMatrix {
 public:
  NumericMatrix() : r_(0), c_(0) {}
  NumericMatrix(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
  double& operator()(size_t i, size_t j) { return d_[j * (size_t)r_ + i]; }  // column-major like R
  int nrow() const { return r_; }
  const double* data() const { return d_.data(); }
 private:
  int r_, c_;
  std::vector<double> d_;
};
}  // namespace Rcpp
#endif
