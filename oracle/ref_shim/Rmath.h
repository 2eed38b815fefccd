// Stand-in for <Rmath.h> (R's nmath is not installed in this image).  Test infrastructure, not product code.
// The two functions the reference's gene code calls (gene.cpp:376,509,648,781: R::pnorm5, R::pchisq) are restated from
// their published definitions -- the normal tail through erfc, the chi-square tail as the regularised upper incomplete
// gamma function Q(df/2, x/2) (series for x < a + 1, Lentz continued fraction otherwise) -- and pinned against scipy in
// tests/test_jepeg.py.
#ifndef GB_REF_SHIM_RMATH_H
#define GB_REF_SHIM_RMATH_H
#include <cmath>
namespace R {
inline double pnorm5(double x, double mu, double sigma, int lower_tail, int log_p) {
  const double t = (x - mu) / sigma;
  const double p = lower_tail ? 0.5 * std::erfc(-t / std::sqrt(2.0)) : 0.5 * std::erfc(t / std::sqrt(2.0));
  return log_p ? std::log(p) : p;
}
inline double gb_shim_gammq(double a, double x) {   // Q(a, x)
  if (!(x > 0.0)) return 1.0;
  const double gln = std::lgamma(a);
  if (x < a + 1.0) {                                // P by its series, Q = 1 - P
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 100000; n++) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return 1.0 - sum * std::exp(-x + a * std::log(x) - gln);
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; i++) {
    const double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-16) break;
  }
  return std::exp(-x + a * std::log(x) - gln) * h;
}
inline double pchisq(double x, double df, int lower_tail, int log_p) {
  const double q = gb_shim_gammq(0.5 * df, 0.5 * x);
  const double p = lower_tail ? 1.0 - q : q;
  return log_p ? std::log(p) : p;
}
}  // namespace R
#endif
