// Stand-in for <RcppEigen.h>: see Eigen/Dense in this directory.
#ifndef GB_REF_SHIM_RCPPEIGEN_H
#define GB_REF_SHIM_RCPPEIGEN_H
#include "Rcpp.h"
#include "Eigen/Dense"
#endif
