/* Internal to oracle/_ref build (test infrastructure): lets ref_glue.cpp borrow the
 * restated Eigen algorithms under non-clashing names while exporting the go_* API itself. */
#ifndef GAUSS_ORACLE_INTERNAL_H
#define GAUSS_ORACLE_INTERNAL_H
#include "gauss_oracle.h"
int gor_make_pos_def(double *A, int n, double min_abs_eig);
void gor_inv_full_piv_lu(double *inv, const double *A, int n);
int gor_count_pc(const double *A, int n, double eig_cutoff);
#endif
