/* Stand-in for <RcppEigen.h>: see frk_stub_eigen.h (TEST INFRASTRUCTURE). */
#ifndef FRK_STUB_RCPPEIGEN_H
#define FRK_STUB_RCPPEIGEN_H
#include "frk_stub_eigen.h"
#include "Rcpp.h"
#endif
