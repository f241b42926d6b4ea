// Device-side mrts fit (K2 + K3) and shared dense linear-algebra handles.
#pragma once
#include <cublas_v2.h>
#include <cusolverDn.h>
#include "frk_common.cuh"

namespace frk {

#define FRK_CUBLAS(expr)                                                                       \
    do {                                                                                       \
        cublasStatus_t _s = (expr);                                                            \
        if (_s != CUBLAS_STATUS_SUCCESS) ::frk::fail(3, "cuBLAS error %d at %s:%d", int(_s), __FILE__, __LINE__); \
    } while (0)
#define FRK_CUSOLVER(expr)                                                                     \
    do {                                                                                       \
        cusolverStatus_t _s = (expr);                                                          \
        if (_s != CUSOLVER_STATUS_SUCCESS) ::frk::fail(3, "cuSOLVER error %d at %s:%d", int(_s), __FILE__, __LINE__); \
    } while (0)

struct LinalgHandles {
    int device = -1;
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t solver = nullptr;
};
// per-thread cuBLAS / cuSOLVER handles for the current device (created on first use)
LinalgHandles& linalg_handles();

// Result of the fit, resident on the device (column-major):
//   X     m x (k+d+1)        basis at the knots                      src/autoFRK.cpp:185,209-210
//   UZ    (m+d+1) x (k+d+1)  [gammas | unit | xobs_diag] block layout  :212-215
//   BBBH  (d+1) x m          BBB * H                                   :240
//   nconst d                 column norms / sqrt(m)                    :207,216
struct MrtsFitDev {
    int64_t m = 0;
    int d = 0, k = 0;
    DevBuf<double> Xu, X, UZ, BBBH;
    double nconst[3] = {1, 1, 1};
    double shift[3] = {0, 0, 0};
};

// hXu (m x d) and h_xobs_diag (d x d) are HOST pointers (they are tiny); everything m x m runs on the device.
void mrts_fit_dev(const double* hXu, const double* h_xobs_diag, int64_t m, int d, int k, MrtsFitDev& out);

// full symmetric eigendecomposition in place (ascending values; vectors overwrite dA): getASCeigens,
// src/autoFRK.cpp:24-32
void sym_eigen_dev(double* dA, int n, double* dvalues);

}  // namespace frk
