// Device-resident "plan" for evaluating an mrts basis: knots, packed eigenvector block W,
// polynomial correction Cpoly = BBBH * W, and the shift / nconst of the polynomial columns.
// One plan serves any number of predict calls (R: one `mrts` object, many predict.mrts()).
#pragma once
#include <vector>
#include "frk_common.cuh"
#include "mrts_predict_kernel.cuh"

namespace frk {

struct PredictSlice {
    int col0 = 0;   // first column of W covered by this slice
    int ncols = 0;  // valid columns (<= kernel BN)
    DevBuf<double> wpack;
    DevBuf<double> cpoly;
};

struct MrtsPlan {
    int device = 0;
    int sm_count = 0;
    int d = 0;
    int64_t m = 0;          // knots
    int k = 0;              // columns of X1 the caller asked for (reference `k`)
    int k_compute = 0;      // leading columns that are not identically zero (quirk: R passes k = K)
    int nchunks = 0;
    const PredictKernelInfo* ki = nullptr;
    std::vector<PredictSlice> slices;
    DevBuf<double> knots;   // nchunks * KC coordinate tuples (shared by all slices)
    double shift[3] = {0, 0, 0};   // colMeans(Xu)        R/autoFRK.R:1462
    double nconst[3] = {1, 1, 1};  // attr(obj,"nconst")  R/autoFRK.R:1464
    bool has_nconst = false;
    int64_t launches = 0;   // kernels launched through this plan (bench bookkeeping)
    // sources kept for derived plans (plan_derive_dev): knots m x d, BBBH (d+1) x m, W = UZ[0:m, 0:k_compute] (ld m)
    DevBuf<double> src_Xu, src_BBBH, src_W;
};

// Inputs are DEVICE pointers, column-major: Xu m x d (ld m), BBBH (d+1) x m (ld d+1), UZ ld = ldu.
// d_cpoly_sub (optional, (d+1) x k, ld d+1): subtracted from Cpoly = BBBH * W (derived plans).  keep_sources: keep
// device copies of Xu, BBBH and W so that plan_derive_dev can be used on the plan.
MrtsPlan* plan_create_dev(const double* dXu, int64_t m, int d, const double* dBBBH, const double* dUZ,
                          int64_t ldu, int k, const double* h_nconst, const double* h_shift,
                          cudaStream_t st, const double* d_cpoly_sub = nullptr, bool keep_sources = true);
// Plan whose X1 output is basis(x) %*% B for a K x q coefficient matrix B (device, ld K; K = the plan's k, rows in
// basis-column order [1, polynomial columns, X1 columns]): the projection W and the polynomial terms are folded
// through B once (W' = W B_rest, m x q), so evaluating basis %*% B costs q columns of K1 instead of K columns plus
// an n2 x K x q product, and the n2 x K basis is never formed.  Used for predict.FRK (R/autoFRK.R:1216,1220).
MrtsPlan* plan_derive_dev(MrtsPlan* p, const double* dB, int q, cudaStream_t st);
void plan_destroy(MrtsPlan* p);

// X1 (n2 x k, ld ldo) on device from xnew (n2 x d, ld ldx) on device.  Columns >= k_compute are zero-filled.
void plan_predict_x1_dev(MrtsPlan* p, const double* dxnew, int64_t n2, int64_t ldx, double* dout, int64_t ldo,
                         cudaStream_t st);
// full basis [1, (x - shift)/nconst, X1[:, 0:k-d-1]]  (R/autoFRK.R:1460-1466): n2 x k, k = K = NCOL(object)
void plan_predict_basis_dev(MrtsPlan* p, const double* dxnew, int64_t n2, int64_t ldx, double* dout,
                            int64_t ldo, cudaStream_t st);

}  // namespace frk
