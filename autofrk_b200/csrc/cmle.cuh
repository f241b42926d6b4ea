// K5: K x K stage of the FRK fit on the device (see cmle.cu).
#pragma once
#include <vector>
#include "frk_common.cuh"

namespace frk {

struct CmleOut {
    double v = 0, s = 0, negloglik = 0;
    int K = 0, T = 0;
    int nL = 0;              // columns of L = F * LA[:, 0:nL]  (R/autoFRK.R:387-391)
    DevBuf<double> M, w, V;  // K x K, K x T, K x K (column-major)
    DevBuf<double> half;     // G^(-1/2)
    DevBuf<double> LA;       // K x K: half P diag(sqrt(dhat)), columns in R's (descending) order
};

double sol_v(const std::vector<double>& d, double s, double trS, double n);
double neg2llik(const std::vector<double>& d, double s, double v, double trS, double n);

// half = G^(-1/2) with zero eigenvalues -> 0 (getHalf, R/helper.R:19-26); dG is K x K with leading dimension ldg
void half_from_gram_dev(const double* dG, int64_t ldg, int K, double* dhalf);

// cMLE (R/autoFRK.R:333-397) given half and JSJ on the device (JSJ is destroyed).  eigen_asc != nullptr: dJSJ already
// holds JSJ's eigenvectors (columns in ascending eigenvalue order) and *eigen_asc the K eigenvalues.
void cmle_core_dev(const double* dhalf, double* dJSJ, const double* dC, int64_t ldc, int K, int T, double n, double trS,
                   double s, double ldet, bool wsave, bool only_loglik, bool has_vfixed, double vfixed, CmleOut& out,
                   const std::vector<double>* eigen_asc = nullptr);

// cMLEimat (R/autoFRK.R:399-480) from the sufficient statistics G (ld ldg), C (ld ldc), zz
void cmle_imat_from_stats_dev(const double* dG, int64_t ldg, const double* dC, int64_t ldc, double zz, int K, int T,
                              double n, double s, bool wsave, bool only_loglik, CmleOut& out);

// selectBasis' AIC sweep (R/autoFRK.R:305-317): negloglik of cMLEimat(Fk[, 1:K], Data, s, onlylogLike) for every K in
// Ks from ONE Cholesky factor of the max-K Gram matrix.  Entries the factor cannot serve (G not safely positive
// definite up to that K) are set to NaN -- the caller then takes the eigen path (cmle_imat_from_stats_dev) for them.
void negloglik_sweep_dev(const double* dG, int64_t ldg, const double* dC, int64_t ldc, double zz, int T, double n, double s,
                         const int* Ks, int nK, double* negloglik);

}  // namespace frk
