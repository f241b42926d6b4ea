// K5 -- the small K x K stage of the fixed-rank-kriging fit, on the device, as a function of the
// sufficient statistics (G = F'F, C = F'Z, zz = sum Z^2) that K4 produces:
//   getHalf / getEigen / getASCeigens   R/helper.R:12-26, src/autoFRK.cpp:24-32   (eigen of G, G^(-1/2))
//   cMLEimat                            R/autoFRK.R:399-480
//   cMLE                                R/autoFRK.R:333-397  (given half and JSJ)
// ihF %*% Data (R/autoFRK.R:433-435) equals half %*% (F'Z), so JSJ = (half C)(half C)'/T needs no pass
// over the n rows; and with L'L = diag(dhat) the wSave products (:459-474, invCz :845-852) collapse to
//   w = hP diag(dhat/(s+v+dhat)) hP' C,   V = hP diag(dhat (s+v)/(s+v+dhat)) hP',   hP = half P
// (checked against the literal sequence in tests/test_oracle.py).  The two eigensolves are cuSOLVER
// syevd, the K x K products plain cuBLAS GEMMs; the O(K) scalar rule for v (sol.v, :335-351) runs on
// the host on the K eigenvalues.
#include <cmath>
#include <limits>
#include <vector>
#include "cmle.cuh"
#include "mrts_fit.cuh"

namespace frk {

#define FRK_CUBLAS(expr)                                                                       \
    do {                                                                                       \
        cublasStatus_t _s = (expr);                                                            \
        if (_s != CUBLAS_STATUS_SUCCESS) ::frk::fail(3, "cuBLAS error %d at %s:%d", int(_s), __FILE__, __LINE__); \
    } while (0)

namespace {

__global__ void copy_block_kernel(const double* __restrict__ src, int64_t lds, int rows, int cols,
                                  double* __restrict__ dst, int64_t ldd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < rows && j < cols) dst[i + int64_t(j) * ldd] = src[i + int64_t(j) * lds];
}

// B[:, j] = A[:, j] * f(j)
__global__ void scale_cols_kernel(const double* __restrict__ A, int n, int cols, const double* __restrict__ f,
                                  double* __restrict__ B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < cols) B[i + int64_t(j) * n] = A[i + int64_t(j) * n] * f[j];
}
__global__ void scale_rows_kernel(double* __restrict__ A, int n, int cols, const double* __restrict__ f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < cols) A[i + int64_t(j) * n] *= f[i];
}
// R/helper.R:22-24: sroot = sqrt(pmax(value,0)); sroot[sroot==0] = Inf; 1/sroot
__global__ void inv_sqrt_kernel(const double* __restrict__ lam, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = sqrt(fmax(lam[i], 0.0));
    out[i] = (s == 0.0) ? 0.0 : 1.0 / s;
}
__global__ void mirror_lower_kernel(double* __restrict__ A, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < n && i > j) A[j + int64_t(i) * n] = A[i + int64_t(j) * n];
}

void launch2d(int rows, int cols, dim3& grid, dim3& block) {
    block = dim3(128);
    grid = dim3(unsigned((rows + 127) / 128), unsigned(cols));
}

}  // namespace

// R/autoFRK.R:335-351
double sol_v(const std::vector<double>& d, double s, double trS, double n) {
    const int k = int(d.size());
    double dmax = -std::numeric_limits<double>::infinity();
    for (double x : d) dmax = std::fmax(dmax, x);
    if (dmax < std::fmax(trS / n, s)) return std::fmax(trS / n - s, 0.0);
    int L = 0;  // 1-based max(which(pick))
    double cum = 0.0, cumL = 0.0;
    for (int i = 1; i <= k; ++i) {
        cum += d[i - 1];
        double ks = double(i);
        if (double(k) == n && i == k) ks = n - 1.0;
        if (d[i - 1] > (trS - cum) / (n - ks)) { L = i; cumL = cum; }
    }
    if (L == 0) return std::numeric_limits<double>::quiet_NaN();  // max(integer(0)) = -Inf in R -> error there
    double Ld = double(L);
    if (Ld == n) {
        Ld = n - 1.0;
        cumL = 0.0;
        for (int i = 0; i < L - 1; ++i) cumL += d[i];
    }
    return std::fmax((trS - cumL) / (n - Ld) - s, 0.0);
}

// R/autoFRK.R:353-362
double neg2llik(const std::vector<double>& d, double s, double v, double trS, double n) {
    const int k = int(d.size());
    double mx = -std::numeric_limits<double>::infinity();
    for (double x : d) mx = std::fmax(mx, std::fmax(x - s - v, 0.0) / (s + v));
    if (mx > 1e20) return std::numeric_limits<double>::infinity();
    double sumlog = 0.0, sumq = 0.0;
    for (double x : d) {
        const double eta = std::fmax(x - s - v, 0.0);
        sumlog += std::log(eta + s + v);
        sumq += x * eta / (eta + s + v);
    }
    return n * std::log(2.0 * M_PI) + sumlog + std::log(s + v) * (n - k) + 1.0 / (s + v) * trS - 1.0 / (s + v) * sumq;
}

void half_from_gram_dev(const double* dG, int64_t ldg, int K, double* dhalf) {
    LinalgHandles& lh = linalg_handles();
    FRK_CUBLAS(cublasSetStream(lh.blas, nullptr));
    DevBuf<double> U(size_t(K) * K), Us(size_t(K) * K), lam(static_cast<size_t>(K)), inv(static_cast<size_t>(K));
    dim3 grid, block;
    launch2d(K, K, grid, block);
    copy_block_kernel<<<grid, block>>>(dG, ldg, K, K, U.p, K);
    FRK_CUDA(cudaGetLastError());
    sym_eigen_dev(U.p, K, lam.p);                                   // R/helper.R:13 (ascending; order is irrelevant below)
    inv_sqrt_kernel<<<(K + 127) / 128, 128>>>(lam.p, K, inv.p);     // :22-24
    scale_cols_kernel<<<grid, block>>>(U.p, K, K, inv.p, Us.p);
    FRK_CUDA(cudaGetLastError());
    const double one = 1.0, zero = 0.0;                             // :25  vector %*% (sroot * t(vector))
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_T, K, K, K, &one, Us.p, K, U.p, K, &zero, dhalf, K));
}

void cmle_core_dev(const double* dhalf, double* dJSJ /* K x K, lower triangle valid; destroyed */, const double* dC,
                   int64_t ldc, int K, int T, double n, double trS, double s, double ldet, bool wsave, bool only_loglik,
                   bool has_vfixed, double vfixed, CmleOut& out) {
    LinalgHandles& lh = linalg_handles();
    FRK_CUBLAS(cublasSetStream(lh.blas, nullptr));
    const double one = 1.0, zero = 0.0;
    DevBuf<double> dval(static_cast<size_t>(K));
    sym_eigen_dev(dJSJ, K, dval.p);                                 // R/autoFRK.R:368 / :441  (P ascending in dJSJ)
    std::vector<double> asc(K), d(K);
    FRK_CUDA(cudaMemcpy(asc.data(), dval.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
    for (int i = 0; i < K; ++i) d[i] = asc[K - 1 - i];              // R's eigen(): descending
    const double v = has_vfixed ? vfixed : sol_v(d, s, trS, n);     // :371-375 / :444
    std::vector<double> dii(K);
    for (int i = 0; i < K; ++i) dii[i] = std::fmax(d[i], 0.0);      // :376 / :445
    out.v = v;
    out.s = s;
    out.negloglik = neg2llik(dii, s, v, trS, n) * T + ldet * T;     // :379-380 / :447
    out.K = K;
    out.T = T;
    if (only_loglik) return;

    // coefficients in the solver's ascending order
    std::vector<double> sq(K), cw(K), cv(K);
    for (int i = 0; i < K; ++i) {
        const double dh = std::fmax(dii[K - 1 - i] - s - v, 0.0);   // sol.eta :352
        sq[i] = std::sqrt(dh);
        const double den = s + v + dh;
        cw[i] = (dh > 0.0) ? dh / den : 0.0;
        cv[i] = (dh > 0.0) ? std::sqrt(dh * (s + v) / den) : 0.0;
    }
    DevBuf<double> hP(size_t(K) * K), tmp(size_t(K) * K), coef(static_cast<size_t>(K));
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, K, K, &one, dhalf, K, dJSJ, K, &zero, hP.p, K));
    dim3 grid, block;
    launch2d(K, K, grid, block);
    // M = half P diag(dhat) P' half                                   :382 / :451
    FRK_CUDA(cudaMemcpy(coef.p, sq.data(), sizeof(double) * K, cudaMemcpyHostToDevice));
    scale_cols_kernel<<<grid, block>>>(hP.p, K, K, coef.p, tmp.p);
    out.M.alloc(size_t(K) * K);
    FRK_CUBLAS(cublasDsyrk(lh.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, K, K, &one, tmp.p, K, &zero, out.M.p, K));
    mirror_lower_kernel<<<grid, block>>>(out.M.p, K);
    FRK_CUDA(cudaGetLastError());
    if (!wsave) {
        FRK_CUDA(cudaDeviceSynchronize());
        return;
    }
    // LA = half P diag(sqrt(dhat)) in R's descending column order: L = Fk %*% LA[, dhat > 0]   :387-391
    {
        out.LA.alloc(size_t(K) * K);
        int cnt = 0;
        for (int i = 0; i < K; ++i) cnt += (sq[i] > 0.0);
        out.nL = cnt > 0 ? cnt : 1;  // all(dhat == 0): dhat[1] <- 1e-10 keeps one (zero) column  :388-390
        for (int c = 0; c < K; ++c)  // tmp holds hP diag(sq) in ascending order: reverse the columns
            FRK_CUDA(cudaMemcpyAsync(out.LA.p + size_t(c) * K, tmp.p + size_t(K - 1 - c) * K, sizeof(double) * K,
                                     cudaMemcpyDeviceToDevice, nullptr));
        FRK_CUDA(cudaDeviceSynchronize());
    }
    // V = hP diag(dhat (s+v)/(s+v+dhat)) hP'                         closed form of :470-474
    FRK_CUDA(cudaMemcpy(coef.p, cv.data(), sizeof(double) * K, cudaMemcpyHostToDevice));
    scale_cols_kernel<<<grid, block>>>(hP.p, K, K, coef.p, tmp.p);
    out.V.alloc(size_t(K) * K);
    FRK_CUBLAS(cublasDsyrk(lh.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, K, K, &one, tmp.p, K, &zero, out.V.p, K));
    mirror_lower_kernel<<<grid, block>>>(out.V.p, K);
    // w = hP diag(dhat/(s+v+dhat)) hP' C                              closed form of :464-469
    if (T > 0 && dC) {
        DevBuf<double> pc(size_t(K) * T);
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_T, CUBLAS_OP_N, K, T, K, &one, hP.p, K, dC, int(ldc), &zero, pc.p, K));
        FRK_CUDA(cudaMemcpy(coef.p, cw.data(), sizeof(double) * K, cudaMemcpyHostToDevice));
        dim3 g2, b2;
        launch2d(K, T, g2, b2);
        scale_rows_kernel<<<g2, b2>>>(pc.p, K, T, coef.p);
        out.w.alloc(size_t(K) * T);
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, T, K, &one, hP.p, K, pc.p, K, &zero, out.w.p, K));
    }
    FRK_CUDA(cudaGetLastError());
    FRK_CUDA(cudaDeviceSynchronize());
}

void cmle_imat_from_stats_dev(const double* dG, int64_t ldg, const double* dC, int64_t ldc, double zz, int K, int T,
                              double n, double s, bool wsave, bool only_loglik, CmleOut& out) {
    if (K < 1 || T < 1) fail(1, "K and T must be positive");
    LinalgHandles& lh = linalg_handles();
    FRK_CUBLAS(cublasSetStream(lh.blas, nullptr));
    const double one = 1.0, zero = 0.0;
    const double trS = zz / T;                                      // R/autoFRK.R:431
    DevBuf<double> half(size_t(K) * K), t(size_t(K) * T), JSJ(size_t(K) * K);
    half_from_gram_dev(dG, ldg, K, half.p);                         // :432
    // ihF %*% Data = half %*% C                                       :433-435
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, T, K, &one, half.p, K, dC, int(ldc), &zero, t.p, K));
    const double invT = 1.0 / T;                                    // tcrossprod(.)/TT; lower triangle == (JSJ+JSJ')/2 :440
    FRK_CUBLAS(cublasDsyrk(lh.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, K, T, &invT, t.p, K, &zero, JSJ.p, K));
    cmle_core_dev(half.p, JSJ.p, dC, ldc, K, T, n, trS, s, 0.0, wsave, only_loglik, false, 0.0, out);
    if (!only_loglik) {
        out.half = std::move(half);
    }
}

}  // namespace frk
