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
#include <algorithm>
#include <cmath>
#include <limits>
#include <vector>
#include "cmle.cuh"
#include "mrts_fit.cuh"

namespace frk {


namespace {

__global__ void copy_block_kernel(const double* __restrict__ src, int64_t lds, int rows, int cols,
                                  double* __restrict__ dst, int64_t ldd) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows)
        for (int j = blockIdx.y; j < cols; j += gridDim.y) dst[i + int64_t(j) * ldd] = src[i + int64_t(j) * lds];
}

// B[:, j] = A[:, j] * f(j)
__global__ void scale_cols_kernel(const double* __restrict__ A, int n, int cols, const double* __restrict__ f,
                                  double* __restrict__ B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        for (int j = blockIdx.y; j < cols; j += gridDim.y) B[i + int64_t(j) * n] = A[i + int64_t(j) * n] * f[j];
}
__global__ void scale_rows_kernel(double* __restrict__ A, int n, int cols, const double* __restrict__ f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        for (int j = blockIdx.y; j < cols; j += gridDim.y) A[i + int64_t(j) * n] *= f[i];
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
    if (i < n)
        for (int j = blockIdx.y; j < n; j += gridDim.y)
            if (i > j) A[j + int64_t(i) * n] = A[i + int64_t(j) * n];
}

void launch2d(int rows, int cols, dim3& grid, dim3& block) {
    block = dim3(128);
    grid = dim3(unsigned((rows + 127) / 128), unsigned(std::min(cols, 65535)));   // the kernels stride over columns in y
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
    FRK_CUBLAS(cublasSetStream(lh.blas, cudaStreamPerThread));
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
                   bool has_vfixed, double vfixed, CmleOut& out, const std::vector<double>* eigen_asc) {
    LinalgHandles& lh = linalg_handles();
    FRK_CUBLAS(cublasSetStream(lh.blas, cudaStreamPerThread));
    const double one = 1.0, zero = 0.0;
    std::vector<double> asc(K), d(K);
    if (eigen_asc) {                                                // dJSJ already holds P (ascending), see below
        asc = *eigen_asc;
    } else {
        DevBuf<double> dval(static_cast<size_t>(K));
        sym_eigen_dev(dJSJ, K, dval.p);                             // R/autoFRK.R:368 / :441  (P ascending in dJSJ)
        FRK_CUDA(cudaMemcpy(asc.data(), dval.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
    }
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
        FRK_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
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
        FRK_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
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
    FRK_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
}

void cmle_imat_from_stats_dev(const double* dG, int64_t ldg, const double* dC, int64_t ldc, double zz, int K, int T,
                              double n, double s, bool wsave, bool only_loglik, CmleOut& out) {
    if (K < 1 || T < 1) fail(1, "K and T must be positive");
    LinalgHandles& lh = linalg_handles();
    FRK_CUBLAS(cublasSetStream(lh.blas, cudaStreamPerThread));
    const double one = 1.0, zero = 0.0;
    const double trS = zz / T;                                      // R/autoFRK.R:431
    DevBuf<double> half(size_t(K) * K), t(size_t(K) * T), JSJ(size_t(K) * K);
    half_from_gram_dev(dG, ldg, K, half.p);                         // :432
    // ihF %*% Data = half %*% C                                       :433-435
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, T, K, &one, half.p, K, dC, int(ldc), &zero, t.p, K));
    const double invT = 1.0 / T;                                    // tcrossprod(.)/TT; lower triangle == (JSJ+JSJ')/2 :440
    if (T < K) {
        // JSJ = t t'/T has rank <= T: its non-zero eigenpairs come from the T x T matrix S = t't/T = U diag(lam) U'
        // as P_j = t U_j / sqrt(T lam_j); the other K-T eigenvalues are exactly 0 and, with dhat = 0, their
        // eigenvectors never enter M, w, V or L (:376-391, :451-474), so those columns are left zero.
        DevBuf<double> S(size_t(T) * T), lam(static_cast<size_t>(T)), sc(static_cast<size_t>(T));
        FRK_CUBLAS(cublasDsyrk(lh.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, T, K, &invT, t.p, K, &zero, S.p, T));
        sym_eigen_dev(S.p, T, lam.p);                               // ascending
        std::vector<double> lam_h(T), sc_h(T), asc(K, 0.0);
        FRK_CUDA(cudaMemcpy(lam_h.data(), lam.p, sizeof(double) * T, cudaMemcpyDeviceToHost));
        for (int j = 0; j < T; ++j) {
            const double l = std::fmax(lam_h[j], 0.0);
            asc[K - T + j] = l;
            sc_h[j] = (l > 0.0) ? 1.0 / std::sqrt(double(T) * l) : 0.0;
        }
        FRK_CUDA(cudaMemcpy(sc.p, sc_h.data(), sizeof(double) * T, cudaMemcpyHostToDevice));
        FRK_CUDA(cudaMemset(JSJ.p, 0, sizeof(double) * size_t(K) * (K - T)));
        double* Pnz = JSJ.p + size_t(K) * (K - T);
        DevBuf<double> tu(size_t(K) * T);
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, T, T, &one, t.p, K, S.p, T, &zero, tu.p, K));
        dim3 grid, block;
        launch2d(K, T, grid, block);
        scale_cols_kernel<<<grid, block>>>(tu.p, K, T, sc.p, Pnz);
        FRK_CUDA(cudaGetLastError());
        cmle_core_dev(half.p, JSJ.p, dC, ldc, K, T, n, trS, s, 0.0, wsave, only_loglik, false, 0.0, out, &asc);
    } else {
        FRK_CUBLAS(cublasDsyrk(lh.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, K, T, &invT, t.p, K, &zero, JSJ.p, K));
        cmle_core_dev(half.p, JSJ.p, dC, ldc, K, T, n, trS, s, 0.0, wsave, only_loglik, false, 0.0, out);
    }
    if (!only_loglik) {
        out.half = std::move(half);
    }
}

// The K candidates of selectBasis use the leading K columns of one basis, i.e. the leading principal blocks G_K, C_K.
// With G = L L' (Cholesky), L_K is the leading block of L and W_K = L_K^-1 C_K the leading K rows of W = L^-1 C
// (forward substitution never looks ahead), and L_K^-1 = Q_K G_K^(-1/2) for an orthogonal Q_K, so
//   JSJ_K = G_K^(-1/2) C_K C_K' G_K^(-1/2) / T   (R/autoFRK.R:433-435)   and   W_K W_K' / T
// are orthogonally similar: same eigenvalues d, hence the same v (sol.v) and the same likelihood (:441-447).  The
// non-zero eigenvalues are those of the min(K, T)-sized Gram of W_K.  One potrf + one trsm replace 2 x length(Ks)
// K x K eigen-decompositions.
void negloglik_sweep_dev(const double* dG, int64_t ldg, const double* dC, int64_t ldc, double zz, int T, double n, double s,
                         const int* Ks, int nK, double* negloglik) {
    if (nK < 1) return;
    if (T < 1) fail(1, "T must be positive");
    int Kmax = 0;
    for (int i = 0; i < nK; ++i) {
        if (Ks[i] < 1) fail(1, "every K must be positive");
        Kmax = std::max(Kmax, Ks[i]);
    }
    const double nan = std::numeric_limits<double>::quiet_NaN();
    LinalgHandles& lh = linalg_handles();
    FRK_CUBLAS(cublasSetStream(lh.blas, cudaStreamPerThread));
    FRK_CUSOLVER(cusolverDnSetStream(lh.solver, cudaStreamPerThread));
    const double one = 1.0, zero = 0.0;
    const double trS = zz / T;                                       // R/autoFRK.R:431
    DevBuf<double> L(size_t(Kmax) * Kmax), W(size_t(Kmax) * T);
    dim3 grid, block;
    launch2d(Kmax, Kmax, grid, block);
    copy_block_kernel<<<grid, block>>>(dG, ldg, Kmax, Kmax, L.p, Kmax);
    launch2d(Kmax, T, grid, block);
    copy_block_kernel<<<grid, block>>>(dC, ldc, Kmax, T, W.p, Kmax);
    FRK_CUDA(cudaGetLastError());
    std::vector<double> gdiag(Kmax), ldiag(Kmax);
    FRK_CUDA(cudaMemcpy2D(gdiag.data(), sizeof(double), L.p, sizeof(double) * (Kmax + 1), sizeof(double), size_t(Kmax),
                          cudaMemcpyDeviceToHost));
    int lwork = 0;
    FRK_CUSOLVER(cusolverDnDpotrf_bufferSize(lh.solver, CUBLAS_FILL_MODE_LOWER, Kmax, L.p, Kmax, &lwork));
    DevBuf<double> work(static_cast<size_t>(std::max(lwork, 1)));
    DevBuf<int> dinfo(1);
    FRK_CUSOLVER(cusolverDnDpotrf(lh.solver, CUBLAS_FILL_MODE_LOWER, Kmax, L.p, Kmax, work.p, lwork, dinfo.p));
    int info = 0;
    FRK_CUDA(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) {  // not positive definite: nothing of the factor is trusted
        for (int i = 0; i < nK; ++i) negloglik[i] = nan;
        return;
    }
    FRK_CUDA(cudaMemcpy2D(ldiag.data(), sizeof(double), L.p, sizeof(double) * (Kmax + 1), sizeof(double), size_t(Kmax),
                          cudaMemcpyDeviceToHost));
    // largest K whose leading block is safely positive definite: pivots^2 above 1e-10 of the largest diagonal so far
    // (below that the reference's eigenvalue rule, R/helper.R:22-24, and a Cholesky factor may part ways)
    int Ksafe = 0;
    {
        double gmax = 0.0;
        for (int i = 0; i < Kmax; ++i) {
            gmax = std::fmax(gmax, gdiag[i]);
            if (!(ldiag[i] * ldiag[i] > 1e-10 * gmax)) break;
            Ksafe = i + 1;
        }
    }
    FRK_CUBLAS(cublasDtrsm(lh.blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, Kmax, T, &one,
                           L.p, Kmax, W.p, Kmax));
    std::vector<double> w1;
    if (T == 1) {  // rank one: d = (|W_K|^2, 0, ..., 0)
        w1.resize(Kmax);
        FRK_CUDA(cudaMemcpy(w1.data(), W.p, sizeof(double) * Kmax, cudaMemcpyDeviceToHost));
    }
    const int mmax = std::min(Kmax, T);
    DevBuf<double> S(size_t(mmax) * mmax), vals(static_cast<size_t>(mmax));
    DevBuf<double> ework;
    const double invT = 1.0 / T;
    for (int i = 0; i < nK; ++i) {
        const int K = Ks[i];
        if (K > Ksafe) { negloglik[i] = nan; continue; }
        const int m = std::min(K, T);
        std::vector<double> d(K, 0.0);
        if (T == 1) {
            double acc = 0.0;
            for (int j = 0; j < K; ++j) acc += w1[j] * w1[j];
            d[0] = acc;
        } else {
            if (T <= K)
                FRK_CUBLAS(cublasDsyrk(lh.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, m, K, &invT, W.p, Kmax, &zero, S.p, m));
            else
                FRK_CUBLAS(cublasDsyrk(lh.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, m, T, &invT, W.p, Kmax, &zero, S.p, m));
            int lw = 0;
            FRK_CUSOLVER(cusolverDnDsyevd_bufferSize(lh.solver, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, m, S.p, m,
                                                     vals.p, &lw));
            if (ework.n < size_t(lw)) ework.alloc(static_cast<size_t>(lw));
            FRK_CUSOLVER(cusolverDnDsyevd(lh.solver, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, m, S.p, m, vals.p,
                                          ework.p, lw, dinfo.p));
            std::vector<double> asc(m);
            FRK_CUDA(cudaMemcpy(asc.data(), vals.p, sizeof(double) * m, cudaMemcpyDeviceToHost));
            FRK_CUDA(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
            if (info != 0) fail(4, "symmetric eigen decomposition did not converge (info=%d)", info);
            for (int j = 0; j < m; ++j) d[j] = std::fmax(asc[m - 1 - j], 0.0);   // descending; :445
        }
        const double v = sol_v(d, s, trS, n);                        // :444
        negloglik[i] = neg2llik(d, s, v, trS, n) * T;                // :447
    }
}

}  // namespace frk
