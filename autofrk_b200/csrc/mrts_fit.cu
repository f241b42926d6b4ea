// K2 + K3 -- the mrts fit on the device (replaces the C++ `mrts()` at src/autoFRK.cpp:160-217):
//   K2  knot kernel matrix H (tpm2, :60-103 + :189), double centering AHA = (I-P) H (I-P) (:197-198)
//   K3  k largest eigenpairs of AHA (mrtseigencpp :39-54: Spectra Lanczos -> cuSOLVER syevdx on the
//       same matrix; ncv ~ m at every config, so Lanczos degenerates to a dense solve anyway)
//   assembly of X, UZ, BBBH, nconst (:205-216, :240)
// The m x m products are plain library GEMMs (cuBLAS); the (d+1) x (d+1) normal equations
// (:192-196) are solved on the host -- 16 numbers.
#include <cmath>
#include <mutex>
#include <vector>
#include <cublas_v2.h>
#include <cusolverDn.h>
#include "frk_common.cuh"
#include "mrts_fit.cuh"
#include "mrts_predict_kernel.cuh"  // tps_phi

namespace frk {


LinalgHandles& linalg_handles() {
    static thread_local LinalgHandles h;
    int dev = 0;
    FRK_CUDA(cudaGetDevice(&dev));
    if (h.device != dev) {
        if (h.blas) cublasDestroy(h.blas);
        if (h.solver) cusolverDnDestroy(h.solver);
        h.blas = nullptr;
        h.solver = nullptr;
        FRK_CUBLAS(cublasCreate(&h.blas));
        FRK_CUSOLVER(cusolverDnCreate(&h.solver));
        h.device = dev;
    }
    return h;
}

namespace {

// H[i,j] = phi_d(|x_i - x_j|), zero diagonal.  Bitwise symmetric: dx -> -dx only changes a sign that is
// squared (d=2,3) or abs'ed (d=1).  Xu is m x d column-major; coordinates are staged as DP-tuples.
template <int D>
__global__ void knot_matrix_kernel(const double* __restrict__ Xu, int64_t m, double* __restrict__ H) {
    constexpr int DP = (D == 3) ? 4 : D;
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t j = blockIdx.y;
    if (i >= m) return;
    __shared__ __align__(16) double uj[4];
    if (threadIdx.x < DP) uj[threadIdx.x] = (threadIdx.x < D) ? Xu[j + threadIdx.x * m] : 0.0;
    __syncthreads();
    __align__(16) double xi[DP];
#pragma unroll
    for (int c = 0; c < DP; ++c) xi[c] = (c < D) ? Xu[i + c * m] : 0.0;
    H[i + j * m] = (i == j) ? 0.0 : tps_phi<D>(xi, uj);
}

// A -= U * V   with U m x r (ld m), V r x m (ld r), r = d+1 <= 4: the rank-(d+1) centering updates
__global__ void rank_update_kernel(double* __restrict__ A, int64_t m, const double* __restrict__ U,
                                   const double* __restrict__ V, int r) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int64_t j = blockIdx.y;
    if (i >= m) return;
    double s = 0.0;
    for (int a = 0; a < r; ++a) s = fma(U[i + a * m], V[a + j * r], s);
    A[i + j * m] -= s;
}

// per-column sign so that the largest-|.| entry is positive; also reverses ascending -> descending
// order while copying the k eigenvectors out of the solver's workspace.
__global__ void gather_eigvecs_kernel(const double* __restrict__ V, int64_t m, int k, double* __restrict__ gamma) {
    const int c = blockIdx.x;  // output column (descending order)
    const double* src = V + int64_t(k - 1 - c) * m;
    __shared__ double smax[256];
    __shared__ int64_t sidx[256];
    double best = -1.0;
    int64_t bi = 0;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) {
        const double a = fabs(src[i]);
        if (a > best) { best = a; bi = i; }
    }
    smax[threadIdx.x] = best;
    sidx[threadIdx.x] = bi;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const double o = smax[threadIdx.x + s];
            const int64_t oi = sidx[threadIdx.x + s];
            if (o > smax[threadIdx.x] || (o == smax[threadIdx.x] && oi < sidx[threadIdx.x])) {
                smax[threadIdx.x] = o;
                sidx[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    const double sgn = (src[sidx[0]] < 0.0) ? -1.0 : 1.0;
    for (int64_t i = threadIdx.x; i < m; i += blockDim.x) gamma[i + int64_t(c) * m] = sgn * src[i];
}

// X (m x (k+d+1)) and UZ ((m+d+1) x (k+d+1)):  src/autoFRK.cpp:205-215
__global__ void assemble_kernel(const double* __restrict__ gamma, const double* __restrict__ Bg /* (d+1) x k = BBB*gamma */,
                                const double* __restrict__ rho_asc, const double* __restrict__ Xu,
                                const double* __restrict__ xobs_diag, int64_t m, int d, int k, double root,
                                double mean0, double mean1, double mean2, double nc0, double nc1, double nc2,
                                double* __restrict__ X, double* __restrict__ UZ) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int c = blockIdx.y;  // column of X / UZ
    const int K = k + d + 1;
    const int64_t ldu = m + d + 1;
    if (i >= ldu) return;
    const double mean[3] = {mean0, mean1, mean2};
    const double nc[3] = {nc0, nc1, nc2};
    // ---- X ----
    if (i < m) {
        double v;
        if (c == 0) v = 1.0;                                                    // :185
        else if (c <= d) v = (Xu[i + int64_t(c - 1) * m] - mean[c - 1]) * (root / nc[c - 1]);  // :209
        else v = gamma[i + int64_t(c - d - 1) * m] * root;                      // :210
        X[i + int64_t(c) * m] = v;
    }
    // ---- UZ ----
    double u = 0.0;                                                             // :212
    if (i < m) {
        if (c < k) {
            // gammas = (gamma - B (BBB gamma)) / rho * root                     // :205
            double bg = Bg[0 + int64_t(c) * (d + 1)];
            for (int a = 0; a < d; ++a) bg = fma(Xu[i + int64_t(a) * m], Bg[a + 1 + int64_t(c) * (d + 1)], bg);
            u = (gamma[i + int64_t(c) * m] - bg) / rho_asc[k - 1 - c] * root;
        }
    } else if (i == m) {
        if (c == k) u = 1.0;                                                    // :214
    } else {
        const int a = int(i - m - 1);
        if (c > k) u = xobs_diag[a + (c - k - 1) * d];                           // :215
    }
    UZ[i + int64_t(c) * ldu] = u;
}

}  // namespace

void mrts_fit_dev(const double* hXu, const double* h_xobs_diag, int64_t m, int d, int k, MrtsFitDev& out) {
    // the reference's checks and messages, src/autoFRK.cpp:173-181
    if (k <= 0) fail(1, "k must be positive");
    if (m <= 1) fail(1, "n must be greater than 1");
    if (k >= m) fail(1, "k must be less than n (got k=%d, n=%d)", k, int(m));
    if (d < 1 || d > 3) fail(1, "Unsupported dimension d: %d", d);  // :64-66

    const double root = std::sqrt(double(m));  // :170
    const int r = d + 1;
    LinalgHandles& lh = linalg_handles();
    cudaStream_t st = nullptr;
    FRK_CUBLAS(cublasSetStream(lh.blas, st));
    FRK_CUSOLVER(cusolverDnSetStream(lh.solver, st));

    // ---- host: B'B, its Cholesky, BBB = (B'B)^{-1} B'  (:190-196); column means / norms (:206-207) ----
    std::vector<double> BtB(r * r, 0.0), BBB(size_t(r) * m);
    for (int a = 0; a < r; ++a)
        for (int b = 0; b <= a; ++b) {
            double s = 0.0;
            for (int64_t j = 0; j < m; ++j) {
                const double ba = a ? hXu[j + int64_t(a - 1) * m] : 1.0;
                const double bb = b ? hXu[j + int64_t(b - 1) * m] : 1.0;
                s += ba * bb;
            }
            BtB[a + b * r] = BtB[b + a * r] = s;
        }
    std::vector<double> Lc(r * r, 0.0);  // lower Cholesky factor
    for (int j = 0; j < r; ++j) {
        double s = BtB[j + j * r];
        for (int p = 0; p < j; ++p) s -= Lc[j + p * r] * Lc[j + p * r];
        if (!(s > 0.0)) fail(1, "knot design matrix [1, Xu] is rank deficient (B'B not positive definite)");
        Lc[j + j * r] = std::sqrt(s);
        for (int i = j + 1; i < r; ++i) {
            double t = BtB[i + j * r];
            for (int p = 0; p < j; ++p) t -= Lc[i + p * r] * Lc[j + p * r];
            Lc[i + j * r] = t / Lc[j + j * r];
        }
    }
    for (int64_t j = 0; j < m; ++j) {
        double y[4];
        for (int i = 0; i < r; ++i) {  // forward substitution L y = B'[:, j]
            double t = i ? hXu[j + int64_t(i - 1) * m] : 1.0;
            for (int p = 0; p < i; ++p) t -= Lc[i + p * r] * y[p];
            y[i] = t / Lc[i + i * r];
        }
        for (int i = r - 1; i >= 0; --i) {  // back substitution L' x = y
            double t = y[i];
            for (int p = i + 1; p < r; ++p) t -= Lc[p + i * r] * BBB[p + j * r];
            BBB[i + j * r] = t / Lc[i + i * r];
        }
    }
    double mean[3] = {0, 0, 0}, nc[3] = {1, 1, 1};
    for (int c = 0; c < d; ++c) {
        double s = 0.0;
        for (int64_t j = 0; j < m; ++j) s += hXu[j + int64_t(c) * m];
        mean[c] = s / double(m);
        double q = 0.0;
        for (int64_t j = 0; j < m; ++j) {
            const double t = hXu[j + int64_t(c) * m] - mean[c];
            q += t * t;
        }
        nc[c] = std::sqrt(q);                         // :207 (undivided; divided by root at :216)
    }

    // ---- device ----
    out.m = m; out.d = d; out.k = k;
    out.Xu.alloc(size_t(m) * d);
    FRK_CUDA(cudaMemcpyAsync(out.Xu.p, hXu, sizeof(double) * m * d, cudaMemcpyHostToDevice, st));
    DevBuf<double> dBBB(size_t(r) * m), dB(size_t(m) * r), dH(size_t(m) * m), dHB(size_t(m) * r), dBtAH(size_t(r) * m);
    DevBuf<double> dxd(size_t(d) * d);
    FRK_CUDA(cudaMemcpyAsync(dBBB.p, BBB.data(), sizeof(double) * r * m, cudaMemcpyHostToDevice, st));
    FRK_CUDA(cudaMemcpyAsync(dxd.p, h_xobs_diag, sizeof(double) * d * d, cudaMemcpyHostToDevice, st));
    {
        std::vector<double> hB(size_t(m) * r);
        for (int64_t j = 0; j < m; ++j) hB[j] = 1.0;
        for (int c = 0; c < d; ++c)
            for (int64_t j = 0; j < m; ++j) hB[j + int64_t(c + 1) * m] = hXu[j + int64_t(c) * m];
        FRK_CUDA(cudaMemcpyAsync(dB.p, hB.data(), sizeof(double) * m * r, cudaMemcpyHostToDevice, st));
        FRK_CUDA(cudaStreamSynchronize(st));
    }
    {
        dim3 grid(unsigned((m + 127) / 128), unsigned(m));
        if (d == 1) knot_matrix_kernel<1><<<grid, 128, 0, st>>>(out.Xu.p, m, dH.p);
        else if (d == 2) knot_matrix_kernel<2><<<grid, 128, 0, st>>>(out.Xu.p, m, dH.p);
        else knot_matrix_kernel<3><<<grid, 128, 0, st>>>(out.Xu.p, m, dH.p);
        FRK_CUDA(cudaGetLastError());
    }
    const double one = 1.0, zero = 0.0;
    const int mi = int(m);
    // BBBH = BBB * H   ((d+1) x m)                                             :240
    out.BBBH.alloc(size_t(r) * m);
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, r, mi, mi, &one, dBBB.p, r, dH.p, mi, &zero, out.BBBH.p, r));
    // AH = H - (H B) BBB                                                       :197
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, mi, r, mi, &one, dH.p, mi, dB.p, mi, &zero, dHB.p, mi));
    {
        dim3 grid(unsigned((m + 127) / 128), unsigned(m));
        rank_update_kernel<<<grid, 128, 0, st>>>(dH.p, m, dHB.p, dBBB.p, r);
        FRK_CUDA(cudaGetLastError());
    }
    // AHA = AH - BBB' (B' AH)                                                  :198
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_T, CUBLAS_OP_N, r, mi, mi, &one, dB.p, mi, dH.p, mi, &zero, dBtAH.p, r));
    {
        // U = BBB' (m x r): reuse dHB as the transposed copy
        FRK_CUBLAS(cublasDgeam(lh.blas, CUBLAS_OP_T, CUBLAS_OP_N, mi, r, &one, dBBB.p, r, &zero, dHB.p, mi, dHB.p, mi));
        dim3 grid(unsigned((m + 127) / 128), unsigned(m));
        rank_update_kernel<<<grid, 128, 0, st>>>(dH.p, m, dHB.p, dBtAH.p, r);
        FRK_CUDA(cudaGetLastError());
    }
    // k largest eigenpairs                                                     :200-202
    DevBuf<double> dW(static_cast<size_t>(m));
    DevBuf<int> dinfo(1);
    int lwork = 0, meig = 0;
    const int il = mi - k + 1, iu = mi;
    FRK_CUSOLVER(cusolverDnDsyevdx_bufferSize(lh.solver, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I,
                                              CUBLAS_FILL_MODE_LOWER, mi, dH.p, mi, 0.0, 0.0, il, iu, &meig, dW.p, &lwork));
    DevBuf<double> dwork(static_cast<size_t>(lwork));
    FRK_CUSOLVER(cusolverDnDsyevdx(lh.solver, CUSOLVER_EIG_MODE_VECTOR, CUSOLVER_EIG_RANGE_I, CUBLAS_FILL_MODE_LOWER, mi,
                                   dH.p, mi, 0.0, 0.0, il, iu, &meig, dW.p, dwork.p, lwork, dinfo.p));
    int info = 0;
    FRK_CUDA(cudaMemcpyAsync(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    FRK_CUDA(cudaStreamSynchronize(st));
    if (info != 0 || meig != k) fail(4, "Spectra eigen decomposition did not converge");  // :48-50 (message kept)

    DevBuf<double> dgamma(size_t(m) * k), dBg(size_t(r) * k);
    gather_eigvecs_kernel<<<k, 256, 0, st>>>(dH.p, m, k, dgamma.p);
    FRK_CUDA(cudaGetLastError());
    // BBB * gamma  ((d+1) x k)
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, r, k, mi, &one, dBBB.p, r, dgamma.p, mi, &zero, dBg.p, r));
    const int K = k + d + 1;
    out.X.alloc(size_t(m) * K);
    out.UZ.alloc(size_t(m + d + 1) * K);
    {
        dim3 grid(unsigned((m + d + 1 + 127) / 128), unsigned(K));
        assemble_kernel<<<grid, 128, 0, st>>>(dgamma.p, dBg.p, dW.p, out.Xu.p, dxd.p, m, d, k, root, mean[0], mean[1],
                                              mean[2], nc[0], nc[1], nc[2], out.X.p, out.UZ.p);
        FRK_CUDA(cudaGetLastError());
    }
    for (int c = 0; c < d; ++c) {
        out.nconst[c] = nc[c] / root;  // :216
        out.shift[c] = mean[c];
    }
    FRK_CUDA(cudaStreamSynchronize(st));
}

void sym_eigen_dev(double* dA, int n, double* dvalues) {
    LinalgHandles& lh = linalg_handles();
    FRK_CUSOLVER(cusolverDnSetStream(lh.solver, cudaStreamPerThread));
    int lwork = 0;
    FRK_CUSOLVER(cusolverDnDsyevd_bufferSize(lh.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dvalues,
                                             &lwork));
    DevBuf<double> work(static_cast<size_t>(lwork));
    DevBuf<int> dinfo(1);
    FRK_CUSOLVER(cusolverDnDsyevd(lh.solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, dA, n, dvalues, work.p,
                                  lwork, dinfo.p));
    int info = 0;
    FRK_CUDA(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) fail(4, "symmetric eigen decomposition did not converge (info=%d)", info);
}

}  // namespace frk
