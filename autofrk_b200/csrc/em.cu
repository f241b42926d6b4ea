// EM0miss (R/autoFRK.R:560-737, DfromLK = NULL, diagonal Depsilon) on per-replicate NA-masked statistics.
//   frk_masked_stats : BiDBt = t(Bt) iDt Bt, ziDB = t(zt) iDt Bt, ziDz = t(zt) iDt zt for every replicate (:587-616),
//                      one pass over the rows of Fk with 0/1-over-D row weights (K4)
//   frk_EM0miss_stats: the EM iteration itself (:652-687) -- K x K work per (iteration, replicate): one
//                      eigendecomposition per iteration for ginv(mkpd(M)), one Cholesky + inverse per replicate.
#include <cmath>
#include <vector>
#include "capi_internal.cuh"

using frk_capi::guarded;
using frk_capi::require_device;

namespace {


__global__ void em_axpy_kernel(double* __restrict__ acc, const double* __restrict__ x, int64_t len, int first) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < len) acc[i] = first ? x[i] : acc[i] + x[i];
}
// iP = iM + G / s   (lower + upper, full)
__global__ void em_ip_kernel(const double* __restrict__ iM, const double* __restrict__ G, double inv_s, int K,
                             double* __restrict__ iP) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < K && j < K) iP[i + size_t(j) * K] = iM[i + size_t(j) * K] + G[i + size_t(j) * K] * inv_s;
}
__global__ void em_mirror_lower_kernel(double* __restrict__ A, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < K && j < K && i > j) A[j + size_t(i) * K] = A[i + size_t(j) * K];
}
__global__ void em_add_diag_kernel(double* __restrict__ A, int K, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) A[i + size_t(i) * K] += v;
}
// out[0] = sum_ij G[i,j] (eta_i eta_j + P[i,j])  = sum(diag(BiDBt %*% (eta eta' + Ptt)))   (:663, :668)
__global__ void em_s1_kernel(const double* __restrict__ G, const double* __restrict__ P, const double* __restrict__ eta,
                             int K, double* __restrict__ out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t e = threadIdx.x + int64_t(blockIdx.x) * blockDim.x; e < int64_t(K) * K; e += int64_t(gridDim.x) * blockDim.x) {
        const int i = int(e % K), j = int(e / K);
        s = fma(G[e], fma(eta[i], eta[j], P[e]), s);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}
// out[0] = sum |A - B|
__global__ void em_absdiff_kernel(const double* __restrict__ A, const double* __restrict__ B, int64_t len, double* __restrict__ out) {
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t e = threadIdx.x + int64_t(blockIdx.x) * blockDim.x; e < len; e += int64_t(gridDim.x) * blockDim.x) s += fabs(A[e] - B[e]);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}
// M = (E E' + S) / T, symmetrised   (:672, :681); EEt holds E E' (full)
__global__ void em_newm_kernel(const double* __restrict__ EEt, const double* __restrict__ S, double invT, int K, double* __restrict__ M) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i < K && j < K) {
        const double a = (EEt[i + size_t(j) * K] + S[i + size_t(j) * K]) * invT;
        const double b = (EEt[j + size_t(i) * K] + S[j + size_t(i) * K]) * invT;
        M[i + size_t(j) * K] = 0.5 * (a + b);
    }
}

double sum_blocks(const double* dpart, int nb) {
    std::vector<double> h(nb);
    FRK_CUDA(cudaMemcpy(h.data(), dpart, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    double s = 0.0;
    for (double v : h) s += v;
    return s;
}

// iM <- ginv(mkpd(M)):  mkpd (R/helper.R:278-288): min eigenvalue v <= 0 -> M + (max(0,-v) + 10^-7.5) I;
// MASS::ginv: drop singular values <= sqrt(eps) * largest.  M is symmetric, so one eigendecomposition does both.
void ginv_mkpd_dev(const double* dM, int K, double* diM, frk::DevBuf<double>& U, frk::DevBuf<double>& Us, frk::DevBuf<double>& lam) {
    frk::LinalgHandles& lh = frk::linalg_handles();
    FRK_CUDA(cudaMemcpy(U.p, dM, sizeof(double) * K * K, cudaMemcpyDeviceToDevice));
    frk::sym_eigen_dev(U.p, K, lam.p);
    std::vector<double> l(K), hU(size_t(K) * K);
    FRK_CUDA(cudaMemcpy(l.data(), lam.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
    FRK_CUDA(cudaMemcpy(hU.data(), U.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
    const double v = l[0];
    const double shift = (v <= 0.0) ? (std::fmax(0.0, -v) + std::pow(0.1, 7.5)) : 0.0;
    double dmax = 0.0;
    for (int j = 0; j < K; ++j) { l[j] += shift; dmax = std::fmax(dmax, std::fabs(l[j])); }
    const double tol = std::sqrt(2.220446049250313e-16) * dmax;
    for (int j = 0; j < K; ++j) {
        const double inv = (std::fabs(l[j]) > std::fmax(tol, 0.0)) ? 1.0 / l[j] : 0.0;
        for (int i = 0; i < K; ++i) hU[i + size_t(j) * K] *= inv;
    }
    FRK_CUDA(cudaMemcpy(Us.p, hU.data(), sizeof(double) * K * K, cudaMemcpyHostToDevice));
    const double one = 1.0, zero = 0.0;
    FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_T, K, K, K, &one, Us.p, K, U.p, K, &zero, diM, K));
}

}  // namespace

extern "C" {

int32_t frk_masked_stats(const double* Data0, const uint8_t* observed, const double* Fk, const double* Ddiag, int64_t n,
                         int32_t K, int32_t T, double* G_all, double* c_all, double* zz_all, int64_t* nobs_all) {
    return guarded([&] {
        require_device();
        if (n < 1 || K < 1 || T < 1) frk::fail(FRK_EINVAL, "masked_stats: Data (n x T) and Fk (n x K) must be non-empty");
        cudaStream_t st = nullptr;
        const size_t plen = frk::gram_packed_len(K, 1);
        frk::DevBuf<double> stats(plen * T), ptmp(plen);
        int64_t rc = (int64_t(1) << 30) / (int64_t(K + 3) * 8);
        rc = std::max<int64_t>(4096, rc / 16 * 16);
        rc = std::min(rc, (n + 1) / 2 * 2);
        frk::DevBuf<double> dF(size_t(rc) * K), dzt(static_cast<size_t>(rc)), dw(static_cast<size_t>(rc));
        std::vector<double> hw(static_cast<size_t>(rc));
        std::vector<int64_t> nobs(T, 0);
        frk::GramWorkspace& ws = frk_capi::gram_workspace();
        int first = 1;
        for (int64_t r0 = 0; r0 < n; r0 += rc) {
            const int64_t rows = std::min(rc, n - r0);
            FRK_CUDA(cudaMemcpy2DAsync(dF.p, sizeof(double) * rc, Fk + r0, sizeof(double) * n, sizeof(double) * rows, size_t(K),
                                       cudaMemcpyHostToDevice, st));
            for (int t = 0; t < T; ++t) {
                for (int64_t i = 0; i < rows; ++i) {
                    const bool o = observed[(r0 + i) + int64_t(t) * n] != 0;
                    hw[i] = o ? 1.0 / Ddiag[r0 + i] : 0.0;       // iDt = iD[O, O]   (:607)
                    nobs[t] += o;
                }
                FRK_CUDA(cudaMemcpyAsync(dw.p, hw.data(), sizeof(double) * rows, cudaMemcpyHostToDevice, st));
                FRK_CUDA(cudaMemcpyAsync(dzt.p, Data0 + r0 + int64_t(t) * n, sizeof(double) * rows, cudaMemcpyHostToDevice, st));
                frk::gram_stats_dev(dF.p, rows, K, rc, dzt.p, 1, rc, dw.p, ptmp.p, ws, st);
                em_axpy_kernel<<<unsigned((plen + 255) / 256), 256, 0, st>>>(stats.p + plen * t, ptmp.p, int64_t(plen), first);
                FRK_CUDA(cudaGetLastError());
                FRK_CUDA(cudaStreamSynchronize(st));
            }
            first = 0;
        }
        for (int t = 0; t < T; ++t) {
            const double* p = stats.p + plen * t;
            FRK_CUDA(cudaMemcpy(G_all + size_t(t) * K * K, p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
            FRK_CUDA(cudaMemcpy(c_all + size_t(t) * K, p + size_t(K) * K, sizeof(double) * K, cudaMemcpyDeviceToHost));
            FRK_CUDA(cudaMemcpy(zz_all + t, p + size_t(K) * K + K, sizeof(double), cudaMemcpyDeviceToHost));
            nobs_all[t] = nobs[t];
        }
    });
}

int32_t frk_EM0miss_stats(const double* G_all, const double* c_all, const double* zz_all, int64_t sumO, int32_t K, int32_t T,
                          const double* M0, double s0, int32_t maxit, double avgtol, int32_t has_vfixed, double vfixed,
                          double* M_out, double* s_out, double* etatt_out, int32_t* niter) {
    return guarded([&] {
        require_device();
        if (K < 1 || T < 1) frk::fail(FRK_EINVAL, "EM0miss: K and T must be positive");
        frk::LinalgHandles& lh = frk::linalg_handles();
        cudaStream_t st = nullptr;
        FRK_CUBLAS(cublasSetStream(lh.blas, st));
        FRK_CUSOLVER(cusolverDnSetStream(lh.solver, st));
        const size_t KK = size_t(K) * K;
        const double one = 1.0, zero = 0.0;
        frk::DevBuf<double> dG(KK * T), dc(size_t(K) * T), dM(KK), dMn(KK), diM(KK), dP(KK), dsum(KK), dE(size_t(K) * T), dEEt(KK),
            dU(KK), dUs(KK), dlam(static_cast<size_t>(K)), dpart(256);
        frk::DevBuf<int> dinfo(1);
        FRK_CUDA(cudaMemcpy(dG.p, G_all, sizeof(double) * KK * T, cudaMemcpyHostToDevice));
        FRK_CUDA(cudaMemcpy(dc.p, c_all, sizeof(double) * K * T, cudaMemcpyHostToDevice));
        FRK_CUDA(cudaMemcpy(dM.p, M0, sizeof(double) * KK, cudaMemcpyHostToDevice));
        int lwork = 0, lwork2 = 0;
        FRK_CUSOLVER(cusolverDnDpotrf_bufferSize(lh.solver, CUBLAS_FILL_MODE_LOWER, K, dP.p, K, &lwork));
        FRK_CUSOLVER(cusolverDnDpotri_bufferSize(lh.solver, CUBLAS_FILL_MODE_LOWER, K, dP.p, K, &lwork2));
        frk::DevBuf<double> work(static_cast<size_t>(std::max(lwork, lwork2)));
        const int lw = std::max(lwork, lwork2);
        double sumzz = 0.0;
        for (int t = 0; t < T; ++t) sumzz += zz_all[t];                       // sum(ziDz)
        dim3 g2(unsigned((K + 127) / 128), unsigned(K)), b2(128);
        const int nb = 64;

        // old$M <- mkpd(old$M)  (:647): ginv_mkpd_dev applies the same shift at the top of every iteration, so the
        // only place the shifted matrix itself is needed is `dif` of the first iteration.
        {
            FRK_CUDA(cudaMemcpy(dU.p, dM.p, sizeof(double) * KK, cudaMemcpyDeviceToDevice));
            frk::sym_eigen_dev(dU.p, K, dlam.p);
            double v = 0.0;
            FRK_CUDA(cudaMemcpy(&v, dlam.p, sizeof(double), cudaMemcpyDeviceToHost));
            if (v <= 0.0) {
                em_add_diag_kernel<<<(K + 127) / 128, 128, 0, st>>>(dM.p, K, std::fmax(0.0, -v) + std::pow(0.1, 7.5));
                FRK_CUDA(cudaGetLastError());
            }
        }
        double s_old = s0, dif = INFINITY;
        int cnt = 0;
        std::vector<double> heta(K), hc(size_t(K) * T);
        std::memcpy(hc.data(), c_all, sizeof(double) * K * T);
        while (dif > avgtol * (100.0 * K * K) && cnt < maxit) {                // :652
            ginv_mkpd_dev(dM.p, K, diM.p, dU, dUs, dlam);                       // ginv(mkpd(Ptt1))   (:659)
            FRK_CUDA(cudaMemsetAsync(dsum.p, 0, sizeof(double) * KK, st));
            double s1sum = 0.0, cross = 0.0;
            for (int t = 0; t < T; ++t) {                                       // :657-669
                const double* Gt = dG.p + KK * t;
                em_ip_kernel<<<g2, b2, 0, st>>>(diM.p, Gt, 1.0 / s_old, K, dP.p);   // iP = ginv(..) + BiDBt / old$s
                FRK_CUDA(cudaGetLastError());
                FRK_CUSOLVER(cusolverDnDpotrf(lh.solver, CUBLAS_FILL_MODE_LOWER, K, dP.p, K, work.p, lw, dinfo.p));
                int info = 0;
                FRK_CUDA(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
                if (info != 0) {
                    // mkpd(iP) (:659): not positive definite -> shift by the minimum eigenvalue, as the reference does
                    em_ip_kernel<<<g2, b2, 0, st>>>(diM.p, Gt, 1.0 / s_old, K, dP.p);
                    FRK_CUDA(cudaMemcpy(dU.p, dP.p, sizeof(double) * KK, cudaMemcpyDeviceToDevice));
                    frk::sym_eigen_dev(dU.p, K, dlam.p);
                    double v = 0.0;
                    FRK_CUDA(cudaMemcpy(&v, dlam.p, sizeof(double), cudaMemcpyDeviceToHost));
                    em_add_diag_kernel<<<(K + 127) / 128, 128, 0, st>>>(dP.p, K, std::fmax(0.0, -v) + std::pow(0.1, 7.5));
                    FRK_CUSOLVER(cusolverDnDpotrf(lh.solver, CUBLAS_FILL_MODE_LOWER, K, dP.p, K, work.p, lw, dinfo.p));
                    FRK_CUDA(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
                    if (info != 0) frk::fail(FRK_EINTERNAL, "EM0miss: iP is not positive definite after mkpd (potrf info=%d)", info);
                }
                FRK_CUSOLVER(cusolverDnDpotri(lh.solver, CUBLAS_FILL_MODE_LOWER, K, dP.p, K, work.p, lw, dinfo.p));   // Ptt <- solve(iP)
                em_mirror_lower_kernel<<<g2, b2, 0, st>>>(dP.p, K);
                FRK_CUDA(cudaGetLastError());
                // eta = Ptt %*% t(iDBt) %*% zt / old$s = Ptt c_t / old$s   (:661-662)
                const double inv_s = 1.0 / s_old;
                FRK_CUBLAS(cublasDgemv(lh.blas, CUBLAS_OP_N, K, K, &inv_s, dP.p, K, dc.p + size_t(K) * t, 1, &zero, dE.p + size_t(K) * t, 1));
                em_s1_kernel<<<nb, 256, 0, st>>>(Gt, dP.p, dE.p + size_t(K) * t, K, dpart.p);
                FRK_CUDA(cudaGetLastError());
                s1sum += sum_blocks(dpart.p, nb);                                // s1[tt]   (:668)
                em_axpy_kernel<<<unsigned((KK + 255) / 256), 256, 0, st>>>(dsum.p, dP.p, int64_t(KK), 0);   // sumPtt   (:666)
                FRK_CUDA(cudaMemcpy(heta.data(), dE.p + size_t(K) * t, sizeof(double) * K, cudaMemcpyDeviceToHost));
                for (int i = 0; i < K; ++i) cross += hc[i + size_t(K) * t] * heta[i];   // sum(ziDB * t(etatt))
            }
            // new$M = (etatt etatt' + sumPtt) / TT, symmetrised   (:672, :681)
            FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_T, K, K, T, &one, dE.p, K, dE.p, K, &zero, dEEt.p, K));
            em_newm_kernel<<<g2, b2, 0, st>>>(dEEt.p, dsum.p, 1.0 / T, K, dMn.p);
            FRK_CUDA(cudaGetLastError());
            const double s_new = has_vfixed ? vfixed : std::fmax((sumzz - 2.0 * cross + s1sum) / double(sumO), std::pow(0.1, 8));  // :673
            em_absdiff_kernel<<<nb, 256, 0, st>>>(dMn.p, dM.p, int64_t(KK), dpart.p);
            FRK_CUDA(cudaGetLastError());
            dif = sum_blocks(dpart.p, nb) + std::fabs(s_new - s_old);            // :682
            ++cnt;
            FRK_CUDA(cudaMemcpy(dM.p, dMn.p, sizeof(double) * KK, cudaMemcpyDeviceToDevice));
            s_old = s_new;
        }
        FRK_CUDA(cudaMemcpy(M_out, dM.p, sizeof(double) * KK, cudaMemcpyDeviceToHost));
        FRK_CUDA(cudaMemcpy(etatt_out, dE.p, sizeof(double) * K * T, cudaMemcpyDeviceToHost));
        *s_out = s_old;
        *niter = cnt;
    });
}

}  // extern "C"
