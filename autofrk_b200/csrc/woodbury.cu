// Woodbury solves and the masked-data likelihood as functions of (row-weighted) sufficient statistics:
//   invCz         R/autoFRK.R:845-852   (R + L L')^{-1} z,  R diagonal
//   getLikelihood R/helper.R:28-58      -2 log-likelihood with per-replicate NA masks
// Both reduce to K4 passes with row weights 1/R (and 0 for missing rows) plus K x K Cholesky work (cuSOLVER
// potrf/potrs); the only n-sized output is invCz's own result, produced chunk by chunk with a library GEMM.
#include <cmath>
#include <vector>
#include "capi_internal.cuh"
#include "host_staging.cuh"
#include "panel_gemm.cuh"

using frk_capi::guarded;
using frk_capi::require_device;

namespace {


__global__ void recip_kernel(const double* __restrict__ r, int64_t n, double* __restrict__ out) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = 1.0 / r[i];
}
__global__ void add_identity_kernel(double* __restrict__ A, int K) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) A[i + size_t(i) * K] += 1.0;
}
__global__ void axpy_packed_kernel(double* __restrict__ acc, const double* __restrict__ x, int64_t len, int first) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < len) acc[i] = first ? x[i] : acc[i] + x[i];
}
// out[i, t] = (z[i, t] - LI[i, t]) * ir[i]      (iRZ - iR %*% right, R/autoFRK.R:851)
__global__ void invcz_finish_kernel(const double* __restrict__ z, const double* __restrict__ LI, const double* __restrict__ ir,
                                    int64_t rows, int64_t ld, int T, double* __restrict__ out) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < rows)
        for (int t = blockIdx.y; t < T; t += gridDim.y) out[i + int64_t(t) * ld] = (z[i + int64_t(t) * ld] - LI[i + int64_t(t) * ld]) * ir[i];
}

// A (K x K, SPD, lower triangle read; destroyed) <- chol; B (K x nrhs) <- A^{-1} B; returns log det A
double chol_solve_logdet(double* dA, int K, double* dB, int nrhs) {
    frk::LinalgHandles& lh = frk::linalg_handles();
    FRK_CUSOLVER(cusolverDnSetStream(lh.solver, cudaStreamPerThread));
    int lwork = 0;
    FRK_CUSOLVER(cusolverDnDpotrf_bufferSize(lh.solver, CUBLAS_FILL_MODE_LOWER, K, dA, K, &lwork));
    frk::DevBuf<double> work(static_cast<size_t>(lwork));
    frk::DevBuf<int> dinfo(1);
    FRK_CUSOLVER(cusolverDnDpotrf(lh.solver, CUBLAS_FILL_MODE_LOWER, K, dA, K, work.p, lwork, dinfo.p));
    int info = 0;
    FRK_CUDA(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0) frk::fail(FRK_EINTERNAL, "I + L' R^-1 L is not positive definite (potrf info=%d)", info);
    if (dB && nrhs > 0) {
        FRK_CUSOLVER(cusolverDnDpotrs(lh.solver, CUBLAS_FILL_MODE_LOWER, K, nrhs, dA, K, dB, K, dinfo.p));
        FRK_CUDA(cudaMemcpy(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost));
        if (info != 0) frk::fail(FRK_EINTERNAL, "potrs failed (info=%d)", info);
    }
    std::vector<double> diag(K);
    FRK_CUDA(cudaMemcpy2D(diag.data(), sizeof(double), dA, sizeof(double) * (K + 1), sizeof(double), size_t(K),
                          cudaMemcpyDeviceToHost));
    double ld = 0.0;
    for (double v : diag) ld += 2.0 * std::log(v);
    return ld;
}

int64_t chunk_rows_for(int64_t n, int cols) {
    int64_t rc = (int64_t(1) << 30) / (int64_t(cols) * 8);
    rc = std::max<int64_t>(4096, rc / 16 * 16);
    return std::min(rc, (n + 1) / 2 * 2);
}

}  // namespace

extern "C" {

int32_t frk_invCz(const double* Rdiag, const double* L, const double* z, int64_t n, int32_t K, int32_t T, double* out) {
    return guarded([&] {
        require_device();
        if (n < 1 || K < 1 || T < 1) frk::fail(FRK_EINVAL, "invCz: L (n x K) and z (n x T) must be non-empty");
        frk::LinalgHandles& lh = frk::linalg_handles();
        cudaStream_t st = nullptr;
        FRK_CUBLAS(cublasSetStream(lh.blas, st));
        const int64_t rc = chunk_rows_for(n, K + 2 * T + 1);
        frk::DevBuf<double> dL(size_t(rc) * K), dz(size_t(rc) * T), dr(static_cast<size_t>(rc)), dir(static_cast<size_t>(rc)),
            dLI(size_t(rc) * T), dout(size_t(rc) * T);
        const size_t plen = frk::gram_packed_len(K, T);
        frk::DevBuf<double> packed(plen), ptmp(plen);
        frk::GramWorkspace& ws = frk_capi::gram_workspace();
        auto upload = [&](int64_t r0, int64_t rows) {
            FRK_CUDA(cudaMemcpy2DAsync(dL.p, sizeof(double) * rc, L + r0, sizeof(double) * n, sizeof(double) * rows, size_t(K),
                                       cudaMemcpyHostToDevice, st));
            FRK_CUDA(cudaMemcpy2DAsync(dz.p, sizeof(double) * rc, z + r0, sizeof(double) * n, sizeof(double) * rows, size_t(T),
                                       cudaMemcpyHostToDevice, st));
            FRK_CUDA(cudaMemcpyAsync(dr.p, Rdiag + r0, sizeof(double) * rows, cudaMemcpyHostToDevice, st));
            recip_kernel<<<unsigned((rows + 255) / 256), 256, 0, st>>>(dr.p, rows, dir.p);   // iR = solve(R)  :847
            FRK_CUDA(cudaGetLastError());
        };
        // pass 1: t(L) %*% iR %*% L and t(L) %*% iRZ   (:849) as row-weighted statistics
        int first = 1;
        for (int64_t r0 = 0; r0 < n; r0 += rc) {
            const int64_t rows = std::min(rc, n - r0);
            upload(r0, rows);
            frk::gram_stats_dev(dL.p, rows, K, rc, dz.p, T, rc, dir.p, ptmp.p, ws, st);
            axpy_packed_kernel<<<unsigned((plen + 255) / 256), 256, 0, st>>>(packed.p, ptmp.p, int64_t(plen), first);
            FRK_CUDA(cudaGetLastError());
            FRK_CUDA(cudaStreamSynchronize(st));
            first = 0;
        }
        // inner = solve(diag(1, K) + t(L) iR L) %*% (t(L) %*% iRZ)   (:849)
        add_identity_kernel<<<(K + 127) / 128, 128, 0, st>>>(packed.p, K);
        FRK_CUDA(cudaGetLastError());
        double* dinner = packed.p + size_t(K) * K;
        chol_solve_logdet(packed.p, K, dinner, T);
        // pass 2: iRZ - iR %*% (L %*% inner)   (:849-851)
        for (int64_t r0 = 0; r0 < n; r0 += rc) {
            const int64_t rows = std::min(rc, n - r0);
            if (n > rc) upload(r0, rows);  // single chunk: still resident from pass 1
            frk::panel_times_small_dev(dL.p, rows, rc, K, dinner, K, T, dLI.p, rc, st);       // L %*% inner
            dim3 grid(unsigned((rows + 255) / 256), unsigned(std::min(T, 65535)));
            invcz_finish_kernel<<<grid, 256, 0, st>>>(dz.p, dLI.p, dir.p, rows, rc, T, dout.p);
            FRK_CUDA(cudaGetLastError());
            FRK_CUDA(cudaMemcpy2DAsync(out + r0, sizeof(double) * n, dout.p, sizeof(double) * rc, sizeof(double) * rows,
                                       size_t(T), cudaMemcpyDeviceToHost, st));
            FRK_CUDA(cudaStreamSynchronize(st));
        }
    });
}

int32_t frk_getLikelihood(const double* Data0, const uint8_t* observed, const double* Fk, const double* M, double s,
                          const double* Ddiag, int64_t n, int32_t K, int32_t T, double* n2loglik) {
    return guarded([&] {
        require_device();
        if (n < 1 || K < 1 || T < 1) frk::fail(FRK_EINVAL, "getLikelihood: Data (n x T) and Fk (n x K) must be non-empty");
        frk::LinalgHandles& lh = frk::linalg_handles();
        cudaStream_t st = nullptr;
        FRK_CUBLAS(cublasSetStream(lh.blas, st));
        const double one = 1.0, zero = 0.0;
        // B = U diag(sqrt(pmax(value, 0))) U'  so that  L = Fk %*% B   (R/helper.R:40-42)
        frk::DevBuf<double> dU(size_t(K) * K), dlam(static_cast<size_t>(K)), dB(size_t(K) * K), dUs(size_t(K) * K);
        FRK_CUDA(cudaMemcpy(dU.p, M, sizeof(double) * K * K, cudaMemcpyHostToDevice));
        frk::sym_eigen_dev(dU.p, K, dlam.p);
        std::vector<double> lam(K), hU(size_t(K) * K), hUs(size_t(K) * K);
        FRK_CUDA(cudaMemcpy(lam.data(), dlam.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
        FRK_CUDA(cudaMemcpy(hU.data(), dU.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
        for (int j = 0; j < K; ++j) {
            const double sq = std::sqrt(std::fmax(lam[j], 0.0));
            for (int i = 0; i < K; ++i) hUs[i + size_t(j) * K] = hU[i + size_t(j) * K] * sq;
        }
        FRK_CUDA(cudaMemcpy(dUs.p, hUs.data(), sizeof(double) * K * K, cudaMemcpyHostToDevice));
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_T, K, K, K, &one, dUs.p, K, dU.p, K, &zero, dB.p, K));

        // per-replicate masked statistics of Fk:  G_t = Fk' W_t Fk, c_t = Fk' W_t z_t, zz_t = sum(W_t z_t^2),
        // W_t = diag(observed[, t] / (s * D)); one pass over the rows of Fk, all replicates per chunk.
        const size_t plen = frk::gram_packed_len(K, 1);
        frk::DevBuf<double> stats(plen * T), ptmp(plen);
        const int64_t rc = chunk_rows_for(n, K + 3);
        frk::DevBuf<double> dF(size_t(rc) * K), dzt(static_cast<size_t>(rc)), dw(static_cast<size_t>(rc));
        std::vector<double> hw(static_cast<size_t>(rc));
        std::vector<double> sumlogR(T, 0.0);
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
                    const double R = s * Ddiag[r0 + i];                      // R = s * Depsilon   (R/helper.R:39)
                    hw[i] = o ? 1.0 / R : 0.0;
                    if (o) { sumlogR[t] += std::log(R); ++nobs[t]; }
                }
                FRK_CUDA(cudaMemcpyAsync(dw.p, hw.data(), sizeof(double) * rows, cudaMemcpyHostToDevice, st));
                FRK_CUDA(cudaMemcpyAsync(dzt.p, Data0 + r0 + int64_t(t) * n, sizeof(double) * rows, cudaMemcpyHostToDevice, st));
                frk::gram_stats_dev(dF.p, rows, K, rc, dzt.p, 1, rc, dw.p, ptmp.p, ws, st);
                axpy_packed_kernel<<<unsigned((plen + 255) / 256), 256, 0, st>>>(stats.p + plen * t, ptmp.p, int64_t(plen), first);
                FRK_CUDA(cudaGetLastError());
                FRK_CUDA(cudaStreamSynchronize(st));  // hw / host staging is reused
            }
            first = 0;
        }
        // K x K stage per replicate
        double total = 0.0;
        frk::DevBuf<double> dA(size_t(K) * K), dT1(size_t(K) * K), drhs(static_cast<size_t>(K)), dsol(static_cast<size_t>(K));
        std::vector<double> hr(K), hs(K);
        for (int t = 0; t < T; ++t) {
            total += double(nobs[t]) * std::log(2.0 * M_PI);                  // sum(O) * log(2 * pi)   (:38)
            if (nobs[t] == 0) continue;
            const double* Gt = stats.p + plen * t;
            const double* ct = Gt + size_t(K) * K;
            double zz = 0.0;
            FRK_CUDA(cudaMemcpy(&zz, ct + K, sizeof(double), cudaMemcpyDeviceToHost));
            // t(L_t) R_t^-1 L_t = B' G_t B ;  t(L_t) R_t^-1 z_t = B' c_t
            FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, K, K, &one, Gt, K, dB.p, K, &zero, dT1.p, K));
            FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_T, CUBLAS_OP_N, K, K, K, &one, dB.p, K, dT1.p, K, &zero, dA.p, K));
            FRK_CUBLAS(cublasDgemv(lh.blas, CUBLAS_OP_T, K, K, &one, dB.p, K, ct, 1, &zero, drhs.p, 1));
            if (nobs[t] == 1) {
                // reference quirk (R/helper.R:48-51): a single observed row adds log(R_t + L_t L_t') and nothing else
                double tr = 0.0;
                std::vector<double> hA(size_t(K) * K);
                FRK_CUDA(cudaMemcpy(hA.data(), dA.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
                for (int i = 0; i < K; ++i) tr += hA[i + size_t(i) * K];      // L_t R^-1 L_t' (scalar) = trace
                total += sumlogR[t] + std::log1p(tr);                          // log(R (1 + L L'/R))
                continue;
            }
            add_identity_kernel<<<(K + 127) / 128, 128, 0, st>>>(dA.p, K);
            FRK_CUDA(cudaGetLastError());
            FRK_CUDA(cudaMemcpy(dsol.p, drhs.p, sizeof(double) * K, cudaMemcpyDeviceToDevice));
            const double logdet = chol_solve_logdet(dA.p, K, dsol.p, 1);      // logdet(I + L' R^-1 L)   (:29-34)
            FRK_CUDA(cudaMemcpy(hr.data(), drhs.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
            FRK_CUDA(cudaMemcpy(hs.data(), dsol.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
            double quad = zz;                                                  // z' R^-1 z - (L'R^-1 z)' (I + ..)^-1 (L'R^-1 z)
            for (int i = 0; i < K; ++i) quad -= hr[i] * hs[i];
            total += logdet + sumlogR[t] + quad;                               // :53-54
        }
        *n2loglik = total;
    });
}

int32_t frk_krige_from_stats(const double* Gg, const double* c, int32_t Tc, const double* M, int32_t K, int32_t want_V,
                             double* coef, double* V) {
    return guarded([&] {
        require_device();
        if (K < 1 || Tc < 1) frk::fail(FRK_EINVAL, "krige_from_stats: K and the number of data columns must be positive");
        frk::LinalgHandles& lh = frk::linalg_handles();
        cudaStream_t st = nullptr;
        FRK_CUBLAS(cublasSetStream(lh.blas, st));
        const double one = 1.0, zero = 0.0, mone = -1.0;
        const int NR = Tc + (want_V ? K : 0);  // right-hand sides: [c | Gg M]
        frk::DevBuf<double> dG(size_t(K) * K), dM(size_t(K) * K), dU(size_t(K) * K), dlam(static_cast<size_t>(K)), dA(size_t(K) * K),
            dGA(size_t(K) * K), dS(size_t(K) * K), dR(size_t(K) * NR), dY(size_t(K) * NR), dT(size_t(K) * NR);
        FRK_CUDA(cudaMemcpy(dG.p, Gg, sizeof(double) * K * K, cudaMemcpyHostToDevice));
        FRK_CUDA(cudaMemcpy(dM.p, M, sizeof(double) * K * K, cudaMemcpyHostToDevice));
        FRK_CUDA(cudaMemcpy(dR.p, c, sizeof(double) * K * Tc, cudaMemcpyHostToDevice));
        if (want_V)  // Gg M
            FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, K, K, &one, dG.p, K, dM.p, K, &zero, dR.p + size_t(K) * Tc, K));
        // dec <- eigen(M); A = vector %*% diag(sqrt(pmax(value, 0)))          R/autoFRK.R:1260-1264
        FRK_CUDA(cudaMemcpy(dU.p, dM.p, sizeof(double) * K * K, cudaMemcpyDeviceToDevice));
        frk::sym_eigen_dev(dU.p, K, dlam.p);
        std::vector<double> lam(K), hU(size_t(K) * K);
        FRK_CUDA(cudaMemcpy(lam.data(), dlam.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
        FRK_CUDA(cudaMemcpy(hU.data(), dU.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
        for (int j = 0; j < K; ++j) {
            const double sq = std::sqrt(std::fmax(lam[j], 0.0));
            for (int i = 0; i < K; ++i) hU[i + size_t(j) * K] *= sq;
        }
        FRK_CUDA(cudaMemcpy(dA.p, hU.data(), sizeof(double) * K * K, cudaMemcpyHostToDevice));
        // S = I + A' Gg A  (= diag(1, K) + t(L) iR L, R/autoFRK.R:849);   Y = S^-1 A' [c | Gg M]
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, K, K, &one, dG.p, K, dA.p, K, &zero, dGA.p, K));
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_T, CUBLAS_OP_N, K, K, K, &one, dA.p, K, dGA.p, K, &zero, dS.p, K));
        add_identity_kernel<<<(K + 127) / 128, 128, 0, st>>>(dS.p, K);
        FRK_CUDA(cudaGetLastError());
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_T, CUBLAS_OP_N, K, NR, K, &one, dA.p, K, dR.p, K, &zero, dY.p, K));
        chol_solve_logdet(dS.p, K, dY.p, NR);
        // T = [c | Gg M] - Gg A Y ;   [coef | M - V] = M T
        FRK_CUDA(cudaMemcpy(dT.p, dR.p, sizeof(double) * K * NR, cudaMemcpyDeviceToDevice));
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, NR, K, &mone, dGA.p, K, dY.p, K, &one, dT.p, K));
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, K, NR, K, &one, dM.p, K, dT.p, K, &zero, dY.p, K));
        FRK_CUDA(cudaMemcpy(coef, dY.p, sizeof(double) * K * Tc, cudaMemcpyDeviceToHost));
        if (want_V && V) {
            std::vector<double> hMV(size_t(K) * K), hM(size_t(K) * K);
            FRK_CUDA(cudaMemcpy(hMV.data(), dY.p + size_t(K) * Tc, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
            FRK_CUDA(cudaMemcpy(hM.data(), dM.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < size_t(K) * K; ++i) V[i] = hM[i] - hMV[i];   // V <- M - t(GM) %*% invCz(...)  :1270
        }
    });
}

int32_t frk_plan_predict_gram(frk_plan* plan, const double* xnew, int64_t n2, const double* Z, int32_t T,
                              const double* weights, double* G, double* C, double* zz) {
    return guarded([&] {
        if (!plan || !plan->mp) frk::fail(FRK_EINVAL, "plan must not be NULL");
        if (n2 < 1) frk::fail(FRK_EINVAL, "xnew must have at least one row");
        frk::MrtsPlan* mp = plan->mp;
        const int K = mp->k, d = mp->d;
        const int64_t ld = (n2 + 1) / 2 * 2;  // even leading dimension: every chunk stays 16-byte aligned
        frk::DevBuf<double> dx(size_t(ld) * d), dZ, dw, packed(frk::gram_packed_len(K, T));
        cudaStream_t st = nullptr;   // this thread's default stream; large pageable inputs go up through the pinned ring
        frk::upload_matrix(dx.p, ld, xnew, n2, n2, d, st);
        if (T > 0) dZ.alloc(size_t(ld) * T);
        if (weights) {
            dw.alloc(static_cast<size_t>(ld));
            frk::upload_matrix(dw.p, ld, weights, n2, n2, 1, st);
        }
        if (T > 0 && sizeof(double) * size_t(n2) * size_t(T) >= (size_t(256) << 20)) {
            // a large Z goes up chunk by chunk behind the basis evaluation instead of in front of it
            frk_capi::predict_gram_host_z(plan, dx.p, n2, ld, Z, n2, dZ.p, T, ld, weights ? dw.p : nullptr, packed.p, st);
        } else {
            if (T > 0) frk::upload_matrix(dZ.p, ld, Z, n2, n2, T, st);
            int32_t rc = frk_plan_predict_gram_dev(plan, dx.p, n2, ld, T > 0 ? dZ.p : nullptr, T, ld,
                                                   weights ? dw.p : nullptr, packed.p, 0, nullptr);
            if (rc != FRK_OK) frk::fail(rc, "%s", frk_last_error());
        }
        FRK_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
        FRK_CUDA(cudaMemcpy(G, packed.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
        if (T > 0 && C) FRK_CUDA(cudaMemcpy(C, packed.p + size_t(K) * K, sizeof(double) * K * T, cudaMemcpyDeviceToHost));
        if (zz) FRK_CUDA(cudaMemcpy(zz, packed.p + size_t(K) * K + size_t(K) * T, sizeof(double), cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
