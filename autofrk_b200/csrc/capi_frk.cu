// C ABI, part 2: fixed-rank-kriging reductions (K4, K5) and the fused predict -> reduce paths.
#include <functional>

#include "capi_internal.cuh"
#include "host_staging.cuh"
#include "panel_gemm.cuh"

using frk_capi::guarded;
using frk_capi::require_device;

namespace frk_capi {
frk_plan* make_plan(const double* Xu, int64_t n, int32_t d, const double* BBBH, const double* UZ, int64_t uz_rows,
                    int64_t uz_cols, const double* nconst, int32_t k, bool on_device);
}

namespace {


__global__ void axpy_kernel(double* __restrict__ acc, const double* __restrict__ x, int64_t len, int first) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i < len) acc[i] = first ? x[i] : acc[i] + x[i];
}

// se[i] = sqrt(max(0, sum_j sgn[j] * Y[i,j]^2)): rowSums((basis V) * basis) with V = R diag(sgn) R', Y = basis R
__global__ void signed_rownorm_kernel(const double* __restrict__ Y, int64_t rows, int64_t ld, int r,
                                      const double* __restrict__ sgn, double* __restrict__ se) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= rows) return;
    double s = 0.0;
    for (int j = 0; j < r; ++j) {
        const double y = Y[i + int64_t(j) * ld];
        s = fma(sgn[j] * y, y, s);
    }
    se[i] = sqrt(fmax(0.0, s));
}

// lower triangle <- (A + A')/2
__global__ void symmetrize_lower_kernel(double* __restrict__ A, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        for (int j = blockIdx.y; j < i; j += gridDim.y) A[i + int64_t(j) * n] = 0.5 * (A[i + int64_t(j) * n] + A[j + int64_t(i) * n]);
}

__global__ void scale_cols_inplace_kernel(double* __restrict__ A, int n, int cols, const double* __restrict__ f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i < n && j < cols) A[i + int64_t(j) * n] *= f[j];
}

// B = [w | R] on the device with V = R diag(sgn) R' (eigen-decomposition of V; |lambda| <= K eps max|lambda| dropped --
// the rank tolerance of LAPACK-style rank tests and the size of the rounding error of the reference's own FP64
// evaluation of the bilinear form), so that
//   yhat = basis %*% w                                                              R/autoFRK.R:1216
//   se^2 = rowSums((basis %*% V) * basis) = sum_j sgn_j (basis %*% R)[, j]^2        R/autoFRK.R:1220
// both come out of ONE product basis %*% B with q = T + rank(V) columns.  V from cMLEimat has rank <= min(K, T).
// Returns rank(V) (0 when V == nullptr).
int coefficient_matrix_dev(const double* w, int T, const double* V, int K, cudaStream_t st, frk::DevBuf<double>& dB,
                           std::vector<double>& sgn) {
    int r = 0;
    sgn.clear();
    if (V) {
        frk::DevBuf<double> dV(size_t(K) * K), lam(static_cast<size_t>(K));
        FRK_CUDA(cudaMemcpy(dV.p, V, sizeof(double) * K * K, cudaMemcpyHostToDevice));
        // b' V b = b' ((V + V')/2) b: a V that is symmetric only up to rounding (the kriging V = M - ... is) must enter
        // through its symmetric part, or the eigen-solver's lower triangle would stand for a different quadratic form
        symmetrize_lower_kernel<<<dim3(unsigned((K + 127) / 128), unsigned(std::min(K, 65535))), 128, 0, st>>>(dV.p, K);
        FRK_CUDA(cudaGetLastError());
        FRK_CUDA(cudaStreamSynchronize(st));
        frk::sym_eigen_dev(dV.p, K, lam.p);                     // ascending; lower triangle of V is read
        std::vector<double> lam_h(K);
        FRK_CUDA(cudaMemcpy(lam_h.data(), lam.p, sizeof(double) * K, cudaMemcpyDeviceToHost));
        double amax = 0.0;
        for (double l : lam_h) amax = std::fmax(amax, std::fabs(l));
        std::vector<int> keep;
        for (int i = K - 1; i >= 0; --i)
            if (std::fabs(lam_h[i]) > double(K) * 2.220446049250313e-16 * amax) keep.push_back(i);
        r = int(keep.size());
        dB.alloc(size_t(K) * (T + std::max(r, 1)));
        std::vector<double> scale(std::max(r, 1));
        for (int j = 0; j < r; ++j) {
            scale[j] = std::sqrt(std::fabs(lam_h[keep[j]]));
            sgn.push_back(lam_h[keep[j]] < 0.0 ? -1.0 : 1.0);
            FRK_CUDA(cudaMemcpyAsync(dB.p + size_t(K) * (T + j), dV.p + size_t(K) * keep[j], sizeof(double) * K,
                                     cudaMemcpyDeviceToDevice, st));
        }
        if (r > 0) {
            frk::DevBuf<double> dscale(static_cast<size_t>(r));
            FRK_CUDA(cudaMemcpyAsync(dscale.p, scale.data(), sizeof(double) * r, cudaMemcpyHostToDevice, st));
            dim3 grid(unsigned((K + 127) / 128), unsigned(r));
            scale_cols_inplace_kernel<<<grid, 128, 0, st>>>(dB.p + size_t(K) * T, K, r, dscale.p);
            FRK_CUDA(cudaGetLastError());
            FRK_CUDA(cudaStreamSynchronize(st));
        }
    } else {
        dB.alloc(size_t(K) * T);
    }
    FRK_CUDA(cudaMemcpyAsync(dB.p, w, sizeof(double) * K * T, cudaMemcpyHostToDevice, st));
    return r;
}

// One-shot all-reduce over NVLink peer memory: every rank reads every rank's packed statistics straight from the
// peers' HBM (P2P loads through NVSwitch) and adds them in rank order, so all ranks obtain bitwise-identical sums.
struct PeerPtrs {
    const double* p[16];
};
__global__ void peer_sum_kernel(const PeerPtrs ptrs, int nranks, int64_t len, double* __restrict__ out) {
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < len; i += int64_t(gridDim.x) * blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < nranks; ++r) s += __ldcv(ptrs.p[r] + i);   // .cv: never served from a stale cache line
        out[i] = s;
    }
}

int64_t default_chunk_rows(const frk::MrtsPlan* mp, int K) {
    const int64_t bm = mp->ki ? mp->ki->bm : 128;
    const int64_t wave = bm * mp->sm_count;
    int64_t rc = (int64_t(2) << 30) / (int64_t(K) * 8);  // ~2 GB of basis per chunk
    return std::max<int64_t>(wave, rc / wave * wave);
}

// `before_gram(r0, rows)` (optional) runs on the host between the launch of K1 and of K4 for each row chunk: the
// host-buffer entry point uses it to bring that chunk's rows of Z up while K1 of the same chunk is running.
using BeforeGram = std::function<void(int64_t, int64_t)>;
void predict_gram(frk_plan* plan, const double* dx, int64_t n2, int64_t ldx, const double* dZ, int T, int64_t ldz,
                  const double* dw, double* dpacked, int64_t chunk_rows, cudaStream_t st,
                  const BeforeGram* before_gram = nullptr) {
    frk::MrtsPlan* mp = plan->mp;
    const int K = mp->k;
    if (chunk_rows <= 0) chunk_rows = default_chunk_rows(mp, K);
    chunk_rows = std::min(chunk_rows, n2);
    if (chunk_rows % 2) ++chunk_rows;  // keeps every chunk 16-byte aligned for the Gram kernel's cp.async path
    const frk::StreamGuardScope ordered(plan->guard, st);
    plan->guard.grow(plan->fchunk, size_t(chunk_rows) * K);
    const size_t plen = frk::gram_packed_len(K, T);
    plan->guard.grow(plan->packed_tmp, plen);
    frk::GramWorkspace& ws = frk_capi::gram_workspace();
    int first = 1;
    for (int64_t r0 = 0; r0 < n2; r0 += chunk_rows) {
        const int64_t rows = std::min(chunk_rows, n2 - r0);
        frk::plan_predict_basis_dev(mp, dx + r0, rows, ldx, plan->fchunk.p, chunk_rows, st);
        if (before_gram) (*before_gram)(r0, rows);
        frk::gram_stats_dev(plan->fchunk.p, rows, K, chunk_rows, T > 0 ? dZ + r0 : nullptr, T, ldz, dw ? dw + r0 : nullptr,
                            plan->packed_tmp.p, ws, st);
        axpy_kernel<<<unsigned((plen + 255) / 256), 256, 0, st>>>(dpacked, plan->packed_tmp.p, int64_t(plen), first);
        FRK_CUDA(cudaGetLastError());
        mp->launches += 4;
        first = 0;
    }
}

}  // namespace

namespace frk_capi {

// predict -> Gram with Z still in HOST memory: the rows of Z that a chunk needs go up on a copy stream (through the
// pinned ring when Z is pageable) while K1 evaluates that chunk's basis, and K4 of chunk i runs while the host stages
// chunk i + 1 -- the upload of an n2 x T matrix (3.3 GB at 4 M rows, T = 100) leaves the critical path.
void predict_gram_host_z(frk_plan* plan, const double* dx, int64_t n2, int64_t ldx, const double* hZ, int64_t ldzh,
                         double* dZ, int T, int64_t ldz, const double* dw, double* dpacked, cudaStream_t st) {
    cudaStream_t cs = nullptr;
    cudaEvent_t ev = nullptr;
    FRK_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    FRK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    auto cleanup = [&] {
        cudaStreamSynchronize(cs);
        cudaEventDestroy(ev);
        cudaStreamDestroy(cs);
    };
    const BeforeGram up = [&](int64_t r0, int64_t rows) {
        frk::upload_matrix(dZ + r0, ldz, hZ + r0, ldzh, rows, T, cs);
        FRK_CUDA(cudaEventRecord(ev, cs));
        FRK_CUDA(cudaStreamWaitEvent(st, ev, 0));
    };
    try {
        predict_gram(plan, dx, n2, ldx, dZ, T, ldz, dw, dpacked, 0, st, &up);
        FRK_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace frk_capi

extern "C" {

int64_t frk_gram_packed_len(int32_t K, int32_t T) { return int64_t(frk::gram_packed_len(K, T)); }

int64_t frk_gram_job_plan(int32_t K, int32_t T, int32_t* njobs, int32_t* triples, int64_t cap) {
    if (K < 1 || T < 0) return -1;
    const std::vector<frk::GramJobHost> jobs = frk::gram_pack_jobs(K, T);
    if (njobs) *njobs = int32_t(jobs.size());
    int64_t count = 0;
    for (size_t j = 0; j < jobs.size(); ++j)
        for (const auto& ab : jobs[j].tiles) {
            if (triples && count < cap) {
                triples[3 * count + 0] = int32_t(j);
                triples[3 * count + 1] = ab.first;
                triples[3 * count + 2] = ab.second;
            }
            ++count;
        }
    return count;
}

int64_t frk_gram_mid_plan(int32_t K, int32_t T, int32_t* nsets, int32_t* sp_load, int32_t* tiles, int64_t cap) {
    frk::GramMidPlan P;
    if (!frk::gram_mid_plan(K, T, P)) return -1;
    if (nsets) *nsets = P.nsets;
    if (sp_load)
        for (int p = 0; p < 4; ++p) sp_load[p] = P.sp_load[p];
    for (size_t i = 0; tiles && i < P.tiles.size() && int64_t(i) < cap; ++i) {
        const frk::GramMidTile& t = P.tiles[i];
        tiles[5 * i + 0] = t.ga;
        tiles[5 * i + 1] = t.gb;
        tiles[5 * i + 2] = t.isz;
        tiles[5 * i + 3] = t.warp;
        tiles[5 * i + 4] = t.slot;
    }
    return int64_t(P.tiles.size());
}

int32_t frk_gram_stats_dev(const double* d_F, int64_t n, int32_t K, int64_t ldf, const double* d_Z, int32_t T,
                           int64_t ldz, const double* d_weights, double* d_packed, void* stream) {
    return guarded([&] {
        frk::gram_stats_dev(d_F, n, K, ldf, d_Z, T, ldz, d_weights, d_packed, frk_capi::gram_workspace(),
                            static_cast<cudaStream_t>(stream));
    });
}

int32_t frk_gram_stats(const double* F, int64_t n, int32_t K, const double* Z, int32_t T, const double* weights,
                       double* G, double* C, double* zz) {
    return guarded([&] {
        require_device();
        if (K < 1 || n < 1) frk::fail(FRK_EINVAL, "F must be a non-empty matrix");
        if (T < 0) frk::fail(FRK_EINVAL, "T must be non-negative");
        // stream the rows through the device in chunks (two buffers: H2D of chunk c+1 overlaps the Gram of chunk c)
        int64_t rc = (int64_t(1) << 30) / (int64_t(K + T) * 8);
        rc = std::max<int64_t>(4096, rc / 16 * 16);
        rc = std::min(rc, (n + 1) / 2 * 2);
        frk::DevBuf<double> dF[2], dZ[2], dW[2], packed(frk::gram_packed_len(K, T)), ptmp(frk::gram_packed_len(K, T));
        for (int b = 0; b < 2; ++b) {
            dF[b].alloc(size_t(rc) * K);
            if (T > 0) dZ[b].alloc(size_t(rc) * T);
            if (weights) dW[b].alloc(static_cast<size_t>(rc));
        }
        cudaStream_t s_comp, s_copy;
        FRK_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
        FRK_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
        cudaEvent_t ev_in[2], ev_done[2];
        for (int b = 0; b < 2; ++b) {
            FRK_CUDA(cudaEventCreateWithFlags(&ev_in[b], cudaEventDisableTiming));
            FRK_CUDA(cudaEventCreateWithFlags(&ev_done[b], cudaEventDisableTiming));
        }
        frk::GramWorkspace& ws = frk_capi::gram_workspace();
        const size_t plen = frk::gram_packed_len(K, T);
        try {
            int c = 0, first = 1;
            for (int64_t r0 = 0; r0 < n; r0 += rc, ++c) {
                const int b = c & 1;
                const int64_t rows = std::min(rc, n - r0);
                if (c >= 2) FRK_CUDA(cudaStreamWaitEvent(s_copy, ev_done[b], 0));
                frk::upload_matrix(dF[b].p, rc, F + r0, n, rows, K, s_copy);     // pageable R matrices: through the pinned ring
                if (T > 0) frk::upload_matrix(dZ[b].p, rc, Z + r0, n, rows, T, s_copy);
                if (weights)
                    FRK_CUDA(cudaMemcpyAsync(dW[b].p, weights + r0, sizeof(double) * rows, cudaMemcpyHostToDevice, s_copy));
                FRK_CUDA(cudaEventRecord(ev_in[b], s_copy));
                FRK_CUDA(cudaStreamWaitEvent(s_comp, ev_in[b], 0));
                frk::gram_stats_dev(dF[b].p, rows, K, rc, T > 0 ? dZ[b].p : nullptr, T, rc, weights ? dW[b].p : nullptr,
                                    ptmp.p, ws, s_comp);
                axpy_kernel<<<unsigned((plen + 255) / 256), 256, 0, s_comp>>>(packed.p, ptmp.p, int64_t(plen), first);
                FRK_CUDA(cudaGetLastError());
                FRK_CUDA(cudaEventRecord(ev_done[b], s_comp));
                first = 0;
            }
            FRK_CUDA(cudaStreamSynchronize(s_comp));
            FRK_CUDA(cudaMemcpy(G, packed.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
            if (T > 0 && C) FRK_CUDA(cudaMemcpy(C, packed.p + size_t(K) * K, sizeof(double) * K * T, cudaMemcpyDeviceToHost));
            if (zz) FRK_CUDA(cudaMemcpy(zz, packed.p + size_t(K) * K + size_t(K) * T, sizeof(double), cudaMemcpyDeviceToHost));
        } catch (...) {
            cudaStreamDestroy(s_comp);
            cudaStreamDestroy(s_copy);
            for (int b = 0; b < 2; ++b) { cudaEventDestroy(ev_in[b]); cudaEventDestroy(ev_done[b]); }
            throw;
        }
        cudaStreamDestroy(s_comp);
        cudaStreamDestroy(s_copy);
        for (int b = 0; b < 2; ++b) { cudaEventDestroy(ev_in[b]); cudaEventDestroy(ev_done[b]); }
    });
}

int32_t frk_plan_predict_gram_dev(frk_plan* plan, const double* d_xnew, int64_t n2, int64_t ldx, const double* d_Z,
                                  int32_t T, int64_t ldz, const double* d_weights, double* d_packed,
                                  int64_t chunk_rows, void* stream) {
    return guarded([&] {
        if (!plan || !plan->mp) frk::fail(FRK_EINVAL, "plan must not be NULL");
        if (n2 < 1) frk::fail(FRK_EINVAL, "xnew must have at least one row");
        if (ldx < n2 || (T > 0 && ldz < n2)) frk::fail(FRK_EINVAL, "leading dimensions must be >= n2");
        predict_gram(plan, d_xnew, n2, ldx, d_Z, T, ldz, d_weights, d_packed, chunk_rows,
                     static_cast<cudaStream_t>(stream));
    });
}

int32_t frk_peer_sum_dev(const void* const* d_ptrs, int32_t nranks, int64_t len, double* d_out, void* stream) {
    return guarded([&] {
        if (nranks < 1 || nranks > 16) frk::fail(FRK_EINVAL, "peer sum supports 1..16 ranks (one NVSwitch node)");
        if (len < 1) return;
        PeerPtrs pp;
        for (int r = 0; r < 16; ++r) pp.p[r] = (r < nranks) ? static_cast<const double*>(d_ptrs[r]) : nullptr;
        const int grid = int(std::min<int64_t>((len + 255) / 256, 592));
        peer_sum_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pp, nranks, len, d_out);
        FRK_CUDA(cudaGetLastError());
    });
}

int32_t frk_getHalf(const double* G, int64_t ldg, int32_t K, double* half) {
    return guarded([&] {
        require_device();
        if (K < 1 || ldg < K) frk::fail(FRK_EINVAL, "G must be K x K with ldg >= K");
        frk::DevBuf<double> dG(size_t(K) * K), dh(size_t(K) * K);
        FRK_CUDA(cudaMemcpy2D(dG.p, sizeof(double) * K, G, sizeof(double) * ldg, sizeof(double) * K, size_t(K),
                              cudaMemcpyHostToDevice));
        frk::half_from_gram_dev(dG.p, K, K, dh.p);
        FRK_CUDA(cudaMemcpy(half, dh.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
    });
}

int32_t frk_cMLEimat_stats(const double* G, int64_t ldg, const double* C, int64_t ldc, double zz, int32_t K, int32_t T,
                           int64_t n, double s, int32_t wsave, int32_t only_loglik, int32_t on_device, double* v,
                           double* negloglik, double* M, double* w, double* V) {
    return guarded([&] {
        require_device();
        if (K < 1 || T < 1) frk::fail(FRK_EINVAL, "K and T must be positive");
        if (ldg < K || ldc < K) frk::fail(FRK_EINVAL, "leading dimensions must be >= K");
        frk::DevBuf<double> dG, dC;
        const double *pG = G, *pC = C;
        int64_t lg = ldg, lc = ldc;
        if (!on_device) {
            dG.alloc(size_t(K) * K);
            dC.alloc(size_t(K) * T);
            FRK_CUDA(cudaMemcpy2D(dG.p, sizeof(double) * K, G, sizeof(double) * ldg, sizeof(double) * K, size_t(K),
                                  cudaMemcpyHostToDevice));
            FRK_CUDA(cudaMemcpy2D(dC.p, sizeof(double) * K, C, sizeof(double) * ldc, sizeof(double) * K, size_t(T),
                                  cudaMemcpyHostToDevice));
            pG = dG.p; pC = dC.p; lg = K; lc = K;
        }
        frk::CmleOut out;
        frk::cmle_imat_from_stats_dev(pG, lg, pC, lc, zz, K, T, double(n), s, wsave != 0, only_loglik != 0, out);
        if (v) *v = out.v;
        if (negloglik) *negloglik = out.negloglik;
        if (!only_loglik && M) FRK_CUDA(cudaMemcpy(M, out.M.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
        if (!only_loglik && wsave) {
            if (w) FRK_CUDA(cudaMemcpy(w, out.w.p, sizeof(double) * K * T, cudaMemcpyDeviceToHost));
            if (V) FRK_CUDA(cudaMemcpy(V, out.V.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
        }
    });
}

int32_t frk_cMLEimat_sweep(const double* G, int64_t ldg, const double* C, int64_t ldc, double zz, int32_t T, int64_t n,
                           double s, const int32_t* Ks, int32_t nK, double* negloglik) {
    return guarded([&] {
        require_device();
        if (nK < 1) return;
        if (!G || !C || !Ks || !negloglik) frk::fail(FRK_EINVAL, "G, C, Ks and negloglik are required");
        if (T < 1) frk::fail(FRK_EINVAL, "T must be positive");
        int Kmax = 0;
        for (int i = 0; i < nK; ++i) {
            if (Ks[i] < 1) frk::fail(FRK_EINVAL, "every K must be positive");
            Kmax = std::max(Kmax, int(Ks[i]));
        }
        if (ldg < Kmax || ldc < Kmax) frk::fail(FRK_EINVAL, "leading dimensions must be >= max(Ks)");
        frk::DevBuf<double> dG(size_t(Kmax) * Kmax), dC(size_t(Kmax) * T);
        FRK_CUDA(cudaMemcpy2D(dG.p, sizeof(double) * Kmax, G, sizeof(double) * ldg, sizeof(double) * Kmax, size_t(Kmax),
                              cudaMemcpyHostToDevice));
        FRK_CUDA(cudaMemcpy2D(dC.p, sizeof(double) * Kmax, C, sizeof(double) * ldc, sizeof(double) * Kmax, size_t(T),
                              cudaMemcpyHostToDevice));
        std::vector<int> ks(Ks, Ks + nK);
        frk::negloglik_sweep_dev(dG.p, Kmax, dC.p, Kmax, zz, T, double(n), s, ks.data(), nK, negloglik);
    });
}

int32_t frk_cMLE(const double* Fk, int64_t n, int32_t K, int32_t TT, double trS, const double* half, const double* JSJ,
                 double s, double ldet, int32_t wsave, int32_t only_loglik, int32_t has_vfixed, double vfixed, double* v,
                 double* negloglik, double* M, double* L, int32_t* nL) {
    return guarded([&] {
        require_device();
        if (K < 1) frk::fail(FRK_EINVAL, "K must be positive");
        frk::DevBuf<double> dh(size_t(K) * K), dJ(size_t(K) * K);
        FRK_CUDA(cudaMemcpy(dh.p, half, sizeof(double) * K * K, cudaMemcpyHostToDevice));
        FRK_CUDA(cudaMemcpy(dJ.p, JSJ, sizeof(double) * K * K, cudaMemcpyHostToDevice));
        frk::CmleOut out;
        frk::cmle_core_dev(dh.p, dJ.p, nullptr, K, K, TT, double(n), trS, s, ldet, wsave != 0, only_loglik != 0,
                           has_vfixed != 0, vfixed, out);
        if (v) *v = out.v;
        if (negloglik) *negloglik = out.negloglik;
        if (only_loglik) return;
        if (M) FRK_CUDA(cudaMemcpy(M, out.M.p, sizeof(double) * K * K, cudaMemcpyDeviceToHost));
        if (nL) *nL = wsave ? out.nL : 0;
        if (wsave && Fk && L) {
            // L = Fk %*% LA[, 1:nL]   (R/autoFRK.R:387-391): the panel kernel, streamed over row chunks
            int64_t rc = std::min<int64_t>(n, std::max<int64_t>(1024, (int64_t(512) << 20) / (int64_t(K) * 8)));
            rc += rc & 1;
            frk::DevBuf<double> dF(size_t(rc) * K), dL(size_t(rc) * out.nL);
            for (int64_t r0 = 0; r0 < n; r0 += rc) {
                const int64_t rows = std::min(rc, n - r0);
                FRK_CUDA(cudaMemcpy2D(dF.p, sizeof(double) * rc, Fk + r0, sizeof(double) * n, sizeof(double) * rows,
                                      size_t(K), cudaMemcpyHostToDevice));
                frk::panel_times_small_dev(dF.p, rows, rc, K, out.LA.p, K, out.nL, dL.p, rc, nullptr);
                FRK_CUDA(cudaMemcpy2D(L + r0, sizeof(double) * n, dL.p, sizeof(double) * rc, sizeof(double) * rows,
                                      size_t(out.nL), cudaMemcpyDeviceToHost));
            }
        }
    });
}

int32_t frk_plan_predict_frk(frk_plan* plan, const double* xnew, int64_t n2, const double* w, int32_t T, const double* V,
                             double* yhat, double* se) {
    return guarded([&] {
        if (!plan || !plan->mp) frk::fail(FRK_EINVAL, "plan must not be NULL");
        if (n2 <= 0) return;
        frk::MrtsPlan* mp = plan->mp;
        const int K = mp->k, d = mp->d;
        if (T < 1 || !w || !yhat) frk::fail(FRK_EINVAL, "w (K x T) and yhat (n2 x T) are required");
        cudaStream_t st = nullptr;
        const bool want_se = (V != nullptr && se != nullptr);
        // B = [w | R] (coefficient_matrix_dev): basis %*% B comes straight out of K1 through a derived plan (W' = W B_rest):
        // q = T + rank(V) columns of tensor work instead of K columns plus an n2 x K x K product, and no n2 x K basis anywhere.
        std::vector<double> sgn;
        frk::DevBuf<double> dB;
        const int r = coefficient_matrix_dev(w, T, want_se ? V : nullptr, K, st, dB, sgn);
        const int q = T + r;
        struct PlanGuard {
            frk::MrtsPlan* p;
            ~PlanGuard() { frk::plan_destroy(p); }
        } derived{frk::plan_derive_dev(mp, dB.p, q, st)};

        // Row chunks, software-pipelined on two streams: while chunk c runs through K1 (+ the se reduction), the
        // coordinates of chunk c+1 go up and the results of chunk c-1 come down.  Copies from / to pageable host
        // memory block the host thread, so the next chunk's kernels are always queued before a copy is issued.
        plan->ensure_streams();
        cudaStream_t sc = plan->s_comp, sx = plan->s_copy;
        const int64_t bm = derived.p->ki ? derived.p->ki->bm : 128;
        const int64_t wave = bm * derived.p->sm_count;
        int64_t rc = std::max<int64_t>(wave, ((int64_t(256) << 20) / (int64_t(q + d) * 8)) / wave * wave);
        rc = std::min(rc, n2);
        if (rc % 2) ++rc;
        const int nchunk = int((n2 + rc - 1) / rc);
        frk::DevBuf<double> dx[2], dy[2], dse[2], dsgn;
        for (int b2 = 0; b2 < 2; ++b2) {
            dx[b2].alloc(size_t(rc) * d);
            dy[b2].alloc(size_t(rc) * q);
            if (want_se) dse[b2].alloc(static_cast<size_t>(rc));
        }
        if (want_se) {
            dsgn.alloc(static_cast<size_t>(std::max(r, 1)));
            if (r > 0) FRK_CUDA(cudaMemcpyAsync(dsgn.p, sgn.data(), sizeof(double) * r, cudaMemcpyHostToDevice, sc));
        }
        cudaEvent_t ev_up[2], ev_done[2], ev_down[2];
        for (int b2 = 0; b2 < 2; ++b2) {
            FRK_CUDA(cudaEventCreateWithFlags(&ev_up[b2], cudaEventDisableTiming));
            FRK_CUDA(cudaEventCreateWithFlags(&ev_done[b2], cudaEventDisableTiming));
            FRK_CUDA(cudaEventCreateWithFlags(&ev_down[b2], cudaEventDisableTiming));
        }
        struct EvGuard {
            cudaEvent_t* e[3];
            ~EvGuard() {
                for (auto* arr : e)
                    for (int i = 0; i < 2; ++i) cudaEventDestroy(arr[i]);
            }
        } evguard{{ev_up, ev_done, ev_down}};
        auto rows_of = [&](int c) { return std::min(rc, n2 - int64_t(c) * rc); };
        auto upload = [&](int c) {   // coordinates of chunk c -> dx[c & 1] (d column pieces of the n2 x d matrix)
            const int b2 = c & 1;
            if (c >= 2) FRK_CUDA(cudaStreamWaitEvent(sx, ev_done[b2], 0));   // chunk c-2 has consumed this buffer
            FRK_CUDA(cudaMemcpy2DAsync(dx[b2].p, sizeof(double) * rc, xnew + int64_t(c) * rc, sizeof(double) * n2,
                                       sizeof(double) * rows_of(c), size_t(d), cudaMemcpyHostToDevice, sx));
            FRK_CUDA(cudaEventRecord(ev_up[b2], sx));
        };
        auto compute = [&](int c) {
            const int b2 = c & 1;
            const int64_t rows = rows_of(c);
            FRK_CUDA(cudaStreamWaitEvent(sc, ev_up[b2], 0));
            if (c >= 2) FRK_CUDA(cudaStreamWaitEvent(sc, ev_down[b2], 0));   // results of chunk c-2 have left dy / dse
            frk::plan_predict_x1_dev(derived.p, dx[b2].p, rows, rc, dy[b2].p, rc, sc);
            if (want_se) {
                signed_rownorm_kernel<<<unsigned((rows + 255) / 256), 256, 0, sc>>>(dy[b2].p + size_t(rc) * T, rows, rc, r, dsgn.p,
                                                                                   dse[b2].p);
                FRK_CUDA(cudaGetLastError());
            }
            FRK_CUDA(cudaEventRecord(ev_done[b2], sc));
        };
        auto download = [&](int c) {
            const int b2 = c & 1;
            const int64_t rows = rows_of(c), r0 = int64_t(c) * rc;
            FRK_CUDA(cudaStreamWaitEvent(sx, ev_done[b2], 0));
            FRK_CUDA(cudaMemcpy2DAsync(yhat + r0, sizeof(double) * n2, dy[b2].p, sizeof(double) * rc, sizeof(double) * rows,
                                       size_t(T), cudaMemcpyDeviceToHost, sx));
            if (want_se) FRK_CUDA(cudaMemcpyAsync(se + r0, dse[b2].p, sizeof(double) * rows, cudaMemcpyDeviceToHost, sx));
            FRK_CUDA(cudaEventRecord(ev_down[b2], sx));
        };
        upload(0);
        compute(0);
        for (int c = 0; c < nchunk; ++c) {
            if (c + 1 < nchunk) {
                upload(c + 1);
                compute(c + 1);
            }
            download(c);
        }
        FRK_CUDA(cudaStreamSynchronize(sc));
        FRK_CUDA(cudaStreamSynchronize(sx));
        mp->launches += derived.p->launches;
    });
}

int32_t frk_basis_predict_frk(const double* F, int64_t n, int32_t K, const double* w, int32_t T, const double* V,
                              double* yhat, double* se) {
    return guarded([&] {
        require_device();
        if (n <= 0) return;
        if (K < 1 || T < 1 || !F || !w || !yhat) frk::fail(FRK_EINVAL, "F (n x K), w (K x T) and yhat (n x T) are required");
        cudaStream_t st = nullptr;
        const bool want_se = (V != nullptr && se != nullptr);
        // one product basis %*% [w | R] per row chunk (panel kernel) instead of basis %*% w plus the n x K x K product
        // basis %*% V: q = T + rank(V) columns
        std::vector<double> sgn;
        frk::DevBuf<double> dB;
        const int r = coefficient_matrix_dev(w, T, want_se ? V : nullptr, K, st, dB, sgn);
        const int q = T + r;
        int64_t rc = std::min<int64_t>(n, std::max<int64_t>(1024, (int64_t(512) << 20) / (int64_t(K) * 8)));
        rc += rc & 1;
        frk::DevBuf<double> dF(size_t(rc) * K), dy(size_t(rc) * q), dse, dsgn;
        if (want_se) {
            dse.alloc(static_cast<size_t>(rc));
            dsgn.alloc(static_cast<size_t>(std::max(r, 1)));
            if (r > 0) FRK_CUDA(cudaMemcpyAsync(dsgn.p, sgn.data(), sizeof(double) * r, cudaMemcpyHostToDevice, st));
        }
        for (int64_t r0 = 0; r0 < n; r0 += rc) {
            const int64_t rows = std::min(rc, n - r0);
            FRK_CUDA(cudaMemcpy2DAsync(dF.p, sizeof(double) * rc, F + r0, sizeof(double) * n, sizeof(double) * rows,
                                       size_t(K), cudaMemcpyHostToDevice, st));
            frk::panel_times_small_dev(dF.p, rows, rc, K, dB.p, K, q, dy.p, rc, st);            // R/autoFRK.R:1216,1220
            FRK_CUDA(cudaMemcpy2DAsync(yhat + r0, sizeof(double) * n, dy.p, sizeof(double) * rc, sizeof(double) * rows,
                                       size_t(T), cudaMemcpyDeviceToHost, st));
            if (want_se) {
                signed_rownorm_kernel<<<unsigned((rows + 255) / 256), 256, 0, st>>>(dy.p + size_t(rc) * T, rows, rc, r, dsgn.p, dse.p);
                FRK_CUDA(cudaGetLastError());
                FRK_CUDA(cudaMemcpyAsync(se + r0, dse.p, sizeof(double) * rows, cudaMemcpyDeviceToHost, st));
            }
            FRK_CUDA(cudaStreamSynchronize(st));
        }
    });
}

}  // extern "C"
