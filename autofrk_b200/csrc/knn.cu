// Nearest-neighbour mean imputation of missing observations: the `method = "fast"` pre-step of autoFRK() and
// selectBasis() (R/autoFRK.R:135-148, :273-286), which call FNN::get.knnx(loc[observed, ], loc[missing, ], k) and
// replace each missing value by the mean of its k nearest observed values.  Exact brute force on the device: one
// thread per query keeps its k best squared distances while the reference points stream through shared memory
// (ties go to the lower index; FNN's tie order is unspecified).
#include <cmath>
#include <vector>
#include "capi_internal.cuh"

using frk_capi::guarded;
using frk_capi::require_device;

namespace {

constexpr int KMAX = 16;
constexpr int TILE = 256;

__global__ void __launch_bounds__(TILE) knn_mean_kernel(const double* __restrict__ q, int64_t nq, const double* __restrict__ r,
                                                        const double* __restrict__ rval, int64_t nr, int k, double* __restrict__ out) {
    __shared__ double sx[TILE], sy[TILE], sz[TILE], sv[TILE];
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const bool live = i < nq;
    const double x = live ? q[3 * i] : 0.0, y = live ? q[3 * i + 1] : 0.0, z = live ? q[3 * i + 2] : 0.0;
    double bd[KMAX], bv[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) { bd[j] = INFINITY; bv[j] = 0.0; }
    for (int64_t base = 0; base < nr; base += TILE) {
        const int64_t j = base + threadIdx.x;
        if (j < nr) {
            sx[threadIdx.x] = r[3 * j]; sy[threadIdx.x] = r[3 * j + 1]; sz[threadIdx.x] = r[3 * j + 2]; sv[threadIdx.x] = rval[j];
        }
        __syncthreads();
        const int cnt = int(min(int64_t(TILE), nr - base));
        if (live) {
            for (int c = 0; c < cnt; ++c) {
                const double dx = x - sx[c], dy = y - sy[c], dz = z - sz[c];
                const double d2 = fma(dx, dx, fma(dy, dy, dz * dz));
                if (d2 < bd[KMAX - 1]) {        // candidate: insert into the sorted list (only the first k slots matter)
                    double cd = d2, cv = sv[c];
#pragma unroll
                    for (int s = 0; s < KMAX; ++s) {
                        if (cd < bd[s]) { const double td = bd[s], tv = bv[s]; bd[s] = cd; bv[s] = cv; cd = td; cv = tv; }
                    }
                }
            }
        }
        __syncthreads();
    }
    if (live) {
        double s = 0.0;
        for (int j = 0; j < k; ++j) s += bv[j];
        out[i] = s / double(k);
    }
}

}  // namespace

extern "C" {

int32_t frk_knn_impute(const double* loc, int64_t n, int32_t d, double* Data, int32_t T, int32_t k) {
    return guarded([&] {
        require_device();
        if (d < 1 || d > 3) frk::fail(FRK_EINVAL, "Unsupported dimension d: %d", d);
        if (k < 1 || k > KMAX) frk::fail(FRK_EINVAL, "n.neighbor must be between 1 and %d", KMAX);
        std::vector<double> hq, hr, hv, ho;
        std::vector<int64_t> where;
        frk::DevBuf<double> dq, dr, dv, dout;
        for (int t = 0; t < T; ++t) {
            double* col = Data + int64_t(t) * n;
            hq.clear(); hr.clear(); hv.clear(); where.clear();
            for (int64_t i = 0; i < n; ++i) {
                std::vector<double>& dst = std::isnan(col[i]) ? hq : hr;
                for (int c = 0; c < 3; ++c) dst.push_back(c < d ? loc[i + int64_t(c) * n] : 0.0);
                if (std::isnan(col[i])) where.push_back(i); else hv.push_back(col[i]);
            }
            const int64_t nq = int64_t(where.size()), nr = int64_t(hv.size());
            if (nq == 0) continue;                                            // :139-141
            if (nr < k) frk::fail(FRK_EINVAL, "replicate %d has fewer observed values (%lld) than n.neighbor (%d)", t + 1,
                                  (long long)nr, k);
            if (dq.n < size_t(3 * nq)) dq.alloc(3 * nq);
            if (dout.n < size_t(nq)) dout.alloc(nq);
            if (dr.n < size_t(3 * nr)) dr.alloc(3 * nr);
            if (dv.n < size_t(nr)) dv.alloc(nr);
            FRK_CUDA(cudaMemcpy(dq.p, hq.data(), sizeof(double) * 3 * nq, cudaMemcpyHostToDevice));
            FRK_CUDA(cudaMemcpy(dr.p, hr.data(), sizeof(double) * 3 * nr, cudaMemcpyHostToDevice));
            FRK_CUDA(cudaMemcpy(dv.p, hv.data(), sizeof(double) * nr, cudaMemcpyHostToDevice));
            knn_mean_kernel<<<unsigned((nq + TILE - 1) / TILE), TILE>>>(dq.p, nq, dr.p, dv.p, nr, k, dout.p);
            FRK_CUDA(cudaGetLastError());
            ho.resize(nq);
            FRK_CUDA(cudaMemcpy(ho.data(), dout.p, sizeof(double) * nq, cudaMemcpyDeviceToHost));
            for (int64_t j = 0; j < nq; ++j) col[where[j]] = ho[j];            // Data[where, tt] <- rowMeans(nnval)  :146
        }
    });
}

}  // extern "C"
