// Plan construction (W packing, Cpoly, zero-column detection) and the K1 launch logic.
#include <algorithm>
#include <cstdlib>
#include <vector>
#include "mrts_plan.cuh"
#include "mrts_fit.cuh"

namespace frk {

namespace {

// flags[c] = 1 if W[:, c] has any non-zero entry (W = UZ[0:m, 0:k], ld = ldu)
__global__ void column_nonzero_kernel(const double* __restrict__ UZ, int64_t ldu, int64_t m, int k,
                                      int* __restrict__ flags) {
    const int c = blockIdx.x;
    if (c >= k) return;
    int any = 0;
    for (int64_t j = threadIdx.x; j < m; j += blockDim.x) any |= (UZ[j + int64_t(c) * ldu] != 0.0);
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) flags[c] = any;
}

// Packed image of one slice: for every chunk of KC knots
//   [q = 0..KC/4)[col = 0..BN)[t = 0..4)   W[chunk*KC + 4q + t][col0 + col]   (0 past m / past ncols)
__global__ void pack_w_kernel(const double* __restrict__ UZ, int64_t ldu, int64_t m, int col0, int ncols, int bn,
                              int kc, int nchunks, double* __restrict__ wpack) {
    const int stage = kc * bn;
    const int64_t total = int64_t(nchunks) * stage;
    for (int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; idx < total;
         idx += int64_t(gridDim.x) * blockDim.x) {
        const int chunk = int(idx / stage);
        const int r = int(idx % stage);
        const int q = r / (bn * 4), rem = r % (bn * 4);
        const int col = rem >> 2, tt = rem & 3;
        const int64_t j = int64_t(chunk) * kc + q * 4 + tt;
        wpack[idx] = (j < m && col < ncols) ? UZ[j + int64_t(col0 + col) * ldu] : 0.0;
    }
}

// knot coordinates as DP-tuples, padded to nchunks*KC knots with knot 0 (finite phi, multiplied by W = 0)
__global__ void pack_knots_kernel(const double* __restrict__ Xu, int64_t m, int d, int dp, int64_t npad,
                                  double* __restrict__ knots) {
    const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (idx >= npad * dp) return;
    int64_t j = idx / dp;
    const int c = int(idx % dp);
    if (j >= m) j = 0;
    knots[idx] = (c < d) ? Xu[j + int64_t(c) * m] : 0.0;
}

// cpoly[a*bn + col] = sum_j BBBH[a, j] * W[j, col0+col]   (src/autoFRK.cpp:324, (BBBH) * UZ.block(0,0,n,k))
__global__ void cpoly_kernel(const double* __restrict__ BBBH, const double* __restrict__ UZ, int64_t ldu,
                             int64_t m, int d, int col0, int ncols, int bn, const double* __restrict__ sub,
                             double* __restrict__ cpoly) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (d + 1) * bn) return;
    const int a = idx / bn, col = idx % bn;
    double s = 0.0;
    if (col < ncols) {
        const double* w = UZ + int64_t(col0 + col) * ldu;
        for (int64_t j = 0; j < m; ++j) s = fma(BBBH[a + j * (d + 1)], w[j], s);
        if (sub) s -= sub[a + int64_t(col0 + col) * (d + 1)];
    }
    cpoly[idx] = s;
}

// sub[0, j] = B[0, j] - sum_a shift_a / nconst_a * B[1+a, j];  sub[1+a, j] = B[1+a, j] / nconst_a
// (the polynomial basis columns 1, (x_a - shift_a)/nconst_a of R/autoFRK.R:1462-1464 written as [1, x] %*% sub)
__global__ void derive_sub_kernel(const double* __restrict__ B, int64_t ldb, int q, int d, double s0, double s1, double s2,
                                  double c0, double c1, double c2, double* __restrict__ sub) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= q) return;
    const double sh[3] = {s0, s1, s2}, nc[3] = {c0, c1, c2};
    const double* b = B + int64_t(j) * ldb;
    double v0 = b[0];
    for (int a = 0; a < d; ++a) {
        v0 -= sh[a] / nc[a] * b[1 + a];
        sub[(1 + a) + int64_t(j) * (d + 1)] = b[1 + a] / nc[a];
    }
    sub[int64_t(j) * (d + 1)] = v0;
}

__global__ void copy_cols_kernel(const double* __restrict__ src, int64_t lds, int64_t rows, int cols, double* __restrict__ dst,
                                 int64_t ldd) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    const int j = blockIdx.y;
    if (i < rows && j < cols) dst[i + int64_t(j) * ldd] = src[i + int64_t(j) * lds];
}

__global__ void zero_columns_kernel(double* out, int64_t ldo, int64_t n2, int c0, int c1) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n2) return;
    for (int c = c0 + blockIdx.y; c < c1; c += gridDim.y) out[int64_t(c) * ldo + i] = 0.0;
}

// out[:,0] = 1, out[:,1+a] = (x[:,a] - shift[a]) / nconst[a]     R/autoFRK.R:1462-1464
__global__ void poly_columns_kernel(const double* __restrict__ x, int64_t ldx, int64_t n2, int d, double s0,
                                    double s1, double s2, double c0, double c1, double c2,
                                    double* __restrict__ out, int64_t ldo) {
    const int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
    if (i >= n2) return;
    out[i] = 1.0;
    out[ldo + i] = (x[i] - s0) / c0;
    if (d > 1) out[2 * ldo + i] = (x[i + ldx] - s1) / c1;
    if (d > 2) out[3 * ldo + i] = (x[i + 2 * ldx] - s2) / c2;
}

const PredictKernelInfo* kernels_for(int d, int* count) {
    switch (d) {
        case 1: return predict_kernels_d1(count);
        case 2: return predict_kernels_d2(count);
        case 3: return predict_kernels_d3(count);
    }
    fail(1, "Unsupported dimension d: %d", d);  // src/autoFRK.cpp:114-116
}

}  // namespace

MrtsPlan* plan_create_dev(const double* dXu, int64_t m, int d, const double* dBBBH, const double* dUZ,
                          int64_t ldu, int k, const double* h_nconst, const double* h_shift,
                          cudaStream_t st, const double* d_cpoly_sub, bool keep_sources) {
    if (d < 1 || d > 3) fail(1, "Unsupported dimension d: %d", d);
    if (m < 1) fail(1, "n must be greater than 1");
    if (k < 1) fail(1, "k must be positive");
    auto* p = new MrtsPlan();
    try {
        FRK_CUDA(cudaGetDevice(&p->device));
        FRK_CUDA(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, p->device));
        p->d = d;
        p->m = m;
        p->k = k;
        if (h_nconst) {
            for (int c = 0; c < d; ++c) p->nconst[c] = h_nconst[c];
            p->has_nconst = true;
        }
        if (h_shift)
            for (int c = 0; c < d; ++c) p->shift[c] = h_shift[c];

        // trailing all-zero columns of W (R passes k = K; the last d+1 columns of UZ[0:m,] are zero,
        // src/autoFRK.cpp:212-215, R/autoFRK.R:1458-1460) are not worth a DMMA
        DevBuf<int> dflags(k);
        column_nonzero_kernel<<<k, 256, 0, st>>>(dUZ, ldu, m, k, dflags.p);
        FRK_CUDA(cudaGetLastError());
        std::vector<int> flags(k);
        FRK_CUDA(cudaMemcpyAsync(flags.data(), dflags.p, sizeof(int) * k, cudaMemcpyDeviceToHost, st));
        FRK_CUDA(cudaStreamSynchronize(st));
        int kc = k;
        while (kc > 0 && !flags[kc - 1]) --kc;
        p->k_compute = kc;
        if (d_cpoly_sub) p->k_compute = kc = k;  // derived plan: the polynomial part alone can make a column non-zero
        if (keep_sources) {
            p->src_Xu.alloc(size_t(m) * d);
            p->src_BBBH.alloc(size_t(d + 1) * m);
            FRK_CUDA(cudaMemcpyAsync(p->src_Xu.p, dXu, sizeof(double) * m * d, cudaMemcpyDeviceToDevice, st));
            FRK_CUDA(cudaMemcpyAsync(p->src_BBBH.p, dBBBH, sizeof(double) * (d + 1) * m, cudaMemcpyDeviceToDevice, st));
            if (kc > 0) {
                p->src_W.alloc(size_t(m) * kc);
                dim3 grid(unsigned((m + 255) / 256), unsigned(kc));
                copy_cols_kernel<<<grid, 256, 0, st>>>(dUZ, ldu, m, kc, p->src_W.p, m);
                FRK_CUDA(cudaGetLastError());
            }
        }
        if (kc == 0) return p;  // nothing to project: X1 is identically zero

        // kernel configuration: smallest padded slice width that holds the columns
        int nk = 0;
        const PredictKernelInfo* tab = kernels_for(d, &nk);
        int max_bn = 0;
        for (int i = 0; i < nk; ++i) max_bn = std::max(max_bn, tab[i].bn);
        const int ntile8 = (kc + 7) / 8;
        const int nslices = (ntile8 * 8 + max_bn - 1) / max_bn;
        const int per_slice = ((ntile8 + nslices - 1) / nslices) * 8;  // columns per slice, multiple of 8
        const PredictKernelInfo* best = nullptr;
        for (int i = 0; i < nk; ++i)  // first match wins among equals: later table entries are experimental variants
            if (tab[i].bn >= per_slice && (!best || tab[i].bn < best->bn || (tab[i].bn == best->bn && tab[i].bm > best->bm)))
                best = &tab[i];
        if (const char* force = std::getenv("FRK_K1_CONFIG")) {  // tuning aid: index into the configuration table
            const int i = std::atoi(force);
            if (i >= 0 && i < nk && tab[i].bn >= per_slice) best = &tab[i];
        }
        if (!best) fail(2, "internal: no K1 configuration for %d columns", per_slice);
        p->ki = best;
        best->prepare();
        p->nchunks = std::max(best->min_chunks, int((m + best->kc - 1) / best->kc));
        p->knots.alloc(size_t(p->nchunks) * best->kc * best->dp);
        {
            const int64_t npad = int64_t(p->nchunks) * best->kc;
            pack_knots_kernel<<<unsigned((npad * best->dp + 255) / 256), 256, 0, st>>>(dXu, m, d, best->dp, npad,
                                                                                     p->knots.p);
            FRK_CUDA(cudaGetLastError());
        }

        for (int s = 0; s < nslices; ++s) {
            PredictSlice sl;
            sl.col0 = s * per_slice;
            sl.ncols = std::min(per_slice, kc - sl.col0);
            if (sl.ncols <= 0) break;
            sl.wpack.alloc(size_t(p->nchunks) * best->stage_doubles);
            sl.cpoly.alloc(size_t(d + 1) * best->bn);
            pack_w_kernel<<<p->sm_count * 4, 256, 0, st>>>(dUZ, ldu, m, sl.col0, sl.ncols, best->bn, best->kc,
                                                           p->nchunks, sl.wpack.p);
            FRK_CUDA(cudaGetLastError());
            const int nthr = (d + 1) * best->bn;
            cpoly_kernel<<<(nthr + 127) / 128, 128, 0, st>>>(dBBBH, dUZ, ldu, m, d, sl.col0, sl.ncols, best->bn,
                                                            d_cpoly_sub, sl.cpoly.p);
            FRK_CUDA(cudaGetLastError());
            p->slices.push_back(std::move(sl));
        }
        FRK_CUDA(cudaStreamSynchronize(st));
    } catch (...) {
        delete p;
        throw;
    }
    return p;
}

void plan_destroy(MrtsPlan* p) { delete p; }

MrtsPlan* plan_derive_dev(MrtsPlan* p, const double* dB, int q, cudaStream_t st) {
    if (!p->has_nconst) fail(1, "plan was created without nconst/shift: basis output unavailable");
    if (!p->src_Xu.p) fail(2, "internal: plan keeps no sources to derive from");
    if (q < 1) fail(1, "the coefficient matrix needs at least one column");
    const int d = p->d, K = p->k;
    const int kstar = K - d - 1;
    if (kstar < 0) fail(1, "k-1 can not be smaller than the number of dimensions!");
    const int kc = std::min(kstar, p->k_compute);
    const int64_t m = p->m;
    DevBuf<double> Wd(size_t(m) * q), sub(size_t(d + 1) * q);
    if (kc > 0) {  // W' = W[:, 0:kc] %*% B[(d+1):(d+1+kc), ]
        LinalgHandles& lh = linalg_handles();
        FRK_CUBLAS(cublasSetStream(lh.blas, st));
        const double one = 1.0, zero = 0.0;
        FRK_CUBLAS(cublasDgemm(lh.blas, CUBLAS_OP_N, CUBLAS_OP_N, int(m), q, kc, &one, p->src_W.p, int(m), dB + (d + 1), K, &zero,
                               Wd.p, int(m)));
    } else {
        FRK_CUDA(cudaMemsetAsync(Wd.p, 0, sizeof(double) * m * q, st));
    }
    derive_sub_kernel<<<(q + 127) / 128, 128, 0, st>>>(dB, K, q, d, p->shift[0], p->shift[1], p->shift[2], p->nconst[0],
                                                      p->nconst[1], p->nconst[2], sub.p);
    FRK_CUDA(cudaGetLastError());
    MrtsPlan* out = plan_create_dev(p->src_Xu.p, m, d, p->src_BBBH.p, Wd.p, m, q, nullptr, nullptr, st, sub.p, false);
    FRK_CUDA(cudaStreamSynchronize(st));  // Wd / sub are temporaries
    return out;
}

// first `kout` columns of X1 into dout (columns past k_compute are exact zeros)
static void predict_cols(MrtsPlan* p, const double* dxnew, int64_t n2, int64_t ldx, double* dout, int64_t ldo,
                         int kout, cudaStream_t st) {
    if (n2 <= 0 || kout <= 0) return;
    const int kc = std::min(kout, p->k_compute);
    if (kc < kout) {
        dim3 grid(unsigned((n2 + 255) / 256), unsigned(std::min(kout - kc, 64)));
        zero_columns_kernel<<<grid, 256, 0, st>>>(dout, ldo, n2, kc, kout);
        FRK_CUDA(cudaGetLastError());
        ++p->launches;
    }
    if (kc == 0) return;
    const PredictKernelInfo* ki = p->ki;
    const int64_t ntiles = (n2 + ki->bm - 1) / ki->bm;
    const int grid = int(std::min<int64_t>(ntiles, p->sm_count));
    if (((ntiles + grid - 1) / grid) * int64_t(p->nchunks) >= (int64_t(1) << 31))
        fail(1, "xnew has too many rows for one launch (%lld): evaluate it in row blocks", (long long)n2);
    for (auto& sl : p->slices) {
        const int ncols = std::min(sl.ncols, kc - sl.col0);
        if (ncols <= 0) break;
        PredictArgs a;
        a.xnew = dxnew;
        a.ldx = ldx;
        a.n2 = n2;
        a.wpack = sl.wpack.p;
        a.knots = p->knots.p;
        a.nchunks = p->nchunks;
        a.cpoly = sl.cpoly.p;
        a.out = dout + int64_t(sl.col0) * ldo;
        a.ldo = ldo;
        a.ncols = ncols;
        a.ntiles = ntiles;
        ki->launch(a, grid, st);
        FRK_CUDA(cudaGetLastError());
        ++p->launches;
    }
}

void plan_predict_x1_dev(MrtsPlan* p, const double* dxnew, int64_t n2, int64_t ldx, double* dout, int64_t ldo,
                         cudaStream_t st) {
    predict_cols(p, dxnew, n2, ldx, dout, ldo, p->k, st);
}

void plan_predict_basis_dev(MrtsPlan* p, const double* dxnew, int64_t n2, int64_t ldx, double* dout,
                            int64_t ldo, cudaStream_t st) {
    if (n2 <= 0) return;
    if (!p->has_nconst) fail(1, "plan was created without nconst/shift: basis output unavailable");
    const int kstar = p->k - p->d - 1;  // R/autoFRK.R:1449 (the plan's k is K = NCOL(object))
    if (kstar < 0) fail(1, "k-1 can not be smaller than the number of dimensions!");
    poly_columns_kernel<<<unsigned((n2 + 255) / 256), 256, 0, st>>>(dxnew, ldx, n2, p->d, p->shift[0], p->shift[1],
                                                                   p->shift[2], p->nconst[0], p->nconst[1],
                                                                   p->nconst[2], dout, ldo);
    FRK_CUDA(cudaGetLastError());
    ++p->launches;
    // X1[, 1:kstar] lands right after the d+1 polynomial columns (R/autoFRK.R:1460,1466)
    predict_cols(p, dxnew, n2, ldx, dout + int64_t(p->d + 1) * ldo, ldo, kstar, st);
}

}  // namespace frk
