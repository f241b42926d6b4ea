// C ABI (include/frk_b200.h).  No exception crosses this boundary: everything is converted to
// (status, thread-local message) -- the C analogue of BEGIN_RCPP/END_RCPP (src/RcppExports.cpp:18,24).
#include <cstdlib>
#include <cstring>
#include <map>
#include <thread>
#include <vector>
#include "capi_internal.cuh"
#include "host_staging.cuh"

namespace frk_capi {

std::string& last_error() {
    static thread_local std::string e;
    return e;
}

void require_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        frk::fail(FRK_ECUDA, "no CUDA device available (frk_b200 has no CPU fallback): %s",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
}

frk::GramWorkspace& gram_workspace() {  // one per host thread and device (the buffers live on that device)
    static thread_local std::map<int, frk::GramWorkspace> ws;
    int dev = 0;
    FRK_CUDA(cudaGetDevice(&dev));
    return ws[dev];
}

}  // namespace frk_capi

using frk_capi::guarded;
using frk_capi::require_device;

namespace frk_capi {

frk_plan* make_plan(const double* Xu, int64_t n, int32_t d, const double* BBBH, const double* UZ,
                    int64_t uz_rows, int64_t uz_cols, const double* nconst, int32_t k, bool on_device) {
    require_device();
    if (d < 1 || d > 3) frk::fail(FRK_EINVAL, "Unsupported dimension d: %d", d);
    if (n < 1) frk::fail(FRK_EINVAL, "n must be greater than 1");
    if (uz_rows < n || uz_cols < k) frk::fail(FRK_EINVAL, "UZ dimensions are inconsistent with n and k");
    if (k < 1) frk::fail(FRK_EINVAL, "k must be positive");
    auto plan = new frk_plan();
    try {
        // colMeans(Xu): R/autoFRK.R:1462
        std::vector<double> hXu(size_t(n) * d);
        double shift[3] = {0, 0, 0};
        frk::DevBuf<double> dXu, dB, dU;
        const double *pXu = Xu, *pB = BBBH, *pU = UZ;
        if (on_device) {
            FRK_CUDA(cudaMemcpy(hXu.data(), Xu, sizeof(double) * n * d, cudaMemcpyDeviceToHost));
        } else {
            std::memcpy(hXu.data(), Xu, sizeof(double) * n * d);
            dXu.alloc(size_t(n) * d);
            dB.alloc(size_t(d + 1) * n);
            dU.alloc(size_t(uz_rows) * k);
            FRK_CUDA(cudaMemcpy(dXu.p, Xu, sizeof(double) * n * d, cudaMemcpyHostToDevice));
            FRK_CUDA(cudaMemcpy(dB.p, BBBH, sizeof(double) * (d + 1) * n, cudaMemcpyHostToDevice));
            FRK_CUDA(cudaMemcpy(dU.p, UZ, sizeof(double) * uz_rows * k, cudaMemcpyHostToDevice));
            pXu = dXu.p;
            pB = dB.p;
            pU = dU.p;
        }
        for (int c = 0; c < d; ++c) {
            long double s = 0.0L;  // R's colMeans accumulates in long double
            for (int64_t j = 0; j < n; ++j) s += hXu[j + size_t(c) * n];
            shift[c] = double(s / (long double)n);
        }
        plan->mp = frk::plan_create_dev(pXu, n, d, pB, pU, uz_rows, k, nconst, shift, nullptr);
    } catch (...) {
        delete plan;
        throw;
    }
    return plan;
}

// R hands `.Call` pageable REALSXP buffers.  A device -> host copy into pageable memory goes through the driver's own
// staging buffer and blocks the calling thread: 11.9 GB/s on this pool's boxes against 50 GB/s into pinned memory
// (profiles/e2e_probe_r02.json).  Page-locking the caller's buffer for the call (cudaHostRegister) pins at ~27 GB/s and
// ends at 17 GB/s overall.  So large results for pageable destinations are staged: the chunk comes down into one of two
// pinned buffers owned by the plan at the PCIe rate while a few host threads scatter the previous chunk's columns into
// the caller's matrix (measured: 25.7 GB/s with 4 threads, 31 GB/s with 8).  FRK_HOST_STAGING=0 turns it off;
// FRK_HOST_COPY_THREADS sets the thread count.
struct ScatterJob {   // columns of one staged chunk (rows x k, dense) -> out[r0 .. r0 + rows, :] with leading dimension ldo
    std::vector<std::thread> threads;
    void start(const double* stage, int64_t rows, int k, double* out, int64_t ldo, int nthreads) {
        join();
        nthreads = std::max(1, std::min(nthreads, k));
        for (int tix = 0; tix < nthreads; ++tix) {
            const int c0 = int(int64_t(k) * tix / nthreads), c1 = int(int64_t(k) * (tix + 1) / nthreads);
            threads.emplace_back([=] {
                for (int c = c0; c < c1; ++c)
                    std::memcpy(out + int64_t(c) * ldo, stage + int64_t(c) * rows, sizeof(double) * size_t(rows));
            });
        }
    }
    void join() {
        for (std::thread& t : threads) t.join();
        threads.clear();
    }
    ~ScatterJob() { join(); }
};

void predict_host(frk_plan* plan, const double* xnew, int64_t n2, double* out, int32_t mode) {
    if (mode != FRK_OUT_X1 && mode != FRK_OUT_BASIS) frk::fail(FRK_EINVAL, "unknown out_mode %d", mode);
    if (n2 <= 0) return;
    frk::MrtsPlan* mp = plan->mp;
    plan->ensure_streams();
    const int d = mp->d, k = mp->k;
    // row chunks: two device buffers of <= ~512 MB each, a multiple of the tile height x SM count
    const int64_t bm = mp->ki ? mp->ki->bm : 128;
    int64_t rc = (int64_t(512) << 20) / (int64_t(k) * 8);
    const int64_t wave = bm * mp->sm_count;
    rc = std::max<int64_t>(wave, rc / wave * wave);
    rc = std::min(rc, n2);
    if (plan->dx.n < size_t(n2) * d) plan->dx.alloc(size_t(n2) * d);
    for (int b = 0; b < 2; ++b)
        if (plan->dbuf[b].n < size_t(rc) * k) plan->dbuf[b].alloc(size_t(rc) * k);
    FRK_CUDA(cudaMemcpyAsync(plan->dx.p, xnew, sizeof(double) * n2 * d, cudaMemcpyHostToDevice, plan->s_comp));
    // The device->host stream is the bottleneck (PCIe): a short first chunk gets it going after two waves of tiles
    // instead of after a full chunk of compute.  The kernels of chunk c+1 are queued BEFORE the copy of chunk c is
    // issued, because a copy to pageable host memory blocks the calling thread.
    std::vector<int64_t> starts;
    for (int64_t r0 = 0; r0 < n2;) {
        starts.push_back(r0);
        r0 += std::min(starts.size() == 1 ? std::min(rc, 2 * wave) : rc, n2 - r0);
    }
    starts.push_back(n2);
    const int nchunk = int(starts.size()) - 1;
    auto compute = [&](int c) {
        const int b = c & 1;
        const int64_t r0 = starts[c], rows = starts[c + 1] - r0;
        if (c >= 2) FRK_CUDA(cudaStreamWaitEvent(plan->s_comp, plan->ev_copy[b], 0));   // chunk c-2 has left the buffer
        if (mode == FRK_OUT_X1)
            frk::plan_predict_x1_dev(mp, plan->dx.p + r0, rows, n2, plan->dbuf[b].p, rc, plan->s_comp);
        else
            frk::plan_predict_basis_dev(mp, plan->dx.p + r0, rows, n2, plan->dbuf[b].p, rc, plan->s_comp);
        FRK_CUDA(cudaEventRecord(plan->ev_comp[b], plan->s_comp));
    };
    static const int staging_env = [] {
        const char* e = std::getenv("FRK_HOST_STAGING");
        return e ? std::atoi(e) : 1;
    }();
    const int copy_threads = frk::host_copy_threads();
    const bool staged = staging_env != 0 && sizeof(double) * size_t(n2) * size_t(k) >= (size_t(64) << 20) &&
                        !frk::host_pointer_is_pinned(out);
    if (staged && plan->hstage_doubles < size_t(rc) * k) {
        for (int b = 0; b < 2; ++b) {
            if (plan->hstage[b]) cudaFreeHost(plan->hstage[b]);
            plan->hstage[b] = nullptr;
            FRK_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&plan->hstage[b]), sizeof(double) * size_t(rc) * k, cudaHostAllocDefault));
        }
        plan->hstage_doubles = size_t(rc) * k;
    }
    ScatterJob scatter[2];
    compute(0);
    for (int c = 0; c < nchunk; ++c) {
        const int b = c & 1;
        const int64_t r0 = starts[c], rows = starts[c + 1] - r0;
        if (c + 1 < nchunk) compute(c + 1);
        FRK_CUDA(cudaStreamWaitEvent(plan->s_copy, plan->ev_comp[b], 0));
        if (staged) {
            scatter[b].join();                                   // chunk c-2 has left this staging buffer
            FRK_CUDA(cudaMemcpy2DAsync(plan->hstage[b], sizeof(double) * rows, plan->dbuf[b].p, sizeof(double) * rc,
                                       sizeof(double) * rows, size_t(k), cudaMemcpyDeviceToHost, plan->s_copy));
            FRK_CUDA(cudaEventRecord(plan->ev_copy[b], plan->s_copy));
            if (c >= 1) {                                        // while chunk c comes down, scatter chunk c-1
                FRK_CUDA(cudaEventSynchronize(plan->ev_copy[b ^ 1]));
                scatter[b ^ 1].start(plan->hstage[b ^ 1], starts[c] - starts[c - 1], k, out + starts[c - 1], n2, copy_threads);
            }
        } else {
            FRK_CUDA(cudaMemcpy2DAsync(out + r0, sizeof(double) * n2, plan->dbuf[b].p, sizeof(double) * rc,
                                       sizeof(double) * rows, size_t(k), cudaMemcpyDeviceToHost, plan->s_copy));
            FRK_CUDA(cudaEventRecord(plan->ev_copy[b], plan->s_copy));
        }
    }
    FRK_CUDA(cudaStreamSynchronize(plan->s_comp));
    FRK_CUDA(cudaStreamSynchronize(plan->s_copy));
    if (staged) {
        const int c = nchunk - 1;
        scatter[c & 1].start(plan->hstage[c & 1], starts[c + 1] - starts[c], k, out + starts[c], n2, copy_threads);
        scatter[0].join();
        scatter[1].join();
    }
}

}  // namespace frk_capi

using frk_capi::make_plan;
using frk_capi::predict_host;

extern "C" {

const char* frk_last_error(void) { return frk_capi::last_error().c_str(); }

int32_t frk_version(void) { return 100; }

int32_t frk_device_count(int32_t* count) {
    return guarded([&] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        *count = n;
    });
}

int32_t frk_release_cached_memory(void) {
    return guarded([&] { frk::pool_release_all(); });
}

int32_t frk_set_device(int32_t device) {
    return guarded([&] {
        require_device();
        FRK_CUDA(cudaSetDevice(device));
    });
}

int32_t frk_plan_create(frk_plan** plan, const double* Xu, int64_t n, int32_t d, const double* BBBH,
                        const double* UZ, int64_t uz_rows, int64_t uz_cols, const double* nconst, int32_t k,
                        int32_t on_device) {
    return guarded([&] {
        if (!plan) frk::fail(FRK_EINVAL, "plan must not be NULL");
        *plan = nullptr;
        *plan = make_plan(Xu, n, d, BBBH, UZ, uz_rows, uz_cols, nconst, k, on_device != 0);
    });
}

int32_t frk_plan_destroy(frk_plan* plan) {
    return guarded([&] { delete plan; });
}

int32_t frk_plan_info(const frk_plan* plan, int64_t info[8]) {
    return guarded([&] {
        if (!plan || !plan->mp) frk::fail(FRK_EINVAL, "plan must not be NULL");
        const frk::MrtsPlan* mp = plan->mp;
        info[0] = mp->d;
        info[1] = mp->m;
        info[2] = mp->k;
        info[3] = mp->k_compute;
        info[4] = mp->ki ? mp->ki->bm : 0;
        info[5] = mp->ki ? mp->ki->bn : 0;
        info[6] = int64_t(mp->slices.size());
        info[7] = mp->nchunks;
    });
}

int64_t frk_plan_launches(const frk_plan* plan) { return (plan && plan->mp) ? plan->mp->launches : 0; }

int32_t frk_plan_predict(frk_plan* plan, const double* xnew, int64_t n2, double* out, int32_t out_mode) {
    return guarded([&] {
        if (!plan || !plan->mp) frk::fail(FRK_EINVAL, "plan must not be NULL");
        predict_host(plan, xnew, n2, out, out_mode);
    });
}

int32_t frk_plan_predict_dev(frk_plan* plan, const double* d_xnew, int64_t n2, int64_t ldx, double* d_out,
                             int64_t ldo, int32_t out_mode, void* stream) {
    return guarded([&] {
        if (!plan || !plan->mp) frk::fail(FRK_EINVAL, "plan must not be NULL");
        if (ldx < n2 || ldo < n2) frk::fail(FRK_EINVAL, "leading dimensions must be >= n2");
        cudaStream_t st = static_cast<cudaStream_t>(stream);
        if (out_mode == FRK_OUT_X1)
            frk::plan_predict_x1_dev(plan->mp, d_xnew, n2, ldx, d_out, ldo, st);
        else if (out_mode == FRK_OUT_BASIS)
            frk::plan_predict_basis_dev(plan->mp, d_xnew, n2, ldx, d_out, ldo, st);
        else
            frk::fail(FRK_EINVAL, "unknown out_mode %d", out_mode);
    });
}

int32_t frk_mrtsrcpp_predict(const double* Xu, int64_t n, int32_t d, const double* xobs_diag, const double* xnew,
                             int64_t n2, int32_t xnew_cols, const double* BBBH, int64_t bbbh_rows,
                             int64_t bbbh_cols, const double* UZ, int64_t uz_rows, int64_t uz_cols,
                             const double* nconst, int64_t nconst_len, int32_t k, double* X1) {
    (void)xobs_diag;  // src/autoFRK.cpp:288: carried by the signature, never read
    return guarded([&] {
        // the reference's checks, in its order, with its messages (src/autoFRK.cpp:298-309, :114-116)
        if (xnew_cols != d) frk::fail(FRK_EINVAL, "xnew must have %d columns", d);
        if (bbbh_rows != d + 1 || bbbh_cols != n) frk::fail(FRK_EINVAL, "BBBH must be %d x %d", d + 1, int(n));
        if (uz_rows < n || uz_cols < k) frk::fail(FRK_EINVAL, "UZ dimensions are inconsistent with n and k");
        if (nconst_len != d) frk::fail(FRK_EINVAL, "nconst must have length %d", d);
        if (d < 1 || d > 3) frk::fail(FRK_EINVAL, "Unsupported dimension d: %d", d);
        if (n2 == 0 || k == 0) return;
        frk_plan* plan = make_plan(Xu, n, d, BBBH, UZ, uz_rows, uz_cols, nconst, k, false);
        try {
            predict_host(plan, xnew, n2, X1, FRK_OUT_X1);
        } catch (...) {
            delete plan;
            throw;
        }
        delete plan;
    });
}

static void fit_to_host(const frk::MrtsFitDev& f, double* X, double* UZ, double* BBBH, double* nconst) {
    const int64_t m = f.m;
    const int d = f.d, K = f.k + f.d + 1;
    FRK_CUDA(cudaMemcpy(X, f.X.p, sizeof(double) * m * K, cudaMemcpyDeviceToHost));
    FRK_CUDA(cudaMemcpy(UZ, f.UZ.p, sizeof(double) * (m + d + 1) * K, cudaMemcpyDeviceToHost));
    FRK_CUDA(cudaMemcpy(BBBH, f.BBBH.p, sizeof(double) * (d + 1) * m, cudaMemcpyDeviceToHost));
    for (int c = 0; c < d; ++c) nconst[c] = f.nconst[c];
}

int32_t frk_mrtsrcpp(const double* Xu, int64_t n, int32_t d, const double* xobs_diag, int32_t xd_rows,
                     int32_t xd_cols, int32_t k, double* X, double* UZ, double* BBBH, double* nconst) {
    return guarded([&] {
        if (xd_rows != d || xd_cols != d) frk::fail(FRK_EINVAL, "xobs_diag must be %d x %d", d, d);  // :231-233
        require_device();
        frk::MrtsFitDev f;
        frk::mrts_fit_dev(Xu, xobs_diag, n, d, k, f);
        fit_to_host(f, X, UZ, BBBH, nconst);
    });
}

int32_t frk_mrtsrcpp_predict0(const double* Xu, int64_t n, int32_t d, const double* xobs_diag, int32_t xd_rows,
                              int32_t xd_cols, const double* xnew, int64_t n2, int32_t xnew_cols, int32_t k,
                              double* X, double* UZ, double* BBBH, double* nconst, double* X1) {
    return guarded([&] {
        if (xd_rows != d || xd_cols != d) frk::fail(FRK_EINVAL, "xobs_diag must be %d x %d", d, d);  // :257-259
        if (xnew_cols != d) frk::fail(FRK_EINVAL, "xnew must have %d columns", d);                     // :260-262
        require_device();
        frk::MrtsFitDev f;
        frk::mrts_fit_dev(Xu, xobs_diag, n, d, k, f);
        fit_to_host(f, X, UZ, BBBH, nconst);
        if (n2 == 0) return;
        frk_plan* plan = new frk_plan();
        try {
            plan->mp = frk::plan_create_dev(f.Xu.p, n, d, f.BBBH.p, f.UZ.p, n + d + 1, k, f.nconst, f.shift, nullptr);
            predict_host(plan, xnew, n2, X1, FRK_OUT_X1);
        } catch (...) {
            delete plan;
            throw;
        }
        delete plan;
    });
}

int32_t frk_getASCeigens(const double* A, int32_t n, double* value, double* vector) {
    return guarded([&] {
        if (n < 1) frk::fail(FRK_EINVAL, "A must be a non-empty square matrix");
        require_device();
        frk::DevBuf<double> dA(size_t(n) * n), dW(static_cast<size_t>(n));
        FRK_CUDA(cudaMemcpy(dA.p, A, sizeof(double) * n * n, cudaMemcpyHostToDevice));
        frk::sym_eigen_dev(dA.p, n, dW.p);
        FRK_CUDA(cudaMemcpy(value, dW.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
        FRK_CUDA(cudaMemcpy(vector, dA.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    });
}

}  // extern "C"
