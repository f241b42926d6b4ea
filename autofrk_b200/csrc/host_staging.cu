// R hands `.Call` pageable vectors.  cudaMemcpy from pageable memory goes through the driver's single staging thread:
// 11-12 GB/s on this pool's boxes against 50 GB/s from pinned memory (profiles/e2e_probe_r02.json).  Large inputs
// (the n x T data matrix of a fit, an n x K basis handed to frk_gram_stats) are therefore gathered by a few host threads
// into a two-slot pinned ring owned by the calling thread and sent from there, the gather of piece i+1 overlapping the
// transfer of piece i.  Small or already pinned inputs take the plain cudaMemcpy2DAsync.
// FRK_HOST_STAGING=0 turns the staging off; FRK_HOST_COPY_THREADS sets the thread count.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include "host_staging.cuh"

namespace frk {

bool host_pointer_is_pinned(const void* p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

int host_copy_threads() {   // default: the host cores divided among the visible GPUs (one process each), 2..8
    static const int n = [] {
        if (const char* e = std::getenv("FRK_HOST_COPY_THREADS")) return std::max(1, std::atoi(e));
        int ndev = 1;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess) {
            cudaGetLastError();
            ndev = 1;
        }
        const int hw = int(std::thread::hardware_concurrency());
        return std::max(2, std::min(8, hw / std::max(1, ndev)));
    }();
    return n;
}

namespace {

constexpr size_t kPieceBytes = size_t(128) << 20;

struct UploadRing {   // per host thread; lives until the thread ends
    void* slot[2] = {nullptr, nullptr};
    cudaEvent_t sent[2] = {nullptr, nullptr};
    bool busy[2] = {false, false};
    int device = -1;
    ~UploadRing() {
        for (int i = 0; i < 2; ++i) {
            if (sent[i]) cudaEventDestroy(sent[i]);
            if (slot[i]) cudaFreeHost(slot[i]);
        }
    }
    void ensure() {
        int dev = 0;
        FRK_CUDA(cudaGetDevice(&dev));
        if (slot[0] && dev == device) return;
        for (int i = 0; i < 2; ++i) {
            if (!slot[i]) FRK_CUDA(cudaHostAlloc(&slot[i], kPieceBytes, cudaHostAllocPortable));
            if (sent[i]) cudaEventDestroy(sent[i]);
            FRK_CUDA(cudaEventCreateWithFlags(&sent[i], cudaEventDisableTiming));
            busy[i] = false;
        }
        device = dev;
    }
};

void parallel_memcpy(void* dst, const void* src, size_t bytes, int nthreads) {
    nthreads = int(std::max<size_t>(1, std::min<size_t>(size_t(nthreads), bytes >> 20)));
    if (nthreads == 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    const size_t per = ((bytes + nthreads - 1) / nthreads + 63) & ~size_t(63);
    for (int i = 0; i < nthreads; ++i) {
        const size_t b0 = std::min(bytes, per * i), b1 = std::min(bytes, per * (i + 1));
        if (b1 > b0) th.emplace_back([=] { std::memcpy(static_cast<char*>(dst) + b0, static_cast<const char*>(src) + b0, b1 - b0); });
    }
    for (std::thread& t : th) t.join();
}

}  // namespace

void upload_matrix(double* d_dst, int64_t ld_d, const double* h_src, int64_t ld_h, int64_t rows, int64_t cols, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return;
    static const int staging_env = [] {
        const char* e = std::getenv("FRK_HOST_STAGING");
        return e ? std::atoi(e) : 1;
    }();
    const size_t total = sizeof(double) * size_t(rows) * size_t(cols);
    if (staging_env == 0 || total < (size_t(64) << 20) || host_pointer_is_pinned(h_src)) {
        FRK_CUDA(cudaMemcpy2DAsync(d_dst, sizeof(double) * ld_d, h_src, sizeof(double) * ld_h, sizeof(double) * rows, size_t(cols),
                                   cudaMemcpyHostToDevice, st));
        return;
    }
    static thread_local UploadRing ring;
    ring.ensure();
    const int nthreads = host_copy_threads();
    // pieces: runs of whole columns when a column fits a slot, else row ranges of one column
    const int64_t max_rows = int64_t(kPieceBytes / sizeof(double));
    int k = 0;
    auto send = [&](const double* src, double* dst, int64_t r, int64_t c, int64_t ldh) {   // r x c block, c == 1 unless ldh-strided
        const int s = k++ & 1;
        if (ring.busy[s]) FRK_CUDA(cudaEventSynchronize(ring.sent[s]));
        double* stage = static_cast<double*>(ring.slot[s]);
        if (c == 1 || ldh == r) {
            parallel_memcpy(stage, src, sizeof(double) * size_t(r) * size_t(c), nthreads);
        } else {   // columns of a strided matrix, gathered densely
            std::vector<std::thread> th;
            const int nt = int(std::min<int64_t>(nthreads, c));
            for (int t = 0; t < nt; ++t)
                th.emplace_back([=] {
                    for (int64_t j = c * t / nt; j < c * (t + 1) / nt; ++j)
                        std::memcpy(stage + j * r, src + j * ldh, sizeof(double) * size_t(r));
                });
            for (std::thread& t : th) t.join();
        }
        FRK_CUDA(cudaMemcpy2DAsync(dst, sizeof(double) * ld_d, stage, sizeof(double) * r, sizeof(double) * r, size_t(c),
                                   cudaMemcpyHostToDevice, st));
        FRK_CUDA(cudaEventRecord(ring.sent[s], st));
        ring.busy[s] = true;
    };
    if (rows <= max_rows) {
        const int64_t cper = std::max<int64_t>(1, max_rows / rows);
        for (int64_t c0 = 0; c0 < cols; c0 += cper) {
            const int64_t c = std::min(cper, cols - c0);
            send(h_src + c0 * ld_h, d_dst + c0 * ld_d, rows, c, ld_h);
        }
    } else {
        for (int64_t c0 = 0; c0 < cols; ++c0)
            for (int64_t r0 = 0; r0 < rows; r0 += max_rows)
                send(h_src + c0 * ld_h + r0, d_dst + c0 * ld_d + r0, std::min(max_rows, rows - r0), 1, ld_h);
    }
}

}  // namespace frk
