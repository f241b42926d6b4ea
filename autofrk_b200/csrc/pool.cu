// Device-memory cache behind frk::DevBuf (see frk_common.cuh).
#include <map>
#include <mutex>
#include <vector>
#include "frk_common.cuh"

namespace frk {

namespace {

struct Pool {
    std::mutex mu;
    std::map<std::pair<int, size_t>, std::vector<void*>> free_blocks;  // (device, size class) -> blocks
    size_t cached_bytes = 0;
};
Pool& pool() {
    static Pool p;
    return p;
}
// size classes: powers of two up to 1 MB, then multiples of 1 MB
size_t size_class(size_t bytes) {
    if (bytes <= 256) return 256;
    if (bytes <= (size_t(1) << 20)) {
        size_t c = 256;
        while (c < bytes) c <<= 1;
        return c;
    }
    const size_t mb = size_t(1) << 20;
    return (bytes + mb - 1) / mb * mb;
}
constexpr size_t kMaxCachedBlock = size_t(256) << 20;  // bigger blocks go straight back to the driver

}  // namespace

void* pool_alloc(size_t bytes) {
    int dev = 0;
    FRK_CUDA(cudaGetDevice(&dev));
    const size_t cls = size_class(bytes);
    if (cls <= kMaxCachedBlock) {
        Pool& P = pool();
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.free_blocks.find({dev, cls});
        if (it != P.free_blocks.end() && !it->second.empty()) {
            void* p = it->second.back();
            it->second.pop_back();
            P.cached_bytes -= cls;
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, cls);
    if (e != cudaSuccess) {  // give cached memory back and retry once
        cudaGetLastError();
        pool_release_all();
        e = cudaMalloc(&p, cls);
    }
    if (e != cudaSuccess) fail(3, "CUDA error %s allocating %zu bytes of device memory", cudaGetErrorString(e), cls);
    return p;
}

void pool_free(void* p, size_t bytes) {
    if (!p) return;
    const size_t cls = size_class(bytes);
    if (cls > kMaxCachedBlock) {
        cudaFree(p);
        return;
    }
    // work queued on a stream may still use the block: blocks only return to the cache (never to the driver) here,
    // and every user of DevBuf in this library is stream-ordered on the stream that will reuse it or synchronises
    // before the buffer goes out of scope.
    cudaPointerAttributes attr;
    int dev = 0;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) dev = attr.device;
    else cudaGetLastError();
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    P.free_blocks[{dev, cls}].push_back(p);
    P.cached_bytes += cls;
}

void pool_release_all() {
    Pool& P = pool();
    std::lock_guard<std::mutex> lk(P.mu);
    for (auto& kv : P.free_blocks)
        for (void* p : kv.second) cudaFree(p);
    P.free_blocks.clear();
    P.cached_bytes = 0;
}

}  // namespace frk
