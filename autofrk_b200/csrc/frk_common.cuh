// Shared device/host helpers for the frk_b200 CUDA library (sm_100a only).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>
#include <stdexcept>
#include <cuda_runtime.h>

namespace frk {

// ---------------------------------------------------------------------------------------------
// error plumbing: C++ exceptions inside the library, converted to (status, message) at the C ABI
// (the reference converts Rcpp::stop() to R errors at src/RcppExports.cpp:18,24 -- same idea).
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw Error(code, buf);
}

#define FRK_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess)                                                                   \
            ::frk::fail(3, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__,      \
                        __LINE__, #expr);                                                        \
    } while (0)

// Size-class cache for device allocations (per host thread and device).  cudaMalloc / cudaFree synchronise the device
// and become very slow once peer mappings exist (symmetric memory, NCCL), while the K x K stages allocate dozens of
// small temporaries per call; cached blocks are reused across calls and released by frk_release_cached_memory().
void* pool_alloc(size_t bytes);
void pool_free(void* p, size_t bytes);
void pool_release_all();

// RAII device buffer backed by the cache
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) p = static_cast<T*>(pool_alloc(count * sizeof(T)));
    }
    void release() {
        if (p) pool_free(p, n * sizeof(T));
        p = nullptr;
        n = 0;
    }
};

// Scratch that is reused by consecutive asynchronous calls (the Gram workspace of a host thread, a plan's chunk
// buffers) is safe only if those calls are ordered.  Calls on ONE stream are; for a caller that alternates streams the
// guard orders them explicitly: every call waits for the event recorded at the end of the previous one when it runs
// on another stream, and buffers are only re-allocated after that event has completed.
struct StreamGuard {
    cudaEvent_t ev = nullptr;
    cudaStream_t last = nullptr;
    bool used = false;
    StreamGuard() = default;
    StreamGuard(const StreamGuard&) = delete;
    StreamGuard& operator=(const StreamGuard&) = delete;
    ~StreamGuard() {
        if (ev) cudaEventDestroy(ev);
    }
    void enter(cudaStream_t st) {
        if (used && st != last) FRK_CUDA(cudaStreamWaitEvent(st, ev, 0));
    }
    void leave(cudaStream_t st) {
        if (!ev) FRK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        FRK_CUDA(cudaEventRecord(ev, st));
        last = st;
        used = true;
    }
    void quiesce() {   // before a buffer that queued work may still use is released
        if (used) FRK_CUDA(cudaEventSynchronize(ev));
    }
    template <typename T>
    void grow(DevBuf<T>& b, size_t need) {
        if (b.n < need) {
            quiesce();
            b.alloc(need);
        }
    }
};
struct StreamGuardScope {
    StreamGuard& g;
    cudaStream_t st;
    StreamGuardScope(StreamGuard& g_, cudaStream_t st_) : g(g_), st(st_) { g.enter(st); }
    ~StreamGuardScope() {
        try {
            g.leave(st);
        } catch (...) {
        }
    }
};

// ---------------------------------------------------------------------------------------------
// device primitives
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// FP64 tensor-core MMA.  On sm_100a every f64 mma.sync shape lowers to DMMA.8x8x4 (checked with
// cuobjdump), so the native shape is used directly.  Fragment layout (g = lane>>2, t = lane&3):
//   A (8x4, row)  : a  = A[g][t]
//   B (4x8, col)  : b  = B[t][g]
//   C/D (8x8)     : c0 = C[g][2t], c1 = C[g][2t+1]
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ---------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// non-blocking probe of a phase (never suspends the warp)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// plain arrive (count 1), release semantics at CTA scope
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory counter (relaxed, CTA scope); returns the old value.  Used only to elect the warp that
// arrives last; the data it guards was already consumed through register dependencies.
__device__ __forceinline__ uint32_t smem_add_relaxed(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}
// global -> shared bulk copy through the TMA unit, completion signalled on an mbarrier.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                             uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

#endif  // __CUDACC__

}  // namespace frk
