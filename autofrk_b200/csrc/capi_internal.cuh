// Internals shared by the C-ABI translation units (capi.cu, capi_frk.cu).
#pragma once
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/frk_b200.h"
#include "frk_common.cuh"
#include "mrts_plan.cuh"
#include "mrts_fit.cuh"
#include "gram.cuh"
#include "cmle.cuh"

namespace frk_capi {

std::string& last_error();

template <typename F>
int32_t guarded(F&& f) {
    try {
        f();
        return FRK_OK;
    } catch (const frk::Error& e) {
        last_error() = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        last_error() = "out of host memory";
        return FRK_EINTERNAL;
    } catch (const std::exception& e) {
        last_error() = e.what();
        return FRK_EINTERNAL;
    } catch (...) {
        last_error() = "unknown error";
        return FRK_EINTERNAL;
    }
}

void require_device();
frk::GramWorkspace& gram_workspace();  // one per host thread; calls sharing it must be stream-ordered

}  // namespace frk_capi

struct frk_plan {
    frk::MrtsPlan* mp = nullptr;
    // resources of the host-buffer path (created lazily)
    cudaStream_t s_comp = nullptr, s_copy = nullptr;
    cudaEvent_t ev_comp[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    frk::DevBuf<double> dx, dbuf[2];
    // pinned staging of results that go to pageable host memory (two chunks; see predict_host)
    double* hstage[2] = {nullptr, nullptr};
    size_t hstage_doubles = 0;
    // fused predict -> Gram / predict -> kriging: one basis chunk + packed partial statistics
    frk::DevBuf<double> fchunk, packed_tmp, ytmp;
    frk::StreamGuard guard;   // orders asynchronous calls that reuse these buffers from different streams
    ~frk_plan() {
        for (int i = 0; i < 2; ++i) {
            if (ev_comp[i]) cudaEventDestroy(ev_comp[i]);
            if (ev_copy[i]) cudaEventDestroy(ev_copy[i]);
        }
        for (int i = 0; i < 2; ++i)
            if (hstage[i]) cudaFreeHost(hstage[i]);
        if (s_comp) cudaStreamDestroy(s_comp);
        if (s_copy) cudaStreamDestroy(s_copy);
        if (mp) frk::plan_destroy(mp);
    }
    void ensure_streams() {
        if (s_comp) return;
        FRK_CUDA(cudaStreamCreateWithFlags(&s_comp, cudaStreamNonBlocking));
        FRK_CUDA(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            FRK_CUDA(cudaEventCreateWithFlags(&ev_comp[i], cudaEventDisableTiming));
            FRK_CUDA(cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming));
        }
    }
};

namespace frk_capi {
// capi_frk.cu: fused predict -> Gram with Z in host memory, uploaded chunk by chunk behind K1 (blocks until done)
void predict_gram_host_z(frk_plan* plan, const double* dx, int64_t n2, int64_t ldx, const double* hZ, int64_t ldzh,
                         double* dZ, int T, int64_t ldz, const double* dw, double* dpacked, cudaStream_t st);
}

