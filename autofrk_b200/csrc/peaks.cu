// Measurement aid: the FP64 tensor-pipe ceiling of the device this library runs on, taken with the instruction every
// kernel here is built from (DMMA.8x8x4, what all f64 mma.sync shapes lower to on sm_100a).  bench.py reports it next
// to the cuBLAS DGEMM figure: cuBLAS picks an sm_80 CUTLASS DGEMM on this toolkit, so "DGEMM peak" is a library
// artefact, while this number is the issue rate of the pipe itself (16 independent accumulators per warp, 16 warps
// per SM, one CTA per SM, no memory traffic).
#include "capi_internal.cuh"

namespace {

__global__ void __launch_bounds__(512, 1) dmma_issue_kernel(double* out, int iters, double seed) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        c[i][0] = seed * i;
        c[i][1] = seed + i;
    }
    const double a = seed + threadIdx.x * 1e-9, b = seed * 0.5 + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int32_t frk_measure_dmma_peak(int32_t iters, double* tflops) {
    return frk_capi::guarded([&] {
        frk_capi::require_device();
        if (iters < 1 || !tflops) frk::fail(FRK_EINVAL, "iters must be positive and tflops non-NULL");
        int dev = 0, sms = 0;
        FRK_CUDA(cudaGetDevice(&dev));
        FRK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        frk::DevBuf<double> out(size_t(sms) * 512);
        cudaEvent_t e0, e1;
        FRK_CUDA(cudaEventCreate(&e0));
        FRK_CUDA(cudaEventCreate(&e1));
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {   // first repetition warms up
            FRK_CUDA(cudaEventRecord(e0, nullptr));
            dmma_issue_kernel<<<sms, 512>>>(out.p, iters, 1.0 + rep);
            FRK_CUDA(cudaEventRecord(e1, nullptr));
            FRK_CUDA(cudaEventSynchronize(e1));
            float ms = 0.f;
            FRK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        FRK_CUDA(cudaGetLastError());
        // one DMMA.8x8x4 = 8 * 8 * 4 multiply-adds = 512 flop per warp instruction
        *tflops = double(sms) * 16.0 * double(iters) * 16.0 * 512.0 / (double(best) * 1e-3) / 1e12;
    });
}
