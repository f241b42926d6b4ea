// How fast can one persistent CTA per SM stream a tall column-major matrix (n x ncols doubles, column stride n*8
// bytes) into shared memory with 2-D TMA boxes, as the Gram kernels do?  No arithmetic: a consumer warp waits for each
// stage and re-arms it at once, so the number is the ceiling the access pattern itself allows.
// Knobs: columns per box (8 or 32), rows per stage (contiguous run per column = rows * 8 bytes), stages in flight.
// Output: JSON lines.  Measurement tooling, not product code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_stream_bench tools/tma_stream_bench.cu -lcuda
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
}

struct Args { int ncols, boxc, rows_stage, nst; long long n, rows_cta; int hold_per_row; int row_fastest; int boxr; };

// one warp does everything: arm all stages, then wait / re-arm round robin
__global__ void __launch_bounds__(32, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const Args a, double* sink) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int nboxc = (a.ncols + a.boxc - 1) / a.boxc;          // boxes across the columns
    const int nboxr = a.rows_stage >= a.boxr ? a.rows_stage / a.boxr : 1;   // boxes down the rows of a stage (a box may overlap the next stage)
    const int box_bytes = a.boxc * a.boxr * 8;
    const int stage_bytes = nboxc * nboxr * box_bytes;
    uint64_t* full = (uint64_t*)(base + (size_t)a.nst * stage_bytes);
    const int lane = threadIdx.x;
    if (lane == 0) {
        for (int s = 0; s < a.nst; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const long long r_begin = (long long)blockIdx.x * a.rows_cta;
    long long r_end = r_begin + a.rows_cta;
    if (r_end > a.n) r_end = a.n;
    const int nsteps = r_end > r_begin ? (int)((r_end - r_begin + a.rows_stage - 1) / a.rows_stage) : 0;
    auto issue = [&](int slot, int step) {
        if (lane == 0) mbar_expect(&full[slot], (uint32_t)stage_bytes);
        __syncwarp();
        // box id = rb * nboxc + cb: consecutive 16-row boxes of the same columns are issued close together
        for (int b = lane; b < nboxc * nboxr; b += 32) {
            const int rb = a.row_fastest ? b % nboxr : b / nboxc, cb = a.row_fastest ? b / nboxr : b % nboxc;
            tma2d(base + (size_t)slot * stage_bytes + (size_t)(rb * nboxc + cb) * box_bytes, &tm,
                  (int)(r_begin + (long long)step * a.rows_stage) + rb * a.boxr, cb * a.boxc, &full[slot]);
        }
    };
    for (int i = 0; i < a.nst && i < nsteps; ++i) issue(i, i);
    int slot = 0;
    uint32_t par = 0;
    double acc = 0.0;
    for (int step = 0; step < nsteps; ++step) {
        while (!mbar_try(&full[slot], par)) {}
        acc += *(const double*)(base + (size_t)slot * stage_bytes + lane * 8);
        if (a.hold_per_row > 0) {
            const long long t0 = clock64(), dt = (long long)a.hold_per_row * a.rows_stage;
            while (clock64() - t0 < dt) {}
        }
        __syncwarp();
        if (step + a.nst < nsteps) issue(slot, step + a.nst);
        if (++slot == a.nst) { slot = 0; par ^= 1u; }
    }
    if (acc == 123.456) sink[0] = acc;
}


__device__ __forceinline__ void bulk1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}

// mode 0: one 1-D bulk copy per column per stage (rows_stage * 8 contiguous bytes each, padded pitch in shared memory)
// mode 1: the CTA's rows of ALL columns taken as one contiguous region (what a plain copy would read): chunked bulk copies
__global__ void __launch_bounds__(512, 1) bulk_kernel(const double* __restrict__ src, const Args a, int mode, double* sink) {
    extern __shared__ unsigned char raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    const int pitch = a.rows_stage * 8 + 32;
    const int stage_bytes = mode == 0 ? a.ncols * pitch : a.ncols * a.rows_stage * 8;
    const uint32_t tx = (uint32_t)a.ncols * a.rows_stage * 8;
    uint64_t* full = (uint64_t*)(base + (size_t)a.nst * stage_bytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nst; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long r_begin = (long long)blockIdx.x * a.rows_cta;
    const int nsteps = (int)(a.rows_cta / a.rows_stage);
    if (r_begin + a.rows_cta > a.n) return;     // whole CTAs only (keeps every copy in bounds)
    if (warp != 0) {   // extra warps (blockDim > 32): poll every stage like consumer warps would, nothing else
        int slot = 0;
        uint32_t par = 0;
        for (int step = 0; step < nsteps; ++step) {
            while (!mbar_try(&full[slot], par)) {}
            if (++slot == a.nst) { slot = 0; par ^= 1u; }
        }
        return;
    }
    auto issue = [&](int slot, int step) {
        if (lane == 0) mbar_expect(&full[slot], tx);
        __syncwarp();
        unsigned char* dst = base + (size_t)slot * stage_bytes;
        if (mode == 0) {
            for (int c = lane; c < a.ncols; c += 32)
                bulk1d(dst + (size_t)c * pitch, src + (size_t)c * a.n + r_begin + (long long)step * a.rows_stage,
                       (uint32_t)a.rows_stage * 8, &full[slot]);
        } else {
            const size_t chunk = (size_t)a.rows_stage * 8;   // ncols chunks of this size, contiguous in memory
            const double* s0 = src + ((size_t)blockIdx.x * nsteps + step) * (size_t)a.ncols * a.rows_stage;
            for (int c = lane; c < a.ncols; c += 32) bulk1d(dst + c * chunk, (const unsigned char*)s0 + c * chunk, (uint32_t)chunk, &full[slot]);
        }
    };
    for (int i = 0; i < a.nst && i < nsteps; ++i) issue(i, i);
    int slot = 0;
    uint32_t par = 0;
    double acc = 0.0;
    for (int step = 0; step < nsteps; ++step) {
        while (!mbar_try(&full[slot], par)) {}
        acc += *(const double*)(base + (size_t)slot * stage_bytes + lane * 8);
        if (a.hold_per_row > 0) {   // stand-in for the arithmetic on the stage: the slot is handed back only afterwards
            const long long t0 = clock64(), dt = (long long)a.hold_per_row * a.rows_stage;
            while (clock64() - t0 < dt) {}
        }
        __syncwarp();
        if (step + a.nst < nsteps) issue(slot, step + a.nst);
        if (++slot == a.nst) { slot = 0; par ^= 1u; }
    }
    if (acc == 123.456) sink[0] = acc;
}

__global__ void __launch_bounds__(512) ldg_kernel(const double2* __restrict__ src, size_t n2, double* sink) {
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
        const double2 v = __ldg(src + i);
        acc += v.x + v.y;
    }
    if (acc == 123.456) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncodeFn enc = (EncodeFn)fp;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const long long n = 5000000;
    const int maxcols = 144;
    double *d = nullptr, *sink = nullptr;
    CK(cudaMalloc(&d, sizeof(double) * n * maxcols));
    CK(cudaMemset(d, 0x3f, sizeof(double) * n * maxcols));   // finite non-zero doubles
    CK(cudaMalloc(&sink, 64));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const size_t budget = 200 * 1024;
    printf("[\n");
    bool first = true;
    // 2-D boxes: swizzled 16-row boxes (what the Gram kernels use) against unswizzled boxes with longer inner extent
    for (int ncols : {36, 68})
        for (int boxr : {16, 32, 36, 64, 68})
            for (int hold : {0, 14})
                for (int rows_stage : {32, 64}) {
                    if (rows_stage < boxr - 4) continue;
                    if ((boxr == 36 && rows_stage != 32) || (boxr == 68 && rows_stage != 64)) continue;
                    const int boxc = 32, rowfast = 1;
                    Args a;
                    a.ncols = ncols; a.boxc = boxc; a.rows_stage = rows_stage; a.n = n; a.hold_per_row = hold; a.row_fastest = rowfast; a.boxr = boxr;
                    const int nboxc = (ncols + boxc - 1) / boxc;
                    const size_t stage_bytes = (size_t)nboxc * (boxr > rows_stage ? boxr : rows_stage) * boxc * 8;
                    a.nst = (int)(budget / stage_bytes);
                    if (a.nst < 2) continue;
                    if (a.nst > 24) a.nst = 24;
                    a.rows_cta = ((n + sms - 1) / sms + rows_stage - 1) / rows_stage * rows_stage;
                    CUtensorMap tm;
                    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)ncols};
                    cuuint64_t gstr[1] = {(cuuint64_t)n * 8};
                    cuuint32_t box[2] = {(cuuint32_t)boxr, (cuuint32_t)boxc};
                    cuuint32_t estr[2] = {1, 1};
                    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            boxr == 16 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                        fprintf(stderr, "encode failed boxr %d\n", boxr);
                        continue;
                    }
                    const size_t smem = (size_t)a.nst * stage_bytes + 1024 + 24 * 8 + 64;
                    cudaEvent_t e0, e1;
                    CK(cudaEventCreate(&e0));
                    CK(cudaEventCreate(&e1));
                    float best = 1e9f;
                    for (int rep = 0; rep < 4; ++rep) {
                        CK(cudaEventRecord(e0));
                        stream_kernel<<<sms, 32, smem>>>(tm, a, sink);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        float ms;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        if (rep > 0 && ms < best) best = ms;
                    }
                    CK(cudaGetLastError());
                    printf("%s{\"mode\": \"tensor2d\", \"box_rows\": %d, \"swizzle\": \"%s\", \"ncols\": %d, \"hold_clk_per_row\": %d, \"rows_per_stage\": %d, \"stages\": %d, \"ms\": %.4f, \"gbs\": %.1f, \"hold_only_ms\": %.4f}",
                           first ? "" : ",\n", boxr, boxr == 16 ? "128B" : "none", ncols, hold, rows_stage, a.nst, best,
                           8.0 * n * ncols / best / 1e6, hold * (double)a.rows_cta / 1.965e6);
                    first = false;
                    fflush(stdout);
                }
    for (int grid_mult : {1, 2, 4, 8}) {
        const size_t n2 = (size_t)n * 132 / 2;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        float best = 1e9f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            ldg_kernel<<<sms * grid_mult, 512>>>((const double2*)d, n2, sink);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best) best = ms;
        }
        printf(",\n{\"mode\": \"ldg128_read_only\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"gbs\": %.1f}", grid_mult, best, 16.0 * n2 / best / 1e6);
    }
    printf("\n]\n");
    return 0;
}
