// Tall panel times small matrix on the FP64 tensor pipe:  Y (rows x q) = F (rows x K) * B (K x q), F column-major with
// rows in the millions and q, K at most a few hundred.  It replaces the library GEMMs behind
//   L = Fk %*% LA                       R/autoFRK.R:387-391   (frk_cMLE)
//   L %*% inner                         R/autoFRK.R:849       (frk_invCz)
//   basis %*% w, basis %*% R            R/autoFRK.R:1216,1220 (frk_basis_predict_frk; V = R diag(+-1) R')
// Every call site feeds it row chunks that arrive from host memory, so it is bound by that copy, not by the kernel;
// the kernel is kept simple: no shared memory, operands straight from global memory / L1.
//   * a warp owns 16 rows and 64 output columns: 2 x 8 DMMA.8x8x4 accumulators;
//   * the MMA m index is permuted (fragment 0 holds rows 2g, fragment 1 rows 2g + 1) so that a lane's two A values
//     are one 16-byte load and a warp's A loads are four 128-byte lines, one per column of F;
//   * B fragments (B[k + t][c + g]) come through the read-only path: B is at most a few MB and stays in L1 / L2.
#include "panel_gemm.cuh"

namespace frk {

namespace {

constexpr int PG_WARPS = 8;
constexpr int PG_NT = 8;   // 8-column groups per warp pass

__device__ __forceinline__ void dmma_pg(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool ALIGNED>
__global__ void __launch_bounds__(PG_WARPS * 32)
panel_gemm_kernel(const double* __restrict__ F, int64_t rows, int64_t ldf, int K, const double* __restrict__ B, int64_t ldb,
                  int q, double* __restrict__ Y, int64_t ldy) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t r0 = (int64_t(blockIdx.x) * PG_WARPS + warp) * 16;
    if (r0 >= rows) return;
    const int c0 = blockIdx.y * (PG_NT * 8);
    const int64_t ra = r0 + 2 * g;                 // this lane's rows: ra (fragment 0), ra + 1 (fragment 1)
    const bool ok0 = ra < rows, ok1 = ra + 1 < rows;
    double acc[2][PG_NT][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < PG_NT; ++j) acc[h][j][0] = acc[h][j][1] = 0.0;
    bool colok[PG_NT];
#pragma unroll
    for (int j = 0; j < PG_NT; ++j) colok[j] = c0 + 8 * j + g < q;

    for (int k = 0; k < K; k += 4) {
        const bool kok = k + t < K;
        double a0 = 0.0, a1 = 0.0;
        const double* pa = F + int64_t(k + t) * ldf + ra;
        if (ALIGNED) {
            if (kok && ok1) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(pa));
                a0 = v.x;
                a1 = v.y;
            } else if (kok && ok0) {
                a0 = __ldg(pa);
            }
        } else {
            if (kok && ok0) a0 = __ldg(pa);
            if (kok && ok1) a1 = __ldg(pa + 1);
        }
        double b[PG_NT];
#pragma unroll
        for (int j = 0; j < PG_NT; ++j) b[j] = (kok && colok[j]) ? __ldg(B + int64_t(c0 + 8 * j + g) * ldb + k + t) : 0.0;
#pragma unroll
        for (int j = 0; j < PG_NT; ++j) {
            dmma_pg(acc[0][j][0], acc[0][j][1], a0, b[j]);
            dmma_pg(acc[1][j][0], acc[1][j][1], a1, b[j]);
        }
    }
    // C fragment: (row g, columns 2t, 2t + 1) of each 8 x 8 tile
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t r = ra + h;
        if (r >= rows) continue;
#pragma unroll
        for (int j = 0; j < PG_NT; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + 8 * j + 2 * t + e;
                if (c < q) Y[r + int64_t(c) * ldy] = acc[h][j][e];
            }
    }
}

}  // namespace

void panel_times_small_dev(const double* dF, int64_t rows, int64_t ldf, int K, const double* dB, int64_t ldb, int q,
                           double* dY, int64_t ldy, cudaStream_t st) {
    if (rows <= 0 || q <= 0) return;
    if (K < 1) fail(1, "panel product: K must be positive");
    const dim3 grid(unsigned((rows + 16 * PG_WARPS - 1) / (16 * PG_WARPS)), unsigned((q + PG_NT * 8 - 1) / (PG_NT * 8)));
    const bool aligned = (reinterpret_cast<uintptr_t>(dF) % 16 == 0) && (ldf % 2 == 0);
    if (aligned) panel_gemm_kernel<true><<<grid, PG_WARPS * 32, 0, st>>>(dF, rows, ldf, K, dB, ldb, q, dY, ldy);
    else panel_gemm_kernel<false><<<grid, PG_WARPS * 32, 0, st>>>(dF, rows, ldf, K, dB, ldb, q, dY, ldy);
    FRK_CUDA(cudaGetLastError());
}

}  // namespace frk
