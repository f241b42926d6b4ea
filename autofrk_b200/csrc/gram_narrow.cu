// K4, narrow variant for few basis functions (K <= 48): G = F' W F, C = F' W Z when the whole upper-triangular tile set
// fits in one warp's registers.  Here the reduction is at the HBM/DMMA ridge (K = 32, T = 4: 288 B and 1,320 credited
// flops per row), so the 128 x 128-tile kernels -- which would do 16x redundant tensor work -- are the wrong shape.
//   * every warp owns a contiguous block of rows and ALL output tiles (MT(MT+1)/2 + MT*NZ DMMA accumulators);
//   * no shared memory, no barriers in the main loop: fragments come straight from global memory as LDG.128
//     (rows 2t, 2t+1 and 8+2t, 8+2t+1 of eight columns per warp instruction = 64 contiguous bytes per column, both
//     halves of each 128 B line back to back), register double-buffered one 16-row group ahead;
//   * the MMA k index is permuted accordingly (the same permutation for both operands, which are the same matrix);
//   * warps are summed through shared memory in warp order, CTAs by gram_narrow_reduce_kernel in CTA order ->
//     bitwise reproducible.
#include <algorithm>
#include "frk_common.cuh"
#include "gram.cuh"

namespace frk {

namespace {

constexpr int NWARPS_N = 8;
constexpr int NTHREADS_N = NWARPS_N * 32;

__device__ __forceinline__ void dmma_n(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

struct NarrowArgs {
    const double* F;
    int64_t ldf;
    int K;
    const double* Z;
    int64_t ldz;
    int T;
    const double* w;
    int64_t n;
    int64_t rows_per_warp;  // multiple of 16
    double* partial;        // [gridDim.x][NTILES * 64]
};

template <int MT, int NZ>
struct NarrowTiles {
    static constexpr int NFF = MT * (MT + 1) / 2;
    static constexpr int NT = NFF + MT * NZ;
};

// one 16-row group of column group `cg` (8 columns starting at 8*cg of F, or of Z for cg >= MT):
// f[0], f[1] = rows 2t, 2t+1; f[2], f[3] = rows 8+2t, 8+2t+1 of column 8*cg + g
template <int MT, int NZ>
__device__ __forceinline__ void load_group(const NarrowArgs& a, int64_t r0, int64_t r_end, int g, int t, double (&f)[MT + NZ][4]) {
#pragma unroll
    for (int cg = 0; cg < MT + NZ; ++cg) {
        const bool isz = cg >= MT;
        const int col = (isz ? (cg - MT) : cg) * 8 + g;
        const bool colok = isz ? (col < a.T) : (col < a.K);
        const double* base = isz ? a.Z + int64_t(col) * a.ldz : a.F + int64_t(col) * a.ldf;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t r = r0 + 8 * h + 2 * t;
            double2 v = make_double2(0.0, 0.0);
            if (colok && r + 1 < r_end) v = __ldg(reinterpret_cast<const double2*>(base + r));
            else if (colok && r < r_end) v.x = __ldg(base + r);
            f[cg][2 * h] = v.x;
            f[cg][2 * h + 1] = v.y;
        }
    }
}

template <int MT, int NZ, bool WEIGHTED>
__global__ void __launch_bounds__(NTHREADS_N, 1) gram_narrow_kernel(const NarrowArgs a) {
    using TI = NarrowTiles<MT, NZ>;
    extern __shared__ double red[];  // TI::NT * 64 doubles
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int64_t gw = int64_t(blockIdx.x) * NWARPS_N + warp;
    const int64_t r_begin = min(a.n, gw * a.rows_per_warp);
    const int64_t r_end = min(a.n, r_begin + a.rows_per_warp);

    double acc[TI::NT][2];
#pragma unroll
    for (int i = 0; i < TI::NT; ++i) acc[i][0] = acc[i][1] = 0.0;

    double cur[MT + NZ][4], nxt[MT + NZ][4];
    double wc[4] = {1.0, 1.0, 1.0, 1.0}, wn[4] = {1.0, 1.0, 1.0, 1.0};
    auto load_w = [&](int64_t r0, double (&wv)[4]) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int64_t r = r0 + 8 * h + 2 * t + j;
                wv[2 * h + j] = (r < r_end) ? __ldg(a.w + r) : 0.0;
            }
    };
    if (r_begin < r_end) {
        load_group<MT, NZ>(a, r_begin, r_end, g, t, cur);
        if (WEIGHTED) load_w(r_begin, wc);
    }
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 16) {
        if (r0 + 16 < r_end) {
            load_group<MT, NZ>(a, r0 + 16, r_end, g, t, nxt);
            if (WEIGHTED) load_w(r0 + 16, wn);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // k4-step q uses element q of every lane's group (a fixed row permutation)
            double af[MT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) af[mt] = WEIGHTED ? cur[mt][q] * wc[q] : cur[mt][q];
            int ti = 0;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = mt; nt < MT; ++nt, ++ti) dmma_n(acc[ti][0], acc[ti][1], af[mt], cur[nt][q]);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nz = 0; nz < NZ; ++nz, ++ti) dmma_n(acc[ti][0], acc[ti][1], af[mt], cur[MT + nz][q]);
        }
#pragma unroll
        for (int cg = 0; cg < MT + NZ; ++cg)
#pragma unroll
            for (int e = 0; e < 4; ++e) cur[cg][e] = nxt[cg][e];
        if (WEIGHTED) {
#pragma unroll
            for (int e = 0; e < 4; ++e) wc[e] = wn[e];
        }
    }
    // warps -> CTA partial, in warp order
    for (int wv = 0; wv < NWARPS_N; ++wv) {
        if (warp == wv) {
#pragma unroll
            for (int i = 0; i < TI::NT; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double* p = red + i * 64 + g * 8 + 2 * t + j;  // tile-local (row g, col 2t+j)
                    *p = (wv == 0) ? acc[i][j] : *p + acc[i][j];
                }
        }
        __syncthreads();
    }
    for (int e = tid; e < TI::NT * 64; e += NTHREADS_N) a.partial[size_t(blockIdx.x) * (TI::NT * 64) + e] = red[e];
}

// sum CTA partials in CTA order and scatter into packed = [G | C | zz]
template <int MT, int NZ>
__global__ void gram_narrow_reduce_kernel(const double* __restrict__ partial, int nctas, int K, int T,
                                          const double* __restrict__ zzpart, int nzz, double* __restrict__ packed) {
    using TI = NarrowTiles<MT, NZ>;
    double* G = packed;
    double* C = packed + size_t(K) * K;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < TI::NT * 64; e += gridDim.x * blockDim.x) {
        const int ti = e / 64, row = (e % 64) / 8, col = e % 8;
        double s = 0.0;
        for (int c = 0; c < nctas; ++c) s += partial[size_t(c) * (TI::NT * 64) + e];
        // decode the tile
        int mt = 0, nt = 0;
        bool isz = false;
        if (ti < TI::NFF) {
            int p = ti;
            for (mt = 0; mt < MT; ++mt) {
                const int cnt = MT - mt;
                if (p < cnt) { nt = mt + p; break; }
                p -= cnt;
            }
        } else {
            const int p = ti - TI::NFF;
            mt = p / (NZ > 0 ? NZ : 1);
            nt = p % (NZ > 0 ? NZ : 1);
            isz = true;
        }
        const int ga = mt * 8 + row, gb = nt * 8 + col;
        if (ga >= K) continue;
        if (isz) {
            if (gb < T) C[ga + size_t(gb) * K] = s;
        } else if (gb < K && ga <= gb) {
            G[ga + size_t(gb) * K] = s;
            G[gb + size_t(ga) * K] = s;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nzz; ++i) s += zzpart[i];
        packed[size_t(K) * K + size_t(K) * T] = s;
    }
}

template <int MT, int NZ>
void launch_narrow(const NarrowArgs& a0, int sms, GramWorkspace& ws, int K, int T, int nzz, double* dpacked, cudaStream_t st) {
    using TI = NarrowTiles<MT, NZ>;
    NarrowArgs a = a0;
    const size_t smem = size_t(TI::NT) * 64 * sizeof(double);
    int per_sm = 1;  // small tile sets leave registers for several CTAs per SM: more loads in flight
    if (a.w) FRK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gram_narrow_kernel<MT, NZ, true>, NTHREADS_N, smem));
    else FRK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gram_narrow_kernel<MT, NZ, false>, NTHREADS_N, smem));
    const int grid = sms * std::max(1, std::min(per_sm, 4));
    const int64_t nwarps = int64_t(grid) * NWARPS_N;
    a.rows_per_warp = std::max<int64_t>(16, ((a.n + nwarps - 1) / nwarps + 15) / 16 * 16);
    const size_t need = size_t(grid) * TI::NT * 64;
    ws.guard.grow(ws.partial, need);
    a.partial = ws.partial.p;
    if (a.w) gram_narrow_kernel<MT, NZ, true><<<grid, NTHREADS_N, smem, st>>>(a);
    else gram_narrow_kernel<MT, NZ, false><<<grid, NTHREADS_N, smem, st>>>(a);
    FRK_CUDA(cudaGetLastError());
    gram_narrow_reduce_kernel<MT, NZ><<<8, 256, 0, st>>>(ws.partial.p, grid, K, T, ws.zzpart.p, nzz, dpacked);
    FRK_CUDA(cudaGetLastError());
}

template <int MT>
bool dispatch_nz(int nz, const NarrowArgs& a, int sms, GramWorkspace& ws, int K, int T, int nzz, double* dpacked, cudaStream_t st) {
    switch (nz) {
        case 0: launch_narrow<MT, 0>(a, sms, ws, K, T, nzz, dpacked, st); return true;
        case 1: launch_narrow<MT, 1>(a, sms, ws, K, T, nzz, dpacked, st); return true;
        case 2: if constexpr (MT <= 4) { launch_narrow<MT, 2>(a, sms, ws, K, T, nzz, dpacked, st); return true; } return false;
        default: return false;
    }
}

}  // namespace

// true if the narrow kernel took the job (zz partials must already be in ws.zzpart)
bool gram_stats_narrow_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                           const double* dw, double* dpacked, GramWorkspace& ws, int sms, int nzz, cudaStream_t st) {
    if (K > 48 || T > 16) return false;  // register budget: MT(MT+1)/2 + MT*NZ accumulator tiles + two fragment groups
    if (reinterpret_cast<uintptr_t>(dF) % 16 || ldf % 2) return false;
    if (T > 0 && (reinterpret_cast<uintptr_t>(dZ) % 16 || ldz % 2)) return false;
    NarrowArgs a;
    a.F = dF; a.ldf = ldf; a.K = K; a.Z = dZ; a.ldz = ldz; a.T = T; a.w = dw; a.n = n;
    a.rows_per_warp = 0; a.partial = nullptr;
    const int mt = (K + 7) / 8, nz = (T + 7) / 8;
    switch (mt) {
        case 1: return dispatch_nz<1>(nz, a, sms, ws, K, T, nzz, dpacked, st);
        case 2: return dispatch_nz<2>(nz, a, sms, ws, K, T, nzz, dpacked, st);
        case 3: return dispatch_nz<3>(nz, a, sms, ws, K, T, nzz, dpacked, st);
        case 4: return dispatch_nz<4>(nz, a, sms, ws, K, T, nzz, dpacked, st);
        case 5: return dispatch_nz<5>(nz, a, sms, ws, K, T, nzz, dpacked, st);
        case 6: return dispatch_nz<6>(nz, a, sms, ws, K, T, nzz, dpacked, st);
    }
    return false;
}

}  // namespace frk
