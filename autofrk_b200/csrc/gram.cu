// K4 -- the sufficient statistics of the fixed-rank-kriging fit in one pass over the rows:
//   G = F'F      (R/helper.R:20  `t(Fk) %*% iDFk`, identity D)              K x K
//   C = F'Z      (what R/autoFRK.R:433-435 needs: ihF %*% Data = half %*% (F'Z))   K x T
//   zz = sum(Z^2) (R/autoFRK.R:431)
// With optional row weights w (diagonal D^-1: cMLEsp, R/autoFRK.R:528-538; NA masks: R/autoFRK.R:609-615) the
// same pass gives G = F' diag(w) F, C = F' diag(w) Z, zz = sum(w Z^2).
// F is n x K, Z is n x T, both column-major (R layout), so the long dimension n is the MMA k-dimension
// and both operands of every product are column panels of the same matrices.
//
//   * FP64 tensor cores (DMMA.8x8x4), 128 x 128 output tiles, 16 warps x (32 x 32) register tiles.
//   * only tile pairs (a <= b) of G are computed (SYRK), plus the F x Z pairs.
//   * rows are split across CTAs; every CTA writes its partial tile to a workspace and a second kernel
//     adds the partials in a fixed order -> bitwise reproducible, no atomics.
//   * panels are staged with cp.async (LDGSTS), 4 stages deep; shared layout [column][16 rows + 4 pad]
//     makes the fragment loads (column g, row t) bank-conflict free and keeps each column's 128 B
//     contiguous for the copy.
//   * CTAs of the same row split are adjacent in launch order, so a row block of F is pulled from
//     HBM once and served to the other tile pairs from the 126 MB L2.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>
#include "frk_common.cuh"
#include "gram.cuh"

namespace frk {

namespace {

constexpr int TB = 128;       // output tile edge
constexpr int KR = 16;        // rows per pipeline stage
constexpr int SP = KR + 4;    // padded column stride in shared memory (doubles); 20 mod 16 == 4
constexpr int NST = 4;        // cp.async stages
constexpr int NTHREADS = 512;
constexpr int PANEL_DOUBLES = TB * SP;
constexpr int STAGE_DOUBLES = 2 * PANEL_DOUBLES + KR;  // A panel, B panel, row weights of the stage
constexpr size_t GRAM_SMEM = size_t(NST) * STAGE_DOUBLES * sizeof(double);

struct GramArgs {
    const double* F;
    int64_t ldf;
    int K;
    const double* Z;
    int64_t ldz;
    int T;
    int64_t n;
    const double* w;  // optional row weights (n): G = F' diag(w) F, C = F' diag(w) Z; nullptr = identity
    int ntf, ntz, npairs, nsplit;
    int64_t rows_per_split;
    double* partial;  // [nsplit][npairs][TB*TB], column-major tiles
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <bool ALIGNED, bool WEIGHTED>
__global__ void __launch_bounds__(NTHREADS, 1) gram_kernel(const GramArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const smem = reinterpret_cast<double*>(smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp >> 2, wc = warp & 3;

    const int pair = blockIdx.x % a.npairs;
    const int split = blockIdx.x / a.npairs;
    // decode the tile pair
    int ta = 0, tb = 0;
    bool bz = false;
    {
        const int nff = a.ntf * (a.ntf + 1) / 2;
        if (pair < nff) {
            int p = pair;
            for (ta = 0; ta < a.ntf; ++ta) {
                const int cnt = a.ntf - ta;
                if (p < cnt) { tb = ta + p; break; }
                p -= cnt;
            }
        } else {
            const int p = pair - nff;
            ta = p / a.ntz;
            tb = p % a.ntz;
            bz = true;
        }
    }
    const bool diag = (!bz && ta == tb);
    const double* Abase = a.F + int64_t(ta) * TB * a.ldf;
    const int Acols = min(TB, a.K - ta * TB);
    const int64_t Ald = a.ldf;
    const double* Bbase = bz ? a.Z + int64_t(tb) * TB * a.ldz : a.F + int64_t(tb) * TB * a.ldf;
    const int Bcols = bz ? min(TB, a.T - tb * TB) : min(TB, a.K - tb * TB);
    const int64_t Bld = bz ? a.ldz : a.ldf;

    const int64_t r_begin = int64_t(split) * a.rows_per_split;
    const int64_t r_end = min(a.n, r_begin + a.rows_per_split);
    const int nsteps = (r_end > r_begin) ? int((r_end - r_begin + KR - 1) / KR) : 0;

    auto load_panel = [&](double* dst, const double* base, int64_t ld, int ncols, int64_t r0) {
        const int rvalid = int(min(int64_t(KR), r_end - r0));
        if (ALIGNED) {
            // 16-byte pieces: 8 per column
#pragma unroll
            for (int i = 0; i < (TB * (KR / 2)) / NTHREADS; ++i) {
                const int pid = tid + i * NTHREADS;
                const int col = pid >> 3, p = pid & 7;
                int bytes = (col < ncols) ? max(0, min(2, rvalid - 2 * p)) * 8 : 0;
                const double* src = (bytes > 0) ? base + int64_t(col) * ld + r0 + 2 * p : base;
                cp_async16(dst + col * SP + 2 * p, src, bytes);
            }
        } else {
#pragma unroll
            for (int i = 0; i < (TB * KR) / NTHREADS; ++i) {
                const int pid = tid + i * NTHREADS;
                const int col = pid >> 4, p = pid & 15;
                const int bytes = (col < ncols && p < rvalid) ? 8 : 0;
                const double* src = bytes ? base + int64_t(col) * ld + r0 + p : base;
                cp_async8(dst + col * SP + p, src, bytes);
            }
        }
    };
    auto load_stage = [&](int stage, int step) {
        double* As = smem + size_t(stage) * STAGE_DOUBLES;
        const int64_t r0 = r_begin + int64_t(step) * KR;
        load_panel(As, Abase, Ald, Acols, r0);
        if (!diag) load_panel(As + PANEL_DOUBLES, Bbase, Bld, Bcols, r0);
        if (WEIGHTED && tid < KR) {  // weights of the stage's rows (0 past the split's end)
            const int bytes = (r0 + tid < r_end) ? 8 : 0;
            cp_async8(As + 2 * PANEL_DOUBLES + tid, bytes ? a.w + r0 + tid : a.w, bytes);
        }
    };

    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nsteps) load_stage(s, s);
        cp_async_commit();
    }
    for (int step = 0; step < nsteps; ++step) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        if (step + NST - 1 < nsteps) load_stage((step + NST - 1) % NST, step + NST - 1);
        cp_async_commit();
        const double* As = smem + size_t(step % NST) * STAGE_DOUBLES;
        const double* Bs = diag ? As : As + PANEL_DOUBLES;
        const double* wsm = As + 2 * PANEL_DOUBLES + t;
        const double* ap = As + (wr * 32 + g) * SP + t;
        const double* bp = Bs + (wc * 32 + g) * SP + t;
#pragma unroll
        for (int q = 0; q < KR / 4; ++q) {
            double af[4], bf[4];
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) af[mt] = ap[mt * 8 * SP + 4 * q];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) bf[nt] = bp[nt * 8 * SP + 4 * q];
            if (WEIGHTED) {  // A operand = F' diag(w): row 4q+t of the stage
                const double wq = wsm[4 * q];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) af[mt] *= wq;
            }
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    cp_async_wait<0>();

    double* out = a.partial + (size_t(split) * a.npairs + pair) * (TB * TB);
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int row = wr * 32 + mt * 8 + g;
                const int col = wc * 32 + nt * 8 + 2 * t + j;
                out[col * TB + row] = acc[mt][nt][j];
            }
}

// sum of squares of Z (n x T, ld ldz): fixed grid (gx blocks per column group, grid.y column groups), per-block
// partials summed later in a fixed order.  Rows are the fast index: coalesced loads, no integer division per element.
__global__ void sumsq_kernel(const double* __restrict__ Z, int64_t ldz, int64_t n, int T, const double* __restrict__ w,
                             double* __restrict__ part) {
    double s = 0.0;
    for (int c = blockIdx.y; c < T; c += gridDim.y) {
        const double* col = Z + int64_t(c) * ldz;
        for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) {
            const double z = col[i];
            s = fma(w ? z * w[i] : z, z, s);
        }
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = sh[0];
}

// add the row-split partials in split order and scatter into packed = [G (K x K) | C (K x T) | zz]
__global__ void gram_reduce_kernel(const double* __restrict__ partial, int nsplit, int npairs, int ntf, int ntz, int K,
                                   int T, const double* __restrict__ zzpart, int nzz, double* __restrict__ packed) {
    const int pair = blockIdx.x;
    int ta = 0, tb = 0;
    bool bz = false;
    const int nff = ntf * (ntf + 1) / 2;
    if (pair < nff) {
        int p = pair;
        for (ta = 0; ta < ntf; ++ta) {
            const int cnt = ntf - ta;
            if (p < cnt) { tb = ta + p; break; }
            p -= cnt;
        }
    } else {
        const int p = pair - nff;
        ta = p / ntz;
        tb = p % ntz;
        bz = true;
    }
    double* G = packed;
    double* C = packed + size_t(K) * K;
    for (int e = threadIdx.x; e < TB * TB; e += blockDim.x) {
        const int row = e % TB, col = e / TB;
        const int ga = ta * TB + row, gb = tb * TB + col;
        if (ga >= K) continue;
        if (bz ? (gb >= T) : (gb >= K)) continue;
        if (!bz && ta == tb && row > col) continue;  // lower triangle of a diagonal tile is mirrored
        double s = 0.0;
        for (int sp = 0; sp < nsplit; ++sp) s += partial[(size_t(sp) * npairs + pair) * (TB * TB) + e];
        if (bz) {
            C[ga + size_t(gb) * K] = s;
        } else {
            G[ga + size_t(gb) * K] = s;
            G[gb + size_t(ga) * K] = s;
        }
    }
    if (pair == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nzz; ++i) s += zzpart[i];
        packed[size_t(K) * K + size_t(K) * T] = s;
    }
}

}  // namespace

size_t gram_packed_len(int K, int T) { return size_t(K) * K + size_t(K) * T + 1; }

void gram_stats_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                    const double* dw, double* dpacked, GramWorkspace& ws, cudaStream_t st) {
    if (K < 1) fail(1, "F must have at least one column");
    if (n < 1) fail(1, "F must have at least one row");
    if (T < 0) fail(1, "T must be non-negative");
    if (ldf < n || (T > 0 && ldz < n)) fail(1, "leading dimensions must be >= n");
    int dev = 0, sms = 0;
    FRK_CUDA(cudaGetDevice(&dev));
    FRK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const StreamGuardScope ordered(ws.guard, st);

    const char* const path_c = std::getenv("FRK_GRAM_PATH");   // tuning / test aid: mid | narrow | tma | cpasync (default: best fit)
    const std::string path_env(path_c ? path_c : "");
    const bool force_cpasync = path_env == "cpasync";
    const bool force_wide = force_cpasync || path_env == "tma";
    // K <= 16: the whole tile set is a handful of accumulators and the pass is pure streaming -- the direct-from-global
    // register kernel (gram_narrow.cu) wins (0.23 vs 0.29 ms at n = 5e6, K = 16, T = 1); from K = 32 on the
    // register-resident kernel is as fast or faster and needs no separate pass over Z
    const bool prefer_narrow = path_env == "narrow" || (path_env.empty() && K <= 16 && T <= 16);
    if ((path_env.empty() || path_env == "mid") && !prefer_narrow &&
        gram_stats_mid_dev(dF, n, K, ldf, dZ, T, ldz, dw, dpacked, ws, sms, st)) {
        ws.launches += 2;
        ws.last_path = "mid";
        return;
    }

    const int nzz = 512;
    ws.guard.grow(ws.zzpart, size_t(nzz));
    if (T > 0) {
        const int gy = std::min(T, 32), gx = nzz / gy;   // gx * gy <= nzz partials; the rest stay zero
        if (gx * gy < nzz) FRK_CUDA(cudaMemsetAsync(ws.zzpart.p + gx * gy, 0, sizeof(double) * (nzz - gx * gy), st));
        sumsq_kernel<<<dim3(unsigned(gx), unsigned(gy)), 256, 0, st>>>(dZ, ldz, n, T, dw, ws.zzpart.p);
    } else {
        FRK_CUDA(cudaMemsetAsync(ws.zzpart.p, 0, sizeof(double) * nzz, st));
    }
    FRK_CUDA(cudaGetLastError());

    if (!force_wide && path_env != "mid" && gram_stats_narrow_dev(dF, n, K, ldf, dZ, T, ldz, dw, dpacked, ws, sms, nzz, st)) {
        ws.launches += 3;
        ws.last_path = "narrow";
        return;
    }
    if (!force_cpasync && gram_tma_usable(dF, ldf, dZ, T, ldz, n)) {
        gram_stats_tma_dev(dF, n, K, ldf, dZ, T, ldz, dw, dpacked, ws, sms, nzz, st);
        ws.launches += 3;
        ws.last_path = "tma";
        return;
    }
    {
        GramArgs a;
        a.F = dF; a.ldf = ldf; a.K = K;
        a.Z = dZ; a.ldz = ldz; a.T = T;
        a.n = n;
        a.w = dw;
        a.ntf = (K + TB - 1) / TB;
        a.ntz = (T + TB - 1) / TB;
        a.npairs = a.ntf * (a.ntf + 1) / 2 + a.ntf * a.ntz;
        // row splits: fill whole waves of one CTA per SM, at least 256 rows per split
        const int max_split = int(std::max<int64_t>(1, n / 256));
        int best_ns = 1;
        double best_eff = 0.0;
        for (int w = 1; w <= 4; ++w) {
            int ns = std::max(1, std::min(max_split, (sms * w) / a.npairs));
            const int ctas = ns * a.npairs;
            const double eff = double(ctas) / (double((ctas + sms - 1) / sms) * sms);
            if (eff > best_eff + 1e-9) { best_eff = eff; best_ns = ns; }
        }
        a.nsplit = best_ns;
        a.rows_per_split = ((n + a.nsplit - 1) / a.nsplit + KR - 1) / KR * KR;
        const size_t need = size_t(a.nsplit) * a.npairs * TB * TB;
        ws.guard.grow(ws.partial, need);
        a.partial = ws.partial.p;
        const bool aligned = (reinterpret_cast<uintptr_t>(dF) % 16 == 0) && (ldf % 2 == 0) &&
                             (T == 0 || ((reinterpret_cast<uintptr_t>(dZ) % 16 == 0) && (ldz % 2 == 0)));
        static std::once_flag attr_once[64];  // the opt-in to large dynamic shared memory is per device; thread-safe
        std::call_once(attr_once[dev & 63], [] {
            FRK_CUDA(cudaFuncSetAttribute(gram_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(GRAM_SMEM)));
            FRK_CUDA(cudaFuncSetAttribute(gram_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(GRAM_SMEM)));
            FRK_CUDA(cudaFuncSetAttribute(gram_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(GRAM_SMEM)));
            FRK_CUDA(cudaFuncSetAttribute(gram_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(GRAM_SMEM)));
        });
        const int grid = a.nsplit * a.npairs;
        if (aligned && !dw) gram_kernel<true, false><<<grid, NTHREADS, GRAM_SMEM, st>>>(a);
        else if (!aligned && !dw) gram_kernel<false, false><<<grid, NTHREADS, GRAM_SMEM, st>>>(a);
        else if (aligned) gram_kernel<true, true><<<grid, NTHREADS, GRAM_SMEM, st>>>(a);
        else gram_kernel<false, true><<<grid, NTHREADS, GRAM_SMEM, st>>>(a);
        FRK_CUDA(cudaGetLastError());
        gram_reduce_kernel<<<a.npairs, 256, 0, st>>>(ws.partial.p, a.nsplit, a.npairs, a.ntf, a.ntz, K, T, ws.zzpart.p, nzz,
                                                     dpacked);
        FRK_CUDA(cudaGetLastError());
        ws.last_nsplit = a.nsplit;
        ws.last_npairs = a.npairs;
    }
    ws.last_path = "cpasync";
    ws.launches += 3;
}

}  // namespace frk
