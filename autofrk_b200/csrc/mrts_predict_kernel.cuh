// K1 -- fused thin-plate-spline kernel evaluation + projection onto the knot eigenvectors.
//
// Replaces, in one pass and without ever writing the n2 x m kernel matrix to HBM:
//   tpm_predict                     src/autoFRK.cpp:109-153   (Phi_new[i,j] = phi_d(|x_i - u_j|))
//   X1 = Hnew * UZ.block(0,0,n,k)   src/autoFRK.cpp:316
//   X1 - [1,xnew] * (BBBH * UZ.block(0,0,n,k))                src/autoFRK.cpp:317-324
//
// Shape of the computation: out(BM x BN tile) = -[1,x]*Cpoly + sum over knot chunks Phi(BM x KC) * W(KC x BN)
//   * FP64 tensor cores (mma.sync m8n8k4.f64 -> SASS DMMA.8x8x4), accumulators in registers
//     (tcgen05 / TMEM have no f64 kind).
//   * W is pre-packed once per plan into the exact shared-memory image of a chunk
//     ([k4-group][column][4 knots], followed by the chunk's knot coordinates) so that
//       - one TMA bulk copy (cp.async.bulk -> UBLKCP) per chunk stages W + knots, tracked by an
//         mbarrier, NSTAGE deep,
//       - every B-fragment load is a contiguous, conflict-free 256 B LDS.64 per warp.
//   * The FP64 ALU and the DMMA unit share one pipe on B200 (measured: profiles/fp64_peaks_r01.json),
//     so kernel evaluation is pure overhead on the binding resource.  Every thread therefore
//     evaluates the same number of Phi values for chunk c+1 INSIDE the DMMA block of chunk c (one
//     basic block, branch-free table log for d = 2), letting the scheduler drop the Phi arithmetic
//     into the issue slots between dependent-free DMMAs instead of serialising it in front of them.
//   * persistent CTAs (one per SM), tiles of BM rows dealt round-robin.
#pragma once
#include "frk_common.cuh"
#include "tps_log_table.h"

namespace frk {

template <int D_, int WR_, int WC_, int MT_, int NT_, int KC_, int NSTAGE_>
struct PredictCfg {
    static constexpr int D = D_, WR = WR_, WC = WC_, MT = MT_, NT = NT_, KC = KC_, NSTAGE = NSTAGE_;
    static constexpr int DP = (D == 3) ? 4 : D;   // doubles per stored coordinate tuple
    static constexpr int NWARPS = WR * WC;
    static constexpr int NTHREADS = NWARPS * 32;
    static constexpr int BM = WR * MT * 8;        // rows (locations) per tile
    static constexpr int BN = WC * NT * 8;        // columns per slice (padded)
    static constexpr int KQ = KC / 4;             // k4 groups per chunk
    static constexpr int VPT = (BM * KC + NTHREADS - 1) / NTHREADS;  // Phi values per thread per chunk
    static constexpr int STAGE_DOUBLES = KC * BN + KC * DP;
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int PHI_DOUBLES = BM * KC;
    static constexpr int XS_DOUBLES = BM * DP;
    static constexpr int LOGTAB_DOUBLES = (D == 2) ? 256 : 0;
    static constexpr size_t SMEM_BYTES = size_t(NSTAGE) * STAGE_BYTES + 2 * PHI_DOUBLES * 8 + 2 * XS_DOUBLES * 8 +
                                         LOGTAB_DOUBLES * 8 + NSTAGE * 8;
    static_assert(KC % 4 == 0, "KC must be a multiple of the DMMA k (4)");
    static_assert(STAGE_BYTES % 16 == 0, "TMA bulk copies need 16 B multiples");
    static_assert(NTHREADS <= 1024, "too many warps");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct PredictArgs {
    const double* xnew;   // n2 x d, column-major, leading dimension ldx
    int64_t ldx;
    int64_t n2;
    const double* wpack;  // nchunks * STAGE_DOUBLES (packed W + knot coordinates)
    int nchunks;
    const double* cpoly;  // (d+1) x BN, row-major, zero padded:  BBBH * W
    double* out;          // n2 x ncols, column-major, leading dimension ldo
    int64_t ldo;
    int ncols;            // valid columns of this slice (<= BN)
    int64_t ntiles;
};

// phi_d(r): src/autoFRK.cpp:127-128 (d=1), :135-138 (d=2), :145-148 (d=3); libm-style log (used by the fit, K2)
template <int D>
__device__ __forceinline__ double tps_phi(const double* __restrict__ x, const double* __restrict__ u) {
    if constexpr (D == 1) {
        const double r = fabs(x[0] - u[0]);
        return (r * r * r) * (1.0 / 12.0);
    } else if constexpr (D == 2) {
        const double2 xv = *reinterpret_cast<const double2*>(x);
        const double2 uv = *reinterpret_cast<const double2*>(u);
        const double dx = xv.x - uv.x, dy = xv.y - uv.y;
        const double r2 = fma(dx, dx, dy * dy);
        // r^2 log(r) / (8 pi) = r^2 log(r^2) / (16 pi); stays 0 at r == 0 like the reference (:137)
        const double v = r2 * log(r2) * 0.019894367886486918;  // 1/(16*pi)
        return (r2 > 0.0) ? v : 0.0;
    } else {
        const double2 xa = *reinterpret_cast<const double2*>(x);
        const double2 ua = *reinterpret_cast<const double2*>(u);
        const double dx = xa.x - ua.x, dy = xa.y - ua.y, dz = x[2] - u[2];
        const double r = sqrt(fma(dx, dx, fma(dy, dy, dz * dz)));
        return -0.125 * r;
    }
}

// Branch-free FP64 natural log of a positive normal double (Tang-style, 128-entry table in shared
// memory): x = 2^e m, m in [1,2); j = top 7 mantissa bits, r = m*rc_j - 1 (one FMA, |r| <= 2^-8),
// log x = e ln2 + lc_j + log1p(r), log1p by a degree-7 Taylor polynomial (|r|^8/8 < 2^-67).
// Absolute error ~1.2e-16 + 1.1e-16 |log x| (tools/logtest notes in DESIGN.md): on par with libm for
// r^2 log r^2.  x == 0, subnormals, inf and nan give finite garbage / nan that the caller masks (x > 0)
// or that is multiplied by x itself.
__device__ __forceinline__ double log_pos_table(double x, const double2* __restrict__ tab) {
    const long long bits = __double_as_longlong(x);
    const int hi = int(bits >> 32);
    const double2 tc = tab[(hi >> 13) & 0x7F];
    const double ef = double((hi >> 20) - 1023);
    const double m = __longlong_as_double((bits & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
    const double r = fma(m, tc.x, -1.0);
    double p = 1.0 / 7.0;
    p = fma(p, r, -1.0 / 6.0);
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    const double q = fma(r * r, p, ef * 1.90821492927058770002e-10);   // + e * ln2_lo
    return fma(ef, 6.93147180369123816490e-01, tc.y + (r + q));        // e * ln2_hi + (lc + log1p(r))
}

// kernel value from pre-loaded coordinates (registers): same formulas as tps_phi, table log for d = 2
template <int D>
__device__ __forceinline__ double tps_phi_fast(const double (&x)[3], const double (&u)[3], const double2* __restrict__ tab) {
    if constexpr (D == 1) {
        const double r = fabs(x[0] - u[0]);
        return (r * r * r) * (1.0 / 12.0);
    } else if constexpr (D == 2) {
        const double dx = x[0] - u[0], dy = x[1] - u[1];
        const double r2 = fma(dx, dx, dy * dy);
        const double v = (r2 * 0.019894367886486918) * log_pos_table(r2, tab);
        return (r2 > 0.0) ? v : 0.0;
    } else {
        const double dx = x[0] - u[0], dy = x[1] - u[1], dz = x[2] - u[2];
        return -0.125 * sqrt(fma(dx, dx, fma(dy, dy, dz * dz)));
    }
}

// DMMA without `volatile`: it is a pure function of its operands, so the scheduler may interleave it
// with the Phi arithmetic of the next chunk.
__device__ __forceinline__ void dmma884_nv(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1) mrts_predict_kernel(const PredictArgs a) {
    constexpr int D = C::D, DP = C::DP, BM = C::BM, BN = C::BN, KC = C::KC, KQ = C::KQ;
    constexpr int MT = C::MT, NT = C::NT, NSTAGE = C::NSTAGE, NTHREADS = C::NTHREADS, VPT = C::VPT;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const stage0 = reinterpret_cast<double*>(smem_raw);
    double* const phib = stage0 + size_t(NSTAGE) * C::STAGE_DOUBLES;
    double* const xs = phib + 2 * C::PHI_DOUBLES;
    double* const logtab_d = xs + 2 * C::XS_DOUBLES;
    uint64_t* const full = reinterpret_cast<uint64_t*>(logtab_d + C::LOGTAB_DOUBLES);
    const double2* const logtab = reinterpret_cast<const double2*>(logtab_d);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp / C::WC, wc = warp % C::WC;

    if (int64_t(blockIdx.x) >= a.ntiles) return;
    const int64_t ntl = (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;  // tiles of this CTA
    const int nchunks = a.nchunks;
    const int64_t total = ntl * nchunks;  // (tile, chunk) iterations of this CTA

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
        for (int s = 0; s < NSTAGE; ++s)
            if (s < total) {
                mbar_expect_tx(&full[s], C::STAGE_BYTES);
                tma_bulk_g2s(stage0 + size_t(s) * C::STAGE_DOUBLES,
                             a.wpack + size_t(s % nchunks) * C::STAGE_DOUBLES, C::STAGE_BYTES, &full[s]);
            }
    }
    if constexpr (D == 2) {
        for (int i = tid; i < 128; i += NTHREADS) {
            logtab_d[2 * i] = kLogTableDev[i].rc;
            logtab_d[2 * i + 1] = kLogTableDev[i].lc;
        }
    }

    // coordinates of one row tile -> shared (zero for rows past n2; their results are never stored)
    auto load_xs = [&](int buf, int64_t tile) {
        double* dst = xs + buf * C::XS_DOUBLES;
        for (int idx = tid; idx < BM * D; idx += NTHREADS) {
            const int r = idx % BM, c = idx / BM;
            const int64_t i = tile * BM + r;
            dst[r * DP + c] = (i < a.n2) ? __ldg(a.xnew + i + c * a.ldx) : 0.0;
        }
    };

    // the Phi values this thread owns in every chunk: slot v = tid + i*NTHREADS of the A-fragment image
    // [k4 group][row][4 knots]  ->  (row, knot-in-chunk)
    int phi_row[VPT], phi_kk[VPT];
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
        const int v = tid + i * NTHREADS;
        const int vv = (v < BM * KC) ? v : 0;
        phi_row[i] = (vv % (BM * 4)) >> 2;
        phi_kk[i] = (vv / (BM * 4)) * 4 + (vv & 3);
    }
    double px[VPT][3], pu[VPT][3];  // coordinates of the pending Phi evaluations
    auto phi_load = [&](int64_t it, int xbuf) {
        const int s = int(it % NSTAGE);
        const double* knots = stage0 + size_t(s) * C::STAGE_DOUBLES + KC * BN;
        const double* xsb = xs + xbuf * C::XS_DOUBLES;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
#pragma unroll
            for (int c = 0; c < D; ++c) {
                px[i][c] = xsb[phi_row[i] * DP + c];
                pu[i][c] = knots[phi_kk[i] * DP + c];
            }
        }
    };
    auto phi_store = [&](int64_t it, const double (&val)[VPT]) {
        double* dst = phib + int(it & 1) * C::PHI_DOUBLES;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = tid + i * NTHREADS;
            if (VPT * NTHREADS == BM * KC || v < BM * KC) dst[v] = val[i];
        }
    };

    load_xs(0, blockIdx.x);
    __syncthreads();  // barrier inits, log table and xs[0] visible
    {
        mbar_wait(&full[0], 0u);
        phi_load(0, 0);
        double val[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) val[i] = tps_phi_fast<D>(px[i], pu[i], logtab);
        phi_store(0, val);
    }
    __syncthreads();

    int64_t it = 0;
    for (int64_t til = 0; til < ntl; ++til) {
        const int64_t tile = blockIdx.x + til * gridDim.x;
        const int xb = int(til & 1);
        if (til + 1 < ntl) {
            load_xs(xb ^ 1, tile + gridDim.x);
            if (nchunks == 1) __syncthreads();
        }
        // accumulators start at -[1, x] * Cpoly   (src/autoFRK.cpp:317-324)
        double acc[MT][NT][2];
        {
            const double* xsb = xs + xb * C::XS_DOUBLES;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                const int row = (wr * MT + mt) * 8 + g;
                double xr[D];
#pragma unroll
                for (int c = 0; c < D; ++c) xr[c] = xsb[row * DP + c];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = (wc * NT + nt) * 8 + 2 * t + j;
                        double p = __ldg(a.cpoly + col);
#pragma unroll
                        for (int c = 0; c < D; ++c) p = fma(xr[c], __ldg(a.cpoly + (c + 1) * BN + col), p);
                        acc[mt][nt][j] = -p;
                    }
            }
        }

        for (int c = 0; c < nchunks; ++c, ++it) {
            // stage of the NEXT chunk: its knots feed the Phi evaluation that overlaps this chunk's DMMAs
            const bool more = (it + 1 < total);
            if (more) mbar_wait(&full[int((it + 1) % NSTAGE)], uint32_t((it + 1) / NSTAGE) & 1u);
            phi_load(more ? it + 1 : it, (c + 1 == nchunks) ? (xb ^ 1) : xb);

            const int s = int(it % NSTAGE);
            const double* ws = stage0 + size_t(s) * C::STAGE_DOUBLES + (wc * NT * 8) * 4 + lane;
            const double* ph = phib + int(it & 1) * C::PHI_DOUBLES + (wr * MT * 8) * 4 + lane;
            double val[VPT];
            // ---- one basic block: DMMAs of chunk `it` + Phi arithmetic of chunk `it + 1` ----
#pragma unroll
            for (int q = 0; q < KQ; ++q) {
                double af[MT], bf[NT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) af[mt] = ph[q * BM * 4 + mt * 32];
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) bf[nt] = ws[q * BN * 4 + nt * 32];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) dmma884_nv(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
                if (q == 0) {
#pragma unroll
                    for (int i = 0; i < VPT; ++i) val[i] = tps_phi_fast<D>(px[i], pu[i], logtab);
                }
            }
            phi_store(it + 1, val);
            __syncthreads();  // stage s and phi buffer (it&1) are free; phi(it+1) is visible
            if (tid == 0 && it + NSTAGE < total) {
                mbar_expect_tx(&full[s], C::STAGE_BYTES);
                tma_bulk_g2s(stage0 + size_t(s) * C::STAGE_DOUBLES,
                             a.wpack + size_t((it + NSTAGE) % nchunks) * C::STAGE_DOUBLES, C::STAGE_BYTES,
                             &full[s]);
            }
        }

        // store the tile, column-major (R layout)
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int64_t i = tile * BM + (wr * MT + mt) * 8 + g;
            if (i < a.n2) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int col = (wc * NT + nt) * 8 + 2 * t + j;
                        if (col < a.ncols) a.out[int64_t(col) * a.ldo + i] = acc[mt][nt][j];
                    }
            }
        }
    }
}

// host-side launcher table entry --------------------------------------------------------------
struct PredictKernelInfo {
    int d, bm, bn, kc, nstage, dp, nthreads, stage_doubles;
    size_t smem_bytes;
    void (*launch)(const PredictArgs&, int grid, cudaStream_t);
    void (*prepare)();  // opt in to large dynamic shared memory (once)
};

template <class C>
void predict_prepare() {
    FRK_CUDA(cudaFuncSetAttribute(mrts_predict_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(C::SMEM_BYTES)));
}
template <class C>
void predict_launch(const PredictArgs& a, int grid, cudaStream_t st) {
    mrts_predict_kernel<C><<<grid, C::NTHREADS, C::SMEM_BYTES, st>>>(a);
}
template <class C>
constexpr PredictKernelInfo predict_info() {
    return PredictKernelInfo{C::D, C::BM, C::BN, C::KC, C::NSTAGE, C::DP, C::NTHREADS, C::STAGE_DOUBLES,
                             C::SMEM_BYTES, &predict_launch<C>, &predict_prepare<C>};
}

// the instantiated configurations, per dimension (defined in mrts_predict_d{1,2,3}.cu)
const PredictKernelInfo* predict_kernels_d1(int* count);
const PredictKernelInfo* predict_kernels_d2(int* count);
const PredictKernelInfo* predict_kernels_d3(int* count);

// <D, WR, WC, MT, NT, KC, NSTAGE>: KC chosen so that every thread evaluates a whole number of Phi
// values per chunk (BM*KC is a multiple of the thread count)
#define FRK_PREDICT_CONFIG_TABLE(D)                                                                        \
    predict_info<PredictCfg<D, 4, 4, 4, 1, 8, 4>>(),   /* BM 128, BN  32, 2 Phi/thread */                  \
    predict_info<PredictCfg<D, 4, 4, 4, 2, 8, 4>>(),   /* BM 128, BN  64 */                                \
    predict_info<PredictCfg<D, 4, 4, 4, 3, 8, 4>>(),   /* BM 128, BN  96 */                                \
    predict_info<PredictCfg<D, 4, 4, 4, 4, 8, 4>>(),   /* BM 128, BN 128 */                                \
    predict_info<PredictCfg<D, 2, 8, 4, 3, 16, 4>>(),  /* BM  64, BN 192, 2 Phi/thread */                  \
    predict_info<PredictCfg<D, 3, 5, 4, 5, 20, 4>>(),  /* BM  96, BN 200, 4 Phi/thread, 15 warps */        \
    predict_info<PredictCfg<D, 2, 8, 4, 4, 16, 4>>(),  /* BM  64, BN 256 */                                \
    predict_info<PredictCfg<D, 2, 8, 4, 5, 16, 4>>(),  /* BM  64, BN 320 */                                \
    predict_info<PredictCfg<D, 1, 16, 4, 3, 16, 4>>(), /* BM  32, BN 384, 1 Phi/thread */                  \
    predict_info<PredictCfg<D, 1, 16, 4, 4, 16, 3>>()  /* BM  32, BN 512, 1 Phi/thread, 3 x 66 KB stages */

}  // namespace frk
