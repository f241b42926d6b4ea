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
//   * Phi for chunk c+1 is evaluated by all threads (one value per (row, knot)) straight into the
//     A-fragment image in shared memory while the DMMAs of chunk c issue; each Phi value is
//     computed exactly once per column slice.
//   * persistent CTAs (one per SM), tiles of BM rows dealt round-robin.
#pragma once
#include "frk_common.cuh"

namespace frk {

template <int D_, int WR_, int WC_, int MT_, int NT_, int KC_ = 8, int NSTAGE_ = 4>
struct PredictCfg {
    static constexpr int D = D_, WR = WR_, WC = WC_, MT = MT_, NT = NT_, KC = KC_, NSTAGE = NSTAGE_;
    static constexpr int DP = (D == 3) ? 4 : D;   // doubles per stored coordinate tuple
    static constexpr int NWARPS = WR * WC;
    static constexpr int NTHREADS = NWARPS * 32;
    static constexpr int BM = WR * MT * 8;        // rows (locations) per tile
    static constexpr int BN = WC * NT * 8;        // columns per slice (padded)
    static constexpr int KQ = KC / 4;             // k4 groups per chunk
    static constexpr int STAGE_DOUBLES = KC * BN + KC * DP;
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int PHI_DOUBLES = BM * KC;
    static constexpr int XS_DOUBLES = BM * DP;
    static constexpr size_t SMEM_BYTES =
        size_t(NSTAGE) * STAGE_BYTES + 2 * PHI_DOUBLES * 8 + 2 * XS_DOUBLES * 8 + NSTAGE * 8;
    static_assert(KC % 4 == 0, "KC must be a multiple of the DMMA k (4)");
    static_assert(STAGE_BYTES % 16 == 0, "TMA bulk copies need 16 B multiples");
    static_assert(NTHREADS <= 1024, "too many warps");
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

// phi_d(r): src/autoFRK.cpp:127-128 (d=1), :135-138 (d=2), :145-148 (d=3)
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

template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1) mrts_predict_kernel(const PredictArgs a) {
    constexpr int D = C::D, DP = C::DP, BM = C::BM, BN = C::BN, KC = C::KC, KQ = C::KQ;
    constexpr int MT = C::MT, NT = C::NT, NSTAGE = C::NSTAGE, NTHREADS = C::NTHREADS;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const stage0 = reinterpret_cast<double*>(smem_raw);
    double* const phib = stage0 + size_t(NSTAGE) * C::STAGE_DOUBLES;
    double* const xs = phib + 2 * C::PHI_DOUBLES;
    uint64_t* const full = reinterpret_cast<uint64_t*>(xs + 2 * C::XS_DOUBLES);

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

    // coordinates of one row tile -> shared (zero for rows past n2; their results are never stored)
    auto load_xs = [&](int buf, int64_t tile) {
        double* dst = xs + buf * C::XS_DOUBLES;
        for (int idx = tid; idx < BM * D; idx += NTHREADS) {
            const int r = idx % BM, c = idx / BM;
            const int64_t i = tile * BM + r;
            dst[r * DP + c] = (i < a.n2) ? __ldg(a.xnew + i + c * a.ldx) : 0.0;
        }
    };
    // Phi(rows of tile-buffer xbuf, knots of pipeline iteration `it`) -> A-fragment image
    auto compute_phi = [&](int64_t it, int xbuf) {
        const int s = int(it % NSTAGE);
        mbar_wait(&full[s], uint32_t(it / NSTAGE) & 1u);
        const double* knots = stage0 + size_t(s) * C::STAGE_DOUBLES + KC * BN;
        const double* xsb = xs + xbuf * C::XS_DOUBLES;
        double* dst = phib + int(it & 1) * C::PHI_DOUBLES;
#pragma unroll
        for (int v = tid; v < BM * KC; v += NTHREADS) {
            const int q = v / (BM * 4), rem = v % (BM * 4);
            const int row = rem >> 2, kk = q * 4 + (rem & 3);
            dst[v] = tps_phi<D>(xsb + row * DP, knots + kk * DP);
        }
    };

    load_xs(0, blockIdx.x);
    __syncthreads();  // barrier inits + xs[0] visible
    compute_phi(0, 0);
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
            if (it + 1 < total) compute_phi(it + 1, (c + 1 == nchunks) ? (xb ^ 1) : xb);

            const int s = int(it % NSTAGE);
            const double* ws = stage0 + size_t(s) * C::STAGE_DOUBLES + (wc * NT * 8) * 4 + lane;
            const double* ph = phib + int(it & 1) * C::PHI_DOUBLES + (wr * MT * 8) * 4 + lane;
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
                    for (int nt = 0; nt < NT; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
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

#define FRK_PREDICT_CONFIG_TABLE(D)                                                                        \
    predict_info<PredictCfg<D, 4, 4, 4, 1>>(),  /* BM 128, BN  32 */                                       \
    predict_info<PredictCfg<D, 4, 4, 4, 2>>(),  /* BM 128, BN  64 */                                       \
    predict_info<PredictCfg<D, 4, 4, 4, 3>>(),  /* BM 128, BN  96 */                                       \
    predict_info<PredictCfg<D, 4, 4, 4, 4>>(),  /* BM 128, BN 128 */                                       \
    predict_info<PredictCfg<D, 2, 8, 4, 3>>(),  /* BM  64, BN 192 */                                       \
    predict_info<PredictCfg<D, 3, 5, 4, 5>>(),  /* BM  96, BN 200 */                                       \
    predict_info<PredictCfg<D, 2, 8, 4, 4>>(),  /* BM  64, BN 256 */                                       \
    predict_info<PredictCfg<D, 2, 8, 4, 5>>(),  /* BM  64, BN 320 */                                       \
    predict_info<PredictCfg<D, 1, 16, 4, 3>>(), /* BM  32, BN 384 */                                       \
    predict_info<PredictCfg<D, 1, 16, 4, 4>>()  /* BM  32, BN 512 */

}  // namespace frk
