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
//       - one TMA bulk copy (cp.async.bulk -> UBLKCP) per chunk stages W, tracked by an mbarrier,
//         NSTAGE deep, refilled by whichever warp is last to finish with a stage,
//       - every B-fragment load is a contiguous, conflict-free 256 B LDS.64 per warp.
//     (The shared-memory bank conflicts ncu counts on this kernel -- 2.69e9 excess wavefronts on the bench launch --
//     are not fragment loads: 92% of them sit on ONE instruction, the LDS.128 gather of the d = 2 log table
//     (phi2_table: a data-dependent index per lane, ~6 excess wavefronts per warp instruction), the rest on the Phi
//     stores; profiles/k1_r02_source_ncu_summary.txt.  The shared-memory pipe runs at 29% of its peak, so they are
//     not on the critical path; a gather from a 16 KB table cannot be conflict-free.)
//   * The FP64 ALU and the DMMA unit share one pipe on B200 (measured: profiles/fp64_peaks_r01.json),
//     so kernel evaluation is pure overhead on the binding resource.  Every thread therefore
//     evaluates the same number of Phi values for chunk c+1 INSIDE the DMMA block of chunk c (one
//     basic block, branch-free table kernel for d = 2 and MUFU-seeded sqrt for d = 3), letting the scheduler drop the Phi arithmetic
//     into the issue slots between dependent-free DMMAs instead of serialising it in front of them.
//   * no CTA-wide barrier in the steady state: Phi is produced two chunks ahead and handed over through
//     mbarriers, so warps drift up to a chunk apart and the pipe does not drain at chunk boundaries.
//   * persistent CTAs (one per SM), tiles of BM rows dealt round-robin.
#pragma once
#include <cstdlib>
#include "frk_common.cuh"
#include "tps_log_table.h"

namespace frk {

template <int D_, int WR_, int WC_, int MT_, int NT_, int KC_, int NSTAGE_, int AHEAD_ = 2, bool SPLIT_ = false,
          bool KSM_ = false>
struct PredictCfg {
    static constexpr int D = D_, WR = WR_, WC = WC_, MT = MT_, NT = NT_, KC = KC_, NSTAGE = NSTAGE_;
    static constexpr int DP = (D == 3) ? 4 : D;   // doubles per stored coordinate tuple
    static constexpr int NWARPS = WR * WC;
    static constexpr int NTHREADS = NWARPS * 32;
    static constexpr int BM = WR * MT * 8;        // rows (locations) per tile
    static constexpr int BN = WC * NT * 8;        // columns per slice (padded)
    static constexpr int KQ = KC / 4;             // k4 groups per chunk
    static constexpr int VPT = (BM * KC + NTHREADS - 1) / NTHREADS;  // Phi values per thread per chunk
    static constexpr int STAGE_DOUBLES = KC * BN;   // packed W chunk: [k4 group][column][4 knots]
    static constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;
    static constexpr int PHI_DOUBLES = BM * KC;
    static constexpr int XS_DOUBLES = BM * DP;
    static constexpr int LOGTAB_DOUBLES = (D == 2) ? 2 * kLogTabSize : 0;
    static constexpr int AHEAD = AHEAD_;          // Phi is produced this many chunks ahead of its DMMAs
    static constexpr bool PHI_SPLIT = SPLIT_;     // d = 2: overlap the halves of consecutive Phi values (see the kernel)
    static constexpr bool KNOTS_SMEM = KSM_;      // knot coordinates staged in shared memory behind SMEM_BYTES (if they fit)
    using Fallback = PredictCfg<D_, WR_, WC_, MT_, NT_, KC_, NSTAGE_, AHEAD_, false, false>;  // when they do not
    static constexpr int NPB = 2 * AHEAD;         // Phi chunk buffers (>= 2 * AHEAD)
    static constexpr int MIN_CHUNKS = (2 * AHEAD > 4) ? 2 * AHEAD : 4;  // the plan pads the knot list to at least this many chunks
    static constexpr size_t SMEM_BYTES = size_t(NSTAGE) * STAGE_BYTES + NPB * PHI_DOUBLES * 8 + 2 * XS_DOUBLES * 8 +
                                         LOGTAB_DOUBLES * 8 + (NSTAGE + NPB) * 8 + NSTAGE * 4 + 16;
    static_assert(KC % 4 == 0, "KC must be a multiple of the DMMA k (4)");
    static_assert(STAGE_BYTES % 16 == 0, "TMA bulk copies need 16 B multiples");
    static_assert(NTHREADS <= 1024, "too many warps");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
    static_assert(NPB >= 2 * AHEAD && MIN_CHUNKS >= 2 * AHEAD, "Phi ring depth");
};

struct PredictArgs {
    const double* xnew;   // n2 x d, column-major, leading dimension ldx
    int64_t ldx;
    int64_t n2;
    const double* wpack;  // nchunks * STAGE_DOUBLES (packed W)
    const double* knots;  // nchunks * KC * DP knot coordinates (tuples of DP doubles), padded with knot 0
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

// Branch-free d = 2 kernel value phi(x) = x log(x) / (16 pi) for x = r^2 >= 0 (Tang-style, 1024-entry table
// in shared memory): x = 2^e m, m in [1,2); j = top 10 mantissa bits, r = m*rc_j - 1 (one FMA, |r| <= 2^-11);
//   log(x)/(16 pi) = e ln2/(16 pi) + lcc_j + c (r - r^2/2 + r^3/3 - r^4/4),   c = 1/(16 pi)   (|c r^5/5| < 2e-19)
// 11 FP64-pipe instructions per pair, r^2 included (libm log alone is ~50 pipe slots); absolute error 5.1e-18 on (0,2) against
// 6.6e-18 for x*log(x)/(16 pi) with libm (tools/gen_log_table.py).  e*ln2 needs no hi/lo split: its
// rounding error is multiplied by x = 2^e m, i.e. it is <= 1 ulp of phi.  x == 0 is masked; subnormal x gives a
// value <= 1e-305 in magnitude; inf / nan propagate through the final multiplication by x.
__device__ __forceinline__ double phi2_table(double x, const double2* __restrict__ tab) {
    const long long bits = __double_as_longlong(x);
    const int hi = int(bits >> 32);
    const double2 tc = tab[(hi >> (20 - kLogTabBits)) & (kLogTabSize - 1)];
    const double ef = double((hi >> 20) - 1023);
    const double m = __longlong_as_double((bits & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
    const double r = fma(m, tc.x, -1.0);
    constexpr double c = 0x1.45f306dc9c883p-6;  // 1/(16 pi) = 0.019894367886486918
    double p = fma(-0.25 * c, r, c / 3.0);
    p = fma(p, r, -0.5 * c);
    p = fma(p, r, c);
    const double s = fma(r, p, tc.y);
    const double L = fma(ef, 0x1.c3dc98f7e969cp-7, s);  // ln2 / (16 pi) = 0.013789725009540725
    return (bits > 0) ? x * L : 0.0;   // x = r^2 >= 0: an integer test on the bits keeps the comparison off the FP64 pipe
}

// phi2_table cut in two so that a thread can overlap the halves of two consecutive values (the first half of the
// next value with the second half of the current one): the 10-deep dependent FP64 chain becomes two independent
// ~5-deep chains in one basic block.  Same operations in the same order: bitwise identical to phi2_table.
struct Phi2Half {
    double r, x, lcc;
    int e;
};
__device__ __forceinline__ Phi2Half phi2_table_head(double x, const double2* __restrict__ tab) {
    const long long bits = __double_as_longlong(x);
    const int hi = int(bits >> 32);
    const double2 tc = tab[(hi >> (20 - kLogTabBits)) & (kLogTabSize - 1)];
    const double m = __longlong_as_double((bits & 0x000FFFFFFFFFFFFFLL) | 0x3FF0000000000000LL);
    Phi2Half h;
    h.r = fma(m, tc.x, -1.0);
    h.x = x;
    h.lcc = tc.y;
    h.e = (hi >> 20) - 1023;
    return h;
}
__device__ __forceinline__ double phi2_table_tail(const Phi2Half& h) {
    constexpr double c = 0x1.45f306dc9c883p-6;  // 1/(16 pi)
    double p = fma(-0.25 * c, h.r, c / 3.0);
    p = fma(p, h.r, -0.5 * c);
    p = fma(p, h.r, c);
    const double s = fma(h.r, p, h.lcc);
    const double L = fma(double(h.e), 0x1.c3dc98f7e969cp-7, s);  // ln2 / (16 pi)
    return (__double_as_longlong(h.x) > 0) ? h.x * L : 0.0;
}

// Branch-free FP64 sqrt of x >= 0: MUFU seed (rsqrt.approx.f64, ~2^-21), two Newton steps on 1/sqrt, one
// correction of x*y.  Correctly rounded on 3e7 random inputs with a 2^-21-perturbed seed; keeps
// the d = 3 kernel in the DMMA basic block (libm sqrt carries a slow-path branch).
__device__ __forceinline__ double sqrt_pos_fast(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    y = y * fma(-(h * y), y, 1.5);
    y = y * fma(-(h * y), y, 1.5);
    double r = x * y;
    r = fma(fma(-r, r, x), 0.5 * y, r);
    r = (x >= 2.2250738585072014e-308) ? r : 0.0;  // x == 0 and subnormal x (r < 1.5e-154) -> 0
    return (x < __longlong_as_double(0x7FF0000000000000LL)) ? r : x;  // inf / nan propagate
}

// kernel value from pre-loaded coordinates (registers): same formulas as tps_phi, table log for d = 2
template <int D>
__device__ __forceinline__ double tps_phi_fast(const double (&x)[3], const double (&u)[3], const double2* __restrict__ tab) {
    if constexpr (D == 1) {
        const double r = fabs(x[0] - u[0]);
        return (r * r * r) * (1.0 / 12.0);
    } else if constexpr (D == 2) {
        const double dx = x[0] - u[0], dy = x[1] - u[1];
        return phi2_table(fma(dx, dx, dy * dy), tab);
    } else {
        const double dx = x[0] - u[0], dy = x[1] - u[1], dz = x[2] - u[2];
        return -0.125 * sqrt_pos_fast(fma(dx, dx, fma(dy, dy, dz * dz)));
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
    constexpr int NPB = C::NPB, NWARPS = C::NWARPS, AHEAD = C::AHEAD;
#ifdef FRK_PHI_Q   // tuning aid: where in the DMMA block the Phi arithmetic is placed in the source
    constexpr int PHI_Q = (FRK_PHI_Q < KQ) ? FRK_PHI_Q : KQ - 1;
#else
    constexpr int PHI_Q = 0;  // k4-group after which the Phi arithmetic is placed in the source (ptxas spreads it anyway)
#endif
    constexpr bool KSM = C::KNOTS_SMEM;       // knots in shared memory: no global loads in the loop, no prefetch registers
    constexpr bool PREFETCH = (VPT <= 2) && !KSM;  // knots one iteration early in registers (if they are affordable)
    // d = 2, one Phi value per thread and chunk: halves of consecutive values overlapped (phi2_table_head / _tail);
    // with two values per thread the extra state spills.  Opt-in per configuration: whether ptxas turns the two
    // shorter chains into a better schedule is measured, not assumed (BN 384: +1.6%, BN 512: -2.5%).
    constexpr bool SPLIT = C::PHI_SPLIT && (VPT == 1) && (D == 2);
    constexpr size_t KNOTS_OFFSET = (C::SMEM_BYTES + 15) / 16 * 16;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* const stage0 = reinterpret_cast<double*>(smem_raw);
    double* const phib = stage0 + size_t(NSTAGE) * C::STAGE_DOUBLES;
    double* const xs = phib + NPB * C::PHI_DOUBLES;
    double* const logtab_d = xs + 2 * C::XS_DOUBLES;
    uint64_t* const w_full = reinterpret_cast<uint64_t*>(logtab_d + C::LOGTAB_DOUBLES);  // TMA landed
    uint64_t* const phi_full = w_full + NSTAGE;    // all warps stored their part of a Phi chunk
    uint32_t* const w_done = reinterpret_cast<uint32_t*>(phi_full + NPB);  // warps done with a W stage
    const double2* const logtab = reinterpret_cast<const double2*>(logtab_d);
    double* const knots_s = reinterpret_cast<double*>(smem_raw + KNOTS_OFFSET);   // KSM only
    const double* const kn = KSM ? knots_s : a.knots;
    auto ldk = [&](const double* p) { return KSM ? *p : __ldg(p); };

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wr = warp / C::WC, wc = warp % C::WC;

    if (int64_t(blockIdx.x) >= a.ntiles) return;
    const int64_t ntl = (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;  // tiles of this CTA
    const int nchunks = a.nchunks;        // >= 4 (the plan pads with zero-weight chunks)
    const int64_t total = ntl * nchunks;  // (tile, chunk) iterations of this CTA

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(&w_full[s], 1);
            w_done[s] = 0;
        }
        for (int b = 0; b < NPB; ++b) mbar_init(&phi_full[b], NWARPS);
        mbar_fence_init();
        for (int s = 0; s < NSTAGE; ++s)
            if (s < total) {
                mbar_expect_tx(&w_full[s], C::STAGE_BYTES);
                tma_bulk_g2s(stage0 + size_t(s) * C::STAGE_DOUBLES,
                             a.wpack + size_t(s % nchunks) * C::STAGE_DOUBLES, C::STAGE_BYTES, &w_full[s]);
            }
    }
    if constexpr (KSM) {
        const int nk = a.nchunks * KC * DP;
        for (int i = tid; i < nk; i += NTHREADS) knots_s[i] = __ldg(a.knots + i);
    }
    if constexpr (D == 2) {
        for (int i = tid; i < kLogTabSize; i += NTHREADS) {
            logtab_d[2 * i] = kLogTableDev[i].rc;
            logtab_d[2 * i + 1] = kLogTableDev[i].lcc;
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
    auto knots_load = [&](int chunk, double (&dst)[VPT][3]) {
        const double* knots = kn + size_t(chunk) * (KC * DP);
#pragma unroll
        for (int i = 0; i < VPT; ++i)
#pragma unroll
            for (int c = 0; c < D; ++c) dst[i][c] = ldk(knots + phi_kk[i] * DP + c);
    };
    auto rows_load = [&](int xbuf) {
        const double* xsb = xs + xbuf * C::XS_DOUBLES;
#pragma unroll
        for (int i = 0; i < VPT; ++i)
#pragma unroll
            for (int c = 0; c < D; ++c) px[i][c] = xsb[phi_row[i] * DP + c];
    };
    auto phi_store = [&](int buf, const double (&val)[VPT]) {
        double* dst = phib + buf * C::PHI_DOUBLES;
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int v = tid + i * NTHREADS;
            if (VPT * NTHREADS == BM * KC || v < BM * KC) dst[v] = val[i];
        }
    };

    double pun[PREFETCH ? VPT : 1][3];  // prefetched knots (global/L2 latency off the slowest warp's path)
    load_xs(0, blockIdx.x);
    __syncthreads();  // barrier inits, log table and xs[0] visible
#pragma unroll
    for (int j = 0; j < AHEAD; ++j) {  // Phi of the first AHEAD chunks (nchunks >= 4: all in tile 0)
        knots_load(j, pu);
        rows_load(0);
        double val[VPT];
#pragma unroll
        for (int i = 0; i < VPT; ++i) val[i] = tps_phi_fast<D>(px[i], pu[i], logtab);
        phi_store(j, val);
        __syncwarp();
        if (lane == 0) mbar_arrive(&phi_full[j]);
    }
    [[maybe_unused]] Phi2Half half[SPLIT ? VPT : 1];   // SPLIT: first half of Phi(it + AHEAD), computed during iteration it - 1
    if constexpr (SPLIT) {            // first half of Phi(AHEAD) (tile 0: nchunks >= 2 AHEAD), knots of chunk AHEAD + 1
        knots_load(AHEAD, pu);
        rows_load(0);
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const double dx = px[i][0] - pu[i][0], dy = px[i][1] - pu[i][1];
            half[i] = phi2_table_head(fma(dx, dx, dy * dy), logtab);
        }
    }
    if constexpr (PREFETCH) {
        const double* k0 = kn + size_t((AHEAD + (SPLIT ? 1 : 0)) % nchunks) * (KC * DP);
#pragma unroll
        for (int i = 0; i < VPT; ++i)
#pragma unroll
            for (int cc = 0; cc < D; ++cc) pun[i][cc] = __ldg(k0 + phi_kk[i] * DP + cc);
    }

    // Warps synchronise only through mbarriers: Phi(it) is produced during iteration it-AHEAD, so a warp
    // may start the DMMAs of chunk `it` as soon as every warp has finished chunk it-AHEAD.  The SM's
    // warp arbiter is strictly prioritised, so the slowest warp regularly runs alone; everything outside
    // the DMMA block is therefore kept off its critical path: 32-bit ring counters instead of 64-bit
    // div/mod, barrier probes and knot loads issued one iteration early, a relaxed counter to elect the
    // warp that refills a W stage.
    int s = 0;              // W stage of the current chunk, parity of its use
    uint32_t wpar = 0;
    int pb = 0;             // Phi buffer of the current chunk, parity of its use
    uint32_t ppar = 0;
    int pbw = AHEAD % NPB;  // Phi buffer written this iteration: (pb + AHEAD) mod NPB
    int cphi = AHEAD % nchunks;            // chunk (within its tile) whose Phi is produced this iteration
    int cref = NSTAGE % nchunks;           // chunk that refills the stage released this iteration
    int left = int(total);                 // iterations still to run, including this one (host checks < 2^31)
    bool okp = false, okw = false;         // early (non-blocking) probes of this iteration's barriers
    for (int64_t til = 0; til < ntl; ++til) {
        const int64_t tile = blockIdx.x + til * gridDim.x;
        const int xb = int(til & 1);
        // Phi of the tile's first chunk complete: every warp has finished iteration it - AHEAD, hence nobody reads
        // the previous tile's rows any more -> stage the next tile's (kept out of the chunk loop)
        if (!okp) mbar_wait(&phi_full[pb], ppar);
        okp = true;
        if (til + 1 < ntl) load_xs(xb ^ 1, tile + gridDim.x);
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

        for (int c = 0; c < nchunks; ++c) {
            // (1) Phi of THIS chunk complete: every warp has finished iteration it - AHEAD
            if (!okp) mbar_wait(&phi_full[pb], ppar);
            // (2) inputs of the Phi chunk produced in this iteration (rows from shared, knots prefetched)
            const bool more = (left > AHEAD);
            // without prefetch: knots of the Phi chunk produced now; with: those of the next iteration's
            // SPLIT: the registers hold the knots of chunk cphi + 1 (its first half runs now), loads fetch cphi + 2
            constexpr int KOFF = (SPLIT ? 1 : 0) + (PREFETCH ? 1 : 0);   // chunk whose knots are read in this iteration
            const int cknot = (cphi + KOFF >= nchunks) ? cphi + KOFF - nchunks : cphi + KOFF;
            const double* knots_c = kn + size_t(cknot) * (KC * DP);
            double pu_now[PREFETCH ? VPT : 1][3];
            if constexpr (PREFETCH) {
#pragma unroll
                for (int i = 0; i < VPT; ++i)
#pragma unroll
                    for (int cc = 0; cc < D; ++cc) {
                        pu_now[i][cc] = pun[i][cc];
                        pun[i][cc] = __ldg(knots_c + phi_kk[i] * DP + cc);
                    }
            }
            const double* xs_c = xs + ((c + AHEAD + (SPLIT ? 1 : 0) >= nchunks) ? (xb ^ 1) : xb) * C::XS_DOUBLES;
            // (3) W of this chunk landed
            if (!okw) mbar_wait(&w_full[s], wpar);

            const double* ws = stage0 + s * C::STAGE_DOUBLES + (wc * NT * 8) * 4 + lane;
            const double* ph = phib + pb * C::PHI_DOUBLES + (wr * MT * 8) * 4 + lane;
            // ring positions of the next iteration
            const int sn = (s + 1 == NSTAGE) ? 0 : s + 1;
            const uint32_t wparn = (s + 1 == NSTAGE) ? (wpar ^ 1u) : wpar;
            const int pbn = (pb + 1 == NPB) ? 0 : pb + 1;
            const uint32_t pparn = (pb + 1 == NPB) ? (ppar ^ 1u) : ppar;
            double val[VPT];
            // ---- one basic block: DMMAs of chunk `it` + Phi arithmetic of chunk `it + AHEAD` ----
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
                if (q == PHI_Q) {  // the scheduler hoists these loads as far as register pressure allows
#pragma unroll
                    for (int i = 0; i < VPT; ++i) {
                        double x1[3], u1[3];
#pragma unroll
                        for (int cc = 0; cc < D; ++cc) {
                            x1[cc] = xs_c[phi_row[i] * DP + cc];
                            if constexpr (PREFETCH) u1[cc] = pu_now[i][cc];
                            else u1[cc] = ldk(knots_c + phi_kk[i] * DP + cc);
                        }
                        if constexpr (SPLIT) {
                            val[i] = phi2_table_tail(half[i]);               // Phi(it + AHEAD), started last iteration
                            const double dx = x1[0] - u1[0], dy = x1[1] - u1[1];
                            half[i] = phi2_table_head(fma(dx, dx, dy * dy), logtab);   // Phi(it + AHEAD + 1)
                        } else {
                            val[i] = tps_phi_fast<D>(x1, u1, logtab);
                        }
                    }
                }
                if (q == KQ - 1) {  // probe the next iteration's barriers while the last DMMAs drain
                    okp = mbar_test_wait(&phi_full[pbn], pparn);
                    okw = mbar_test_wait(&w_full[sn], wparn);
                }
            }
            // (4) hand Phi(it+AHEAD) over.  Its buffer ((pb + AHEAD) mod NPB, NPB = 2 AHEAD) was last read by the DMMAs of chunk
            //     it+AHEAD-NPB, which every warp finished before (1) could pass.
            if (more) phi_store(pbw, val);
            __syncwarp();
            if (lane == 0) {
                if (more) mbar_arrive(&phi_full[pbw]);
                // (5) the last warp to finish with W stage s refills it for chunk it + NSTAGE
                if (smem_add_relaxed(&w_done[s], 1u) == uint32_t(NWARPS - 1)) {
                    w_done[s] = 0;
                    if (left > NSTAGE) {
                        mbar_expect_tx(&w_full[s], C::STAGE_BYTES);
                        tma_bulk_g2s(stage0 + s * C::STAGE_DOUBLES, a.wpack + size_t(cref) * C::STAGE_DOUBLES,
                                     C::STAGE_BYTES, &w_full[s]);
                    }
                }
            }
            s = sn;
            wpar = wparn;
            pb = pbn;
            ppar = pparn;
            pbw = (pbw + 1 == NPB) ? 0 : pbw + 1;
            cphi = (cphi + 1 == nchunks) ? 0 : cphi + 1;
            cref = (cref + 1 == nchunks) ? 0 : cref + 1;
            --left;
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
    int d, bm, bn, kc, nstage, dp, nthreads, stage_doubles, min_chunks;
    size_t smem_bytes;
    void (*launch)(const PredictArgs&, int grid, cudaStream_t);
    void (*prepare)();  // opt in to large dynamic shared memory (once)
};

#ifdef FRK_EXPERIMENTAL_CONFIGS   // FRK_SMEM_PAD=<bytes>: extra (unused) dynamic shared memory, shrinks the L1 carve-out
inline size_t predict_smem_pad(size_t base) {
    const char* e = std::getenv("FRK_SMEM_PAD");
    const size_t pad = e ? size_t(std::atol(e)) : 0;
    return (base + pad <= 227 * 1024) ? pad : 0;
}
#else
inline size_t predict_smem_pad(size_t) { return 0; }
#endif
constexpr size_t kMaxDynSmem = 227 * 1024;
template <class C>
size_t predict_knots_smem_total(const PredictArgs& a) {   // dynamic shared memory with the knots staged behind the rest
    return (C::SMEM_BYTES + 15) / 16 * 16 + size_t(a.nchunks) * C::KC * C::DP * sizeof(double);
}
template <class C>
void predict_prepare() {
    if constexpr (C::KNOTS_SMEM) {
        FRK_CUDA(cudaFuncSetAttribute(mrts_predict_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kMaxDynSmem)));
        predict_prepare<typename C::Fallback>();
    } else {
        FRK_CUDA(cudaFuncSetAttribute(mrts_predict_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(C::SMEM_BYTES + predict_smem_pad(C::SMEM_BYTES))));
    }
}
template <class C>
void predict_launch(const PredictArgs& a, int grid, cudaStream_t st) {
    if constexpr (C::KNOTS_SMEM) {
        const size_t total = predict_knots_smem_total<C>(a);
        if (total <= kMaxDynSmem) mrts_predict_kernel<C><<<grid, C::NTHREADS, total, st>>>(a);
        else predict_launch<typename C::Fallback>(a, grid, st);   // too many knots for shared memory
    } else {
        mrts_predict_kernel<C><<<grid, C::NTHREADS, C::SMEM_BYTES + predict_smem_pad(C::SMEM_BYTES), st>>>(a);
    }
}
template <class C>
constexpr PredictKernelInfo predict_info() {
    return PredictKernelInfo{C::D, C::BM, C::BN, C::KC, C::NSTAGE, C::DP, C::NTHREADS, C::STAGE_DOUBLES,
                             C::MIN_CHUNKS, C::SMEM_BYTES, &predict_launch<C>, &predict_prepare<C>};
}

// the instantiated configurations, per dimension (defined in mrts_predict_d{1,2,3}.cu)
const PredictKernelInfo* predict_kernels_d1(int* count);
const PredictKernelInfo* predict_kernels_d2(int* count);
const PredictKernelInfo* predict_kernels_d3(int* count);

// <D, WR, WC, MT, NT, KC, NSTAGE>: KC chosen so that every thread evaluates a whole number of Phi
// values per chunk (BM*KC is a multiple of the thread count)
#define FRK_PREDICT_CONFIG_TABLE(D)                                                                        \
    predict_info<PredictCfg<D, 16, 1, 4, 1, 8, 4>>(),  /* BM 512, BN   8: Phi-bound (predict.FRK, T = 1)  */  \
    predict_info<PredictCfg<D, 8, 2, 4, 1, 16, 4>>(),  /* BM 256, BN  16 */                                \
    predict_info<PredictCfg<D, 4, 4, 4, 1, 32, 4>>(),  /* BM 128, BN  32, 8 Phi/thread */                  \
    predict_info<PredictCfg<D, 4, 4, 4, 2, 32, 4>>(),  /* BM 128, BN  64 */                                \
    predict_info<PredictCfg<D, 4, 4, 4, 3, 16, 4, 2, false, true>>(), /* BM 128, BN  96, knots in smem */   \
    predict_info<PredictCfg<D, 4, 4, 4, 4, 16, 4, 2, false, true>>(), /* BM 128, BN 128, knots in smem */   \
    predict_info<PredictCfg<D, 2, 8, 4, 3, 16, 4>>(),  /* BM  64, BN 192, 2 Phi/thread */                  \
    predict_info<PredictCfg<D, 3, 5, 4, 5, 20, 3, 3, false, true>>(), /* BM 96, BN 200, 4 Phi/thread, 15 warps, Phi 3 ahead, knots in smem */ \
    predict_info<PredictCfg<D, 2, 8, 4, 4, 16, 4, 2, false, true>>(), /* BM  64, BN 256, knots in smem */   \
    predict_info<PredictCfg<D, 2, 8, 4, 5, 16, 3, 3, false, true>>(), /* BM 64, BN 320, Phi 3 ahead, knots in smem */ \
    predict_info<PredictCfg<D, 1, 16, 4, 3, 16, 3, 3, true>>(), /* BM 32, BN 384, 1 Phi/thread, Phi 3 ahead, split Phi */ \
    predict_info<PredictCfg<D, 1, 16, 4, 4, 16, 2, 3, false, true>>() /* BM 32, BN 512, 1 Phi/thread, 2 x 64 KB stages, Phi 3 ahead, knots in smem */ \
    FRK_PREDICT_EXTRA_CONFIGS(D)
#ifdef FRK_EXPERIMENTAL_CONFIGS   // Phi ring depth study (tools/k1_cfg_probe.py with FRK_K1_CONFIG)
#define FRK_PREDICT_EXTRA_CONFIGS(D)                                                                       \
    , predict_info<PredictCfg<D, 1, 16, 4, 4, 16, 2, 3>>()  /* 12: BN 512, knots from global memory (v7) */  \
    , predict_info<PredictCfg<D, 2, 8, 4, 5, 16, 3, 3>>()   /* 13: BN 320, knots from global memory */      \
    , predict_info<PredictCfg<D, 2, 8, 4, 4, 16, 4, 2>>()   /* 14: BN 256, knots from global memory */      \
    , predict_info<PredictCfg<D, 1, 16, 4, 3, 16, 3, 3, false, true>>() /* 15: BN 384, knots in smem, single Phi chain */ \
    , predict_info<PredictCfg<D, 1, 16, 4, 4, 16, 3, 2>>() /* 16: BN 512, 3 stages, AHEAD 2 (v6) */         \
    , predict_info<PredictCfg<D, 1, 16, 4, 4, 16, 2, 3, true>>() /* 17: BN 512 with split Phi */            \
    , predict_info<PredictCfg<D, 1, 8, 4, 8, 16, 3>>()     /* 18: BN 512, 8 fat warps (32 x 64 tiles) */     \
    , predict_info<PredictCfg<D, 4, 5, 2, 5, 20, 4>>()     /* 19: BN 200, 20 thin warps, 2 Phi/thread */   \
    , predict_info<PredictCfg<D, 3, 5, 4, 5, 20, 4>>()     /* 20: BN 200, 4 stages, Phi 2 ahead, knots from global memory (r01) */ \
    , predict_info<PredictCfg<D, 4, 4, 4, 4, 16, 4>>()     /* 21: BN 128, knots from global memory (r01) */ \
    , predict_info<PredictCfg<D, 4, 4, 4, 3, 16, 4>>()     /* 22: BN 96, knots from global memory (r01) */
#else
#define FRK_PREDICT_EXTRA_CONFIGS(D)
#endif

}  // namespace frk
