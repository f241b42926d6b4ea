// K4, register-resident variant for K + T <= 256 columns (K <= 128 basis functions and few replicates, the shapes
// default autoFRK() selects): G = F' W F, C = F' W Z and zz = sum(w Z^2) in ONE pass and ONE launch (+ the reduction).
//
// Between the HBM roof (K ~ 32: 288 B and 1,792 tensor flops per row) and the DMMA roof (K = 128: 38 DMMAs per row)
// what matters is that (i) no DMMA is spent outside the upper triangle, (ii) the four SM sub-partitions issue the
// same number of DMMAs, (iii) the loads run far ahead.  So:
//   * the output is cut into 8 x 8 tiles (one DMMA accumulator each); a UNIT is up to 2 x 4 of them (two adjacent row
//     groups of F, four adjacent column groups of one 32-column block of F or Z): 6 fragment loads feed 8 DMMAs.  The
//     shape of a unit (nr x nc rectangle, or one of the two triangles of a diagonal block) is a template parameter:
//     a predicated-off DMMA still occupies the FP64 tensor pipe (measured: profiles/gram_mid_r02_v1_ncu_summary.txt),
//     so edge and diagonal units run their own code instead of a mask;
//   * the host deals the units to warps (two per warp), longest first, onto the least loaded sub-partition.  When
//     the units fit 8 (4) warps the 16 warps form 2 (4) SETS that take alternate row stages with the same unit
//     table, rotated across sub-partitions, so the pipe load is equal by construction;
//   * staging: one cp.async.bulk.tensor.2d box of [32 columns][36 rows], NOT swizzled, per column block and 32-row
//     stage.  The four extra rows are never read: they make the column pitch 36 doubles == 4 (mod 16), so the fragment
//     LDS.64 of a shared-memory phase (columns g = 0..3 x rows t = 0..3 of a k4-step, natural k order) fall on 16
//     different banks without any swizzle.  Why not the wide kernel's SWIZZLE_128B boxes of 16 rows: the TMA unit's
//     cost is per box ROW -- one CTA per SM streaming this column-major pattern with no arithmetic reaches 3.3-4.7 TB/s
//     with 128-byte rows against 6.4-7.2 TB/s with 256..544-byte rows (7.0 TB/s for a plain LDG.128 read;
//     tools/tma_stream_bench.cu, profiles/tma_stream_bench_r02.json).  A ring of up to 12 stages, `full` mbarriers,
//     last-finisher re-arm, no CTA barrier in the loop; ragged rows / columns are zero-filled by the TMA unit.
//     Also measured and dropped (profiles/gram_probe_r02_v3..v6): per-column 1-D bulk copies (>= 65 clocks each: need
//     2 KB per copy, and then two 75 KB stages are too coarse a ring), cp.async issued by the warp that re-arms a stage
//     (18..72 LDGSTS on the ring's critical path: 2x slower), 64/128-row stages shared by all sets (-8..-40%);
//   * zz is accumulated from the Z boxes already in shared memory by one warp of the stage's set (in rotation);
//   * one persistent CTA per SM owns a contiguous row range; partial tiles go to a workspace and a second kernel
//     adds them in a fixed order (bitwise reproducible) and mirrors G.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>
#include <type_traits>
#include <vector>
#include <cuda.h>
#include "frk_common.cuh"
#include "gram.cuh"

namespace frk {

namespace {

constexpr int MCB = 32;                    // columns per staged block
constexpr int MBOXR = 36;                  // rows per TMA box: the stage's 32 plus 4 that only pad the column pitch
constexpr int MKR = 32;                    // rows per stage
constexpr int MKQ = MKR / 4;               // k4-steps per stage
constexpr int MPITCH = MBOXR * 8;                     // bytes between columns of a staged block: 36 doubles == 4 (mod 16)
constexpr int MGROUP_BYTES = 8 * MPITCH;             // one 8-column group
constexpr int MSLOT_BYTES = MCB * MPITCH;            // 9 KB: a full 32-column block of one stage = one TMA box
constexpr int MNW = GRAM_MID_WARPS;
constexpr int MNT = MNW * 32;
constexpr int MNU = GRAM_MID_UNITS;        // units per warp
constexpr int MAXBLK = 8;
constexpr int MAXST = 16;
constexpr size_t MID_SMEM_BUDGET = 222 * 1024;   // eight 27 KB stages (three column blocks) must fit

struct MidArgs {
    int K, T, kb, nblocks, nst, nsets, ntiles;
    int stage_bytes;              // sum of the staged blocks (the last block of F and of Z is trimmed to its 8-column groups)
    int blk_off[MAXBLK], blk_col[MAXBLK], blk_map[MAXBLK];   // per block: byte offset in a stage, first column, tensor map
    int64_t n, rows_cta;
    const double* w;
    const GramMidUnit* units;     // [MNW][MNU] (device)
    double* partial;              // [grid][MNW][MNU * 8][64]
    double* zzcta;                // [grid]
};

__device__ __forceinline__ void mid_tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void dmma_m(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}


__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}

// All k4-steps of one stage for one unit of shape NR x NC (TRI: without tile (1,0)), in one basic block, so the loads
// of the next k4-step are scheduled under the DMMAs of this one.  Every fragment address is a register + immediate.
template <int NR, int NC, bool TRI, bool WEIGHTED>
__device__ __forceinline__ void mid_unit(uint32_t As, int lane_off, int ab, int bb, double (&acc)[2][4][2],
                                         const double* __restrict__ w, int64_t r0, int64_t r_end, int t) {
#pragma unroll
    for (int q = 0; q < MKQ; ++q) {
        const uint32_t pa = As + ab + lane_off + q * 32;       // rows 4q + t of this lane's column
        const uint32_t pb = As + bb + lane_off + q * 32;
        double af[NR], bf[NC];
#pragma unroll
        for (int i = 0; i < NR; ++i) af[i] = lds_f64(pa + i * MGROUP_BYTES);
#pragma unroll
        for (int j = 0; j < NC; ++j) bf[j] = lds_f64(pb + j * MGROUP_BYTES);
        if (WEIGHTED) {
            const int64_t r = r0 + 4 * q + t;
            const double wv = (r < r_end) ? __ldg(w + r) : 0.0;
#pragma unroll
            for (int i = 0; i < NR; ++i) af[i] *= wv;
        }
#pragma unroll
        for (int i = 0; i < NR; ++i)
#pragma unroll
            for (int j = 0; j < NC; ++j)
                if (!(TRI && i == 1 && j == 0)) dmma_m(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
}

template <bool WEIGHTED>
__device__ __forceinline__ void mid_unit_dispatch(int kind, uint32_t As, int lane_off, int ab, int bb,
                                                  double (&acc)[2][4][2], const double* __restrict__ w, int64_t r0,
                                                  int64_t r_end, int t) {
    switch (kind) {   // warp-uniform
#define FRK_MID_CASE(id, NR, NC, TRI) \
    case id: mid_unit<NR, NC, TRI, WEIGHTED>(As, lane_off, ab, bb, acc, w, r0, r_end, t); break;
        FRK_MID_CASE(1, 1, 1, false) FRK_MID_CASE(2, 1, 2, false) FRK_MID_CASE(3, 1, 3, false) FRK_MID_CASE(4, 1, 4, false)
        FRK_MID_CASE(5, 2, 1, false) FRK_MID_CASE(6, 2, 2, false) FRK_MID_CASE(7, 2, 3, false) FRK_MID_CASE(8, 2, 4, false)
        FRK_MID_CASE(GRAM_MID_TRI24, 2, 4, true) FRK_MID_CASE(GRAM_MID_TRI22, 2, 2, true)
#undef FRK_MID_CASE
        default: break;
    }
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(MNT, 1)
gram_mid_kernel(const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                const __grid_constant__ CUtensorMap tm2, const __grid_constant__ CUtensorMap tm3, const MidArgs a) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* const base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = a.stage_bytes;
    uint64_t* const full = reinterpret_cast<uint64_t*>(base + size_t(a.nst) * stage_bytes);
    uint32_t* const done = reinterpret_cast<uint32_t*>(full + MAXST);
    double* const zzw = reinterpret_cast<double*>(done + MAXST + (MAXST & 1));   // one partial per warp

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wps = MNW / a.nsets;              // warps per set
    const int set = warp / wps;
    // the unit table is written for a set's logical warps; set s runs logical warp j on physical warp
    // s*wps + (j + s*rot) % wps, which moves it to another SM sub-partition (rot = 1 for 4-warp sets, 2 for 8)
    const int rot = (a.nsets == 4) ? 1 : (a.nsets == 2 ? 2 : 0);
    const int jlog = ((warp - set * wps) - set * rot % wps + wps) % wps;

    int ab[MNU], bb[MNU], kd[MNU];
#pragma unroll
    for (int u = 0; u < MNU; ++u) {
        const GramMidUnit x = a.units[jlog * MNU + u];
        ab[u] = x.a_base;
        bb[u] = x.b_base;
        kd[u] = x.kind;
    }
    bool active = false;
#pragma unroll
    for (int u = 0; u < MNU; ++u) active = active || (kd[u] != 0);
    // active warps per set (identical for every set) and this warp's rank among them, for the stage hand-back and zz duty
    int nact = 0, myrank = 0;
    for (int j = 0; j < wps; ++j) {
        bool act = false;
        for (int u = 0; u < MNU; ++u) act = act || (__ldg(&a.units[j * MNU + u].kind) != 0);
        if (act && j < jlog) ++myrank;
        nact += act ? 1 : 0;
    }

    if (tid == 0) {
        for (int s = 0; s < a.nst; ++s) {
            mbar_init(&full[s], 1);
            done[s] = 0;
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t r_begin = min(a.n, int64_t(blockIdx.x) * a.rows_cta);
    const int64_t r_end = min(a.n, r_begin + a.rows_cta);
    const int nsteps = (r_end > r_begin) ? int((r_end - r_begin + MKR - 1) / MKR) : 0;

    // the box this lane copies when its warp arms a stage: lane -> column block (one 36-row x 32-column box each)
    const int my_blk = lane;
    const bool my_has = my_blk < a.nblocks;
    const int my_map = my_has ? a.blk_map[my_blk] : 0;
    const CUtensorMap* const my_tm = my_map == 0 ? &tm0 : (my_map == 1 ? &tm1 : (my_map == 2 ? &tm2 : &tm3));
    const int my_col = my_has ? a.blk_col[my_blk] : 0;
    const int my_off = my_has ? a.blk_off[my_blk] : 0;
    const uint32_t stage_tx = uint32_t(stage_bytes);
    auto issue = [&](int slot, int step) {   // one whole warp
        if (lane == 0) mbar_expect_tx(&full[slot], stage_tx);
        __syncwarp();
        if (my_has)
            mid_tma_load_2d(base + size_t(slot) * stage_bytes + my_off, my_tm,
                            int(r_begin + int64_t(step) * MKR), my_col, &full[slot]);
    };
    if (warp == 0)
        for (int i = 0; i < a.nst && i < nsteps; ++i) issue(i, i);

    // byte offset of this lane's fragment element inside an 8-column group: column g, row t of a k4-step (rows 4q + t;
    // with a pitch of 36 doubles the 16 lanes of a shared-memory phase, g = 0..3 x t = 0..3, fall on 16 different banks)
    const int lane_off = g * MPITCH + t * 8;

    double acc[MNU][2][4][2];
#pragma unroll
    for (int u = 0; u < MNU; ++u)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[u][i][j][0] = acc[u][i][j][1] = 0.0;
    double zz = 0.0;

    if (active) {
        int slot = set % a.nst;
        uint32_t par = uint32_t(set / a.nst) & 1u;
        int duty = 0;   // counts this set's stages; the warp with rank duty % nact adds the stage's z^2
        for (int step = set; step < nsteps; step += a.nsets) {
            mbar_wait(&full[slot], par);
            const unsigned char* As = base + size_t(slot) * stage_bytes;
#pragma unroll
            for (int u = 0; u < MNU; ++u)
                mid_unit_dispatch<WEIGHTED>(kd[u], smem_u32(As), lane_off, ab[u], bb[u], acc[u], a.w, r_begin + int64_t(step) * MKR, r_end, t);
            if (a.T > 0 && duty % nact == myrank) {
                // 16-byte chunks (two rows) of the Z boxes: chunk e -> column e / 16, row pair e % 16
                const int nchunks = a.T * (MKR / 2);
                for (int e = lane; e < nchunks; e += 32) {
                    const int c = e / (MKR / 2), rp = e % (MKR / 2);
                    const int blk = a.kb + c / MCB, cc = c % MCB;
                    const double2 v = *reinterpret_cast<const double2*>(As + a.blk_off[blk] + cc * MPITCH + rp * 16);
                    if (WEIGHTED) {
                        const int64_t r = r_begin + int64_t(step) * MKR + 2 * rp;
                        const double w0 = (r < r_end) ? __ldg(a.w + r) : 0.0, w1 = (r + 1 < r_end) ? __ldg(a.w + r + 1) : 0.0;
                        zz += fma(w0 * v.x, v.x, (w1 * v.y) * v.y);
                    } else {
                        zz += fma(v.x, v.x, v.y * v.y);
                    }
                }
            }
            ++duty;
            __syncwarp();
            uint32_t last = 0;
            if (lane == 0) last = (smem_add_relaxed(&done[slot], 1u) == uint32_t(nact - 1)) ? 1u : 0u;
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                if (lane == 0) done[slot] = 0;
                if (step + a.nst < nsteps) issue(slot, step + a.nst);
            }
            slot += a.nsets;
            while (slot >= a.nst) {
                slot -= a.nst;
                par ^= 1u;
            }
        }
    }

    // epilogue: tiles to the workspace, zz to one number per CTA (warps added in warp order)
    double* const out = a.partial + (size_t(blockIdx.x) * MNW + warp) * (MNU * 8 * 64);
#pragma unroll
    for (int u = 0; u < MNU; ++u)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (kd[u] != 0) {
                    double2* p = reinterpret_cast<double2*>(out + ((u * 8 + i * 4 + j) * 64 + g * 8 + 2 * t));
                    *p = make_double2(acc[u][i][j][0], acc[u][i][j][1]);
                }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zz += __shfl_xor_sync(0xffffffffu, zz, o);
    if (lane == 0) zzw[warp] = zz;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int wv = 0; wv < MNW; ++wv) s += zzw[wv];
        a.zzcta[blockIdx.x] = s;
    }
}


// One block per 8 x 8 output tile: four thread groups each add every fourth CTA's partials (all sets of a CTA in
// set order), combined as (s0 + s1) + (s2 + s3); block 0 also adds the per-CTA sums of squares in CTA order.
__global__ void __launch_bounds__(256)
gram_mid_reduce_kernel(const MidArgs a, const GramMidTile* __restrict__ tiles, int ncta, double* __restrict__ packed) {
    const GramMidTile tl = tiles[blockIdx.x];
    const int e = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int wps = MNW / a.nsets;
    const int rot = (a.nsets == 4) ? 1 : (a.nsets == 2 ? 2 : 0);
    double s = 0.0;
    for (int c = grp; c < ncta; c += 4)
        for (int st = 0; st < a.nsets; ++st) {
            const int wv = st * wps + (tl.warp + st * rot) % wps;
            s += a.partial[((size_t(c) * MNW + wv) * (MNU * 8) + tl.slot) * 64 + e];
        }
    __shared__ double sh[4][64];
    sh[grp][e] = s;
    __syncthreads();
    if (grp == 0) {
        s = (sh[0][e] + sh[1][e]) + (sh[2][e] + sh[3][e]);
        const int row = e >> 3, col = e & 7;
        const int ga = tl.ga * 8 + row;
        double* G = packed;
        double* C = packed + size_t(a.K) * a.K;
        if (ga < a.K) {
            if (tl.isz) {
                const int gb = tl.gb * 8 + col;
                if (gb < a.T) C[ga + size_t(gb) * a.K] = s;
            } else {
                const int gb = tl.gb * 8 + col;
                if (gb < a.K && ga <= gb) {
                    G[ga + size_t(gb) * a.K] = s;
                    G[gb + size_t(ga) * a.K] = s;
                }
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double z = 0.0;
        for (int c = 0; c < ncta; ++c) z += a.zzcta[c];
        packed[size_t(a.K) * a.K + size_t(a.K) * a.T] = z;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn mid_encode_fn() {
    static EncodeFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FRK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) fail(3, "cuTensorMapEncodeTiled is not available in this driver");
        return reinterpret_cast<EncodeFn>(p);
    }();
    return fn;
}
CUtensorMap mid_make_map(const double* ptr, int64_t rows, int cols, int64_t ld, int box_cols) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {cuuint64_t(rows), cuuint64_t(cols)};
    cuuint64_t gstr[1] = {cuuint64_t(ld) * 8};
    cuuint32_t box[2] = {MBOXR, cuuint32_t(box_cols)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = mid_encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(3, "cuTensorMapEncodeTiled failed (%d)", int(r));
    return tm;
}

}  // namespace

// ---- host: cut the upper triangle of G and the F x Z strip into units and deal them to warps ------------------
bool gram_mid_plan(int K, int T, GramMidPlan& P) {
    P = GramMidPlan{};
    if (K < 1 || T < 0) return false;
    const int MT = (K + 7) / 8, NZ = (T + 7) / 8;
    const int kb = (K + MCB - 1) / MCB, zb = (T + MCB - 1) / MCB;
    if (kb + zb > MAXBLK) return false;
    // units: {first row group, first column group (in block blk of F | Z), kind}; tiles listed as (a, b) local indices
    // stage layout: full 32-column blocks, the last block of F and of Z trimmed to its valid 8-column groups
    {
        int off = 0;
        for (int b = 0; b < kb + zb; ++b) {
            const bool isz = b >= kb;
            const int c0 = (isz ? b - kb : b) * MCB, ncol = isz ? T : K;
            const int w = std::min(MCB, (ncol - c0 + 7) / 8 * 8);
            P.blk_off[b] = off;
            P.blk_cols[b] = w;
            off += w * MPITCH;
        }
        P.stage_bytes = off;
    }
    struct U { GramMidUnit u; int cost; int ga0; bool isz; int gb0; };
    std::vector<U> units;
    auto add = [&](int ga0, bool isz, int gb0, int nr, int nc, int kind) {
        U x{};
        x.ga0 = ga0; x.isz = isz; x.gb0 = gb0;
        x.u.ga0 = ga0; x.u.gb0 = gb0; x.u.isz = isz ? 1 : 0;
        x.u.a_base = P.blk_off[ga0 / 4] + (ga0 % 4) * MGROUP_BYTES;
        x.u.b_base = P.blk_off[(isz ? kb : 0) + gb0 / 4] + (gb0 % 4) * MGROUP_BYTES;
        x.u.kind = kind ? kind : 1 + (nr - 1) * 4 + (nc - 1);
        x.cost = kind == GRAM_MID_TRI24 ? 7 : (kind == GRAM_MID_TRI22 ? 3 : nr * nc);
        units.push_back(x);
    };
    for (int ba = 0; ba < kb; ++ba) {
        const int na = std::min(4, MT - 4 * ba);            // valid row groups of F block ba
        // diagonal block: upper triangle of its na x na tiles
        if (na == 4) {
            add(4 * ba, false, 4 * ba, 2, 4, GRAM_MID_TRI24);
            add(4 * ba + 2, false, 4 * ba + 2, 2, 2, GRAM_MID_TRI22);
        } else if (na == 3) {
            add(4 * ba, false, 4 * ba, 1, 3, 0);
            add(4 * ba + 1, false, 4 * ba + 1, 2, 2, GRAM_MID_TRI22);
        } else if (na == 2) {
            add(4 * ba, false, 4 * ba, 2, 2, GRAM_MID_TRI22);
        } else {
            add(4 * ba, false, 4 * ba, 1, 1, 0);
        }
        // the blocks to its right, then the Z blocks: rectangles of <= 2 x 4
        for (int bb = ba + 1; bb < kb + zb; ++bb) {
            const bool isz = bb >= kb;
            const int g0 = 4 * (isz ? bb - kb : bb);
            const int nc = std::min(4, (isz ? NZ : MT) - g0);
            for (int i = 0; i < na; i += 2) add(4 * ba + i, isz, g0, std::min(2, na - i), nc, 0);
        }
    }
    if (int(units.size()) > MNW * MNU) return false;
    P.nsets = (int(units.size()) <= MNW * MNU / 4) ? 4 : (int(units.size()) <= MNW * MNU / 2 ? 2 : 1);
    {
        static const int max_sets = [] {
            const char* e = std::getenv("FRK_GRAM_MID_MAXSETS");   // tuning aid
            return e ? std::atoi(e) : 4;
        }();
        while (P.nsets > max_sets && P.nsets > 1) P.nsets /= 2;
    }
    const int wps = MNW / P.nsets;
    // Warps of a set finish a stage together only if they carry similar work: while slots are free, the heaviest
    // two-row unit above the per-warp mean is split into its two rows (a 2 x nc rectangle into two 1 x nc ones; the
    // triangles of a diagonal block into their rows).  Costs one more fragment load per DMMA in the split units.
    {
        static const int split_env = [] {
            const char* e = std::getenv("FRK_GRAM_MID_SPLIT");   // tuning aid: threshold in percent of the mean; 0 = never split
            return e ? std::atoi(e) : 115;
        }();
        int total = 0;
        for (const U& x : units) total += x.cost;
        const double mean = double(total) / wps;
        const int split_pct = (P.nsets == 1) ? std::max(split_env, 160) : split_env;   // one set: all 16 warps share every stage anyway
        while (split_env && int(units.size()) < wps * MNU) {
            int h = -1;
            for (int i = 0; i < int(units.size()); ++i) {
                const int kind = units[i].u.kind;
                const bool two_rows = kind == GRAM_MID_TRI24 || kind == GRAM_MID_TRI22 || (kind >= 5 && kind <= 8);
                if (two_rows && units[i].cost * 100.0 > split_pct * mean && (h < 0 || units[i].cost > units[h].cost)) h = i;
            }
            if (h < 0) break;
            const U x = units[h];
            units.erase(units.begin() + h);
            if (x.u.kind == GRAM_MID_TRI24) {
                add(x.ga0, false, x.gb0, 1, 4, 0);
                add(x.ga0 + 1, false, x.gb0 + 1, 1, 3, 0);
            } else if (x.u.kind == GRAM_MID_TRI22) {
                add(x.ga0, false, x.gb0, 1, 2, 0);
                add(x.ga0 + 1, false, x.gb0 + 1, 1, 1, 0);
            } else {
                const int nc = (x.u.kind - 1) % 4 + 1;
                add(x.ga0, x.isz, x.gb0, 1, nc, 0);
                add(x.ga0 + 1, x.isz, x.gb0, 1, nc, 0);
            }
        }
    }
    const int nunits = int(units.size());
    // bins whose loads must be equal: with s sets a logical sub-partition p shares the pipe with p + k*4/s
    const int nbins = 4 / P.nsets;
    std::stable_sort(units.begin(), units.end(), [](const U& x, const U& y) { return x.cost > y.cost; });
    // slot table: warp wv holds units slot_unit[wv][0..MNU-1] (-1 = empty).  Longest first onto the least loaded bin,
    // then pairwise swaps / moves between bins while they lower the spread (longest-first alone leaves e.g. 29|26 at K = 73).
    std::vector<int> slot_unit(size_t(wps) * MNU, -1), wload(wps, 0), bin(std::max(1, nbins), 0);
    auto cost_of = [&](int ui) { return ui < 0 ? 0 : units[ui].cost; };
    for (int ui = 0; ui < nunits; ++ui) {
        int best = -1, best_slot = -1;
        for (int wv = 0; wv < wps; ++wv) {
            int free_slot = -1;
            for (int sl = 0; sl < MNU; ++sl)
                if (slot_unit[wv * MNU + sl] < 0) { free_slot = sl; break; }
            if (free_slot < 0) continue;
            const bool better = best < 0 || bin[wv % nbins] < bin[best % nbins] ||
                                (bin[wv % nbins] == bin[best % nbins] && wload[wv] < wload[best]);
            if (better) { best = wv; best_slot = free_slot; }
        }
        if (best < 0) return false;
        slot_unit[best * MNU + best_slot] = ui;
        wload[best] += units[ui].cost;
        bin[best % nbins] += units[ui].cost;
    }
    for (bool improved = true; improved && nbins > 1;) {   // every swap lowers the sum of squared bin loads: terminates
        improved = false;
        int bx = -1, by = -1, bgain = 0;
        for (int x = 0; x < wps * MNU; ++x)
            for (int y = 0; y < wps * MNU; ++y) {
                const int px = (x / MNU) % nbins, py = (y / MNU) % nbins;
                const int dlt = cost_of(slot_unit[x]) - cost_of(slot_unit[y]);   // load moved from bin px to bin py
                if (px == py || dlt <= 0) continue;
                const int gain = 2 * dlt * (bin[px] - bin[py] - dlt);            // decrease of bin[px]^2 + bin[py]^2
                if (gain > bgain) { bgain = gain; bx = x; by = y; }
            }
        if (bx >= 0) {
            const int dlt = cost_of(slot_unit[bx]) - cost_of(slot_unit[by]);
            std::swap(slot_unit[bx], slot_unit[by]);
            wload[bx / MNU] -= dlt; wload[by / MNU] += dlt;
            bin[(bx / MNU) % nbins] -= dlt; bin[(by / MNU) % nbins] += dlt;
            improved = true;
        }
    }
    P.kb = kb;
    P.nblocks = kb + zb;
    for (int wv = 0; wv < wps; ++wv) {
        // a warp's units are packed to the front (the kernel tests the last slot to know how many it holds)
        int nk = 0;
        for (int sl = 0; sl < MNU; ++sl) {
            const int ui = slot_unit[wv * MNU + sl];
            if (ui < 0) continue;
            const U& x = units[ui];
            P.units[wv * MNU + nk] = x.u;
            const int kind = x.u.kind;
            const bool tri = kind == GRAM_MID_TRI24 || kind == GRAM_MID_TRI22;
            const int nr = tri ? 2 : (kind - 1) / 4 + 1;
            const int nc = kind == GRAM_MID_TRI24 ? 4 : (kind == GRAM_MID_TRI22 ? 2 : (kind - 1) % 4 + 1);
            for (int a = 0; a < nr; ++a)
                for (int b = 0; b < nc; ++b) {
                    if (tri && a == 1 && b == 0) continue;
                    GramMidTile tl;
                    tl.ga = short(x.ga0 + a);
                    tl.gb = short(x.gb0 + b);
                    tl.isz = x.isz ? 1 : 0;
                    tl.warp = (unsigned char)wv;
                    tl.slot = (unsigned char)(nk * 8 + a * 4 + b);
                    tl.pad = 0;
                    P.tiles.push_back(tl);
                }
            ++nk;
        }
    }
    P.nunits = nunits;
    // per sub-partition DMMAs per k4-step (over one round of all sets), for the tests and the probe
    for (int p = 0; p < 4; ++p) P.sp_load[p] = 0;
    const int rot = (P.nsets == 4) ? 1 : (P.nsets == 2 ? 2 : 0);
    for (int s = 0; s < P.nsets; ++s)
        for (int j = 0; j < wps; ++j) P.sp_load[(s * wps + (j + s * rot) % wps) % 4] += wload[j];
    return true;
}

bool gram_stats_mid_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                        const double* dw, double* dpacked, GramWorkspace& ws, int sms, cudaStream_t st) {
    if (!gram_tma_usable(dF, ldf, dZ, T, ldz, n)) return false;
    if (ws.mid_K != K || ws.mid_T != T) {
        GramMidPlan P;
        ws.mid_ok = gram_mid_plan(K, T, P);
        ws.mid_K = K;
        ws.mid_T = T;
        if (ws.mid_ok) {
            const size_t ub = sizeof(GramMidUnit) * MNW * MNU, tb = sizeof(GramMidTile) * P.tiles.size();
            ws.guard.quiesce();
            ws.mid_tab.alloc((ub + tb + 7) / 8);
            FRK_CUDA(cudaMemcpyAsync(ws.mid_tab.p, P.units, ub, cudaMemcpyHostToDevice, st));
            FRK_CUDA(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(ws.mid_tab.p) + ub, P.tiles.data(), tb, cudaMemcpyHostToDevice, st));
            FRK_CUDA(cudaStreamSynchronize(st));   // host temporaries; only on a shape change
            ws.mid_ntiles = int(P.tiles.size());
            ws.mid_nsets = P.nsets;
            ws.mid_kb = P.kb;
            ws.mid_nblocks = P.nblocks;
            ws.mid_stage_bytes = P.stage_bytes;
            for (int b = 0; b < 8; ++b) { ws.mid_blk_off[b] = P.blk_off[b]; ws.mid_blk_cols[b] = P.blk_cols[b]; }
        }
    }
    if (!ws.mid_ok) return false;
    int dev = 0;
    FRK_CUDA(cudaGetDevice(&dev));
    static std::once_flag once[64];   // the opt-in to large dynamic shared memory is per device
    std::call_once(once[dev & 63], [] {
        const int cap = int(MID_SMEM_BUDGET + 2048);
        FRK_CUDA(cudaFuncSetAttribute(gram_mid_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
        FRK_CUDA(cudaFuncSetAttribute(gram_mid_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
    });
    const GramMidTile* dtiles = reinterpret_cast<const GramMidTile*>(reinterpret_cast<const unsigned char*>(ws.mid_tab.p) +
                                                                    sizeof(GramMidUnit) * MNW * MNU);
    MidArgs a;
    a.K = K; a.T = T; a.n = n; a.w = dw;
    a.kb = ws.mid_kb;
    a.nblocks = ws.mid_nblocks;
    a.nsets = ws.mid_nsets;
    a.ntiles = ws.mid_ntiles;
    const int stage_bytes = ws.mid_stage_bytes;
    a.stage_bytes = stage_bytes;
    a.nst = int(std::min<size_t>(MAXST, MID_SMEM_BUDGET / stage_bytes));
    a.nst -= a.nst % a.nsets;   // a ring slot is always consumed by the same set: its warps see the slot's phases in order
    const size_t smem = size_t(a.nst) * stage_bytes + 1024 + MAXST * 8 + (MAXST + 1) * 4 + MNW * 8 + 64;
    // one CTA per SM; at least 8 stages of rows per CTA so that small inputs do not pay 148 pipeline fills
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>(sms, (n + 8 * MKR - 1) / (8 * MKR))));
    a.rows_cta = ((n + grid - 1) / grid + MKR - 1) / MKR * MKR;
    a.units = reinterpret_cast<const GramMidUnit*>(ws.mid_tab.p);
    const size_t need = size_t(grid) * MNW * MNU * 8 * 64 + size_t(grid);
    ws.guard.grow(ws.partial, need);
    a.partial = ws.partial.p;
    a.zzcta = ws.partial.p + size_t(grid) * MNW * MNU * 8 * 64;

    // tensor maps: full-width boxes of F / Z and, where the last block is narrower, a box of exactly its width
    CUtensorMap maps[4];
    int nmaps = 0, key[4] = {-1, -1, -1, -1};
    auto map_for = [&](bool isz, int width) {
        const int k = (isz ? 1000 : 0) + width;
        for (int i = 0; i < nmaps; ++i)
            if (key[i] == k) return i;
        maps[nmaps] = isz ? mid_make_map(dZ, n, T, ldz, width) : mid_make_map(dF, n, K, ldf, width);
        key[nmaps] = k;
        return nmaps++;
    };
    for (int b = 0; b < a.nblocks; ++b) {
        const bool isz = b >= a.kb;
        a.blk_off[b] = ws.mid_blk_off[b];
        a.blk_col[b] = (isz ? b - a.kb : b) * MCB;
        a.blk_map[b] = map_for(isz, ws.mid_blk_cols[b]);
    }
    for (int i = nmaps; i < 4; ++i) maps[i] = maps[0];
    if (dw) gram_mid_kernel<true><<<grid, MNT, smem, st>>>(maps[0], maps[1], maps[2], maps[3], a);
    else gram_mid_kernel<false><<<grid, MNT, smem, st>>>(maps[0], maps[1], maps[2], maps[3], a);
    FRK_CUDA(cudaGetLastError());
    gram_mid_reduce_kernel<<<a.ntiles, 256, 0, st>>>(a, dtiles, grid, dpacked);
    FRK_CUDA(cudaGetLastError());
    ws.last_nsplit = grid;
    ws.last_npairs = 1;
    return true;
}

}  // namespace frk
