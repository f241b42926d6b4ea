// K4, TMA variant: G = F' W F, C = F' W Z from 32 x 32 warp tiles packed into jobs, operands staged by 2-D TMA copies.
//
// Same mathematics as gram.cu (which stays as the path for odd leading dimensions); what changes is the machinery
// around the DMMAs, applying what the K1 profiles taught:
//   * the output is cut into 32-column blocks (F blocks, then Z blocks); the work is the list of block pairs
//     (a <= b) of G plus the F x Z pairs, one 32 x 32 register tile per warp.  A job is up to 16 such tiles that
//     together touch at most 8 column blocks ("slots"); the host packs the pairs into jobs (4 x 4 macro tiles,
//     then the under-filled ones merged: a diagonal macro tile's 10 upper-triangular tiles share a job with the
//     F x Z tiles of the same blocks), so no DMMA is spent on a lower triangle or on padding beyond 31 columns.
//     The 128-column tiling this replaces spent 200 warp tiles on K = 500, T = 1 (now 152) and 68 on K = 200,
//     T = 20 (now 35);
//   * inside a job the unit of work is (tile, set of k4-steps): tiles are dealt to the four SM sub-partitions
//     heaviest first, and while warps are free part of a tile's k4-steps moves from the busiest sub-partition to
//     the idlest one (a 10-tile diagonal job runs as 2.5 tile-times per sub-partition instead of 3,3,2,2; a job
//     with few tiles still occupies all four); the reduction adds a split tile's partials like further row splits;
//   * slots arrive by cp.async.bulk.tensor.2d (UTMALDG) with SWIZZLE_128B into dense [column][16 rows] boxes;
//     the MMA k index is permuted (rows 2q, 2q+1, 2q+8, 2q+9 form k4-step q) so that every fragment LDS.64 is
//     bank-conflict free under that swizzle; out-of-range rows / columns are zero-filled by the TMA unit;
//   * no CTA-wide barrier inside a work item: a 6-deep stage ring, `full` mbarriers, and the warp that is last
//     to finish with a stage (shared counter) re-arms it, one TMA per lane;
//   * persistent CTAs pull (row split, job) items from a global counter, split-major so that CTAs running
//     together read the same rows of F through L2 (items of a few thousand rows when there are many jobs, so that
//     they stay together); jobs are ordered longest first;
//   * every item writes its tiles to a workspace and a second kernel adds them in a fixed order (bitwise
//     reproducible, no atomics) and mirrors G.
#include <algorithm>
#include <cstdlib>
#include <iterator>
#include <mutex>
#include <type_traits>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "frk_common.cuh"
#include "gram.cuh"

namespace frk {

namespace {

constexpr int CB = 32;                            // columns per block = warp tile edge
// Rows per pipeline stage: 16 (six 32 KB stages) for shapes whose jobs stage many column blocks, 32 (three 64 KB
// stages, half the per-stage synchronisation) when all column blocks fit one job's slots (K + T small): measured
// n = 5e6: K = 64, T = 4 1.66 -> 1.35 ms and K = 200, T = 20 9.07 -> 8.72 ms with 32, but K = 500, T = 100 56.9 -> 58.8 ms.
constexpr int BOXR = 16;                          // rows per TMA box: 16 doubles = the 128-byte swizzle span
__host__ __device__ constexpr int stages_for(int kr) { return kr == 16 ? 6 : 3; }
constexpr int NTHREADS = 512;
constexpr int NWARPS = NTHREADS / 32;
constexpr int MAXSLOTS = 8;
constexpr int HALF_BYTES = CB * BOXR * 8;         // one TMA box: 4 KB, 1024-byte aligned (swizzle atom)
constexpr size_t smem_bytes_for(int kr) {
    return size_t(stages_for(kr)) * MAXSLOTS * (CB * kr * 8) + 1024 /* alignment slack */ + 256;
}
constexpr int TILE_DOUBLES = CB * CB;

// one job: column blocks to stage (F blocks are < kb, Z blocks follow) and the block pairs computed from them
struct GramJob {
    int nslots, ntiles, nactive;     // staged blocks, output tiles, warps with work
    int slot_block[MAXSLOTS];
    // per warp: the tile it works on, the k4-steps of every stage it takes (bit q of qmask; 0 = idle warp),
    // and how many 8-column groups of block b hold real columns (1..4)
    unsigned char w_tile[NWARPS], qmask[NWARPS], ntn[NWARPS];   // ntn = 5: full diagonal block, upper 8 x 8 sub-tiles only
    // per tile: its two staged blocks and the warps whose partial sums make it up
    unsigned char slot_a[NWARPS], slot_b[NWARPS], tile_nw[NWARPS], tile_w[NWARPS][4];
};

struct GramTmaArgs {
    int K, T, kb;            // kb = number of 32-column blocks of F
    int kr;                  // rows per stage the job table was built for (16 or 32)
    int64_t n;
    const double* w;
    int njobs, nsplit, nitems;
    int nsplit_big;          // the first nsplit_big row splits hold rows_big rows each, the others rows_small
    int64_t rows_big, rows_small;
    double* partial;         // [nsplit][njobs][NWARPS][32*32], column-major tiles
    unsigned* counter;       // work-item counter (zeroed before the launch)
    const GramJob* jobs;     // longest first
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void dmma_nv(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool WEIGHTED, int KR>
__global__ void __launch_bounds__(NTHREADS, 1)
gram_tma_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmZ, const GramTmaArgs a) {
    constexpr int KQ = KR / 4;                        // k4-steps per stage
    constexpr int NST = stages_for(KR);
    constexpr int SLOT_BYTES = CB * KR * 8;           // KR / 16 boxes per staged column block
    constexpr int STAGE_BYTES = MAXSLOTS * SLOT_BYTES;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* const base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* const full = reinterpret_cast<uint64_t*>(base + size_t(NST) * STAGE_BYTES);
    uint32_t* const done = reinterpret_cast<uint32_t*>(full + NST);
    int* const s_item = reinterpret_cast<int*>(done + NST);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            done[s] = 0;
        }
        mbar_fence_init();
    }
    // byte offset of this lane's 8-byte element inside a 128-byte line for k4-step q (row 2q + (t&1) + 8(t>>1),
    // 16-byte chunk XOR-swizzled with line & 7 == g because every fragment starts at a multiple of 8 columns)
    int off[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) off[q] = (((q + 4 * (t >> 1)) ^ g) << 4) + (t & 1) * 8;

    int s = 0;           // ring position and parity carried across work items
    uint32_t par = 0;

    for (;;) {
        __syncthreads();  // every warp is done with the previous item (ring quiescent, done[] all zero)
        if (tid == 0) *s_item = int(atomicAdd(a.counter, 1u));
        __syncthreads();
        const int item = *s_item;
        if (item >= a.nitems) break;
        const int split = item / a.njobs;
        const int jid = item % a.njobs;
        const GramJob* const job = a.jobs + jid;
        const int nslots = __ldg(&job->nslots);
        const int nactive = __ldg(&job->nactive);
        const int qmask = __ldg(&job->qmask[warp]);
        const bool active = qmask != 0;
        const int tile = __ldg(&job->w_tile[warp]);
        const int sa = __ldg(&job->slot_a[tile]);
        const int sb = __ldg(&job->slot_b[tile]);
        const int ntn = __ldg(&job->ntn[warp]);
        // the slot this lane copies when its warp re-arms a stage
        const int my_slot = lane / (KR / BOXR);
        const int my_block = (my_slot < nslots) ? __ldg(&job->slot_block[my_slot]) : 0;
        const CUtensorMap* const sl_tm = (my_block < a.kb) ? &tmF : &tmZ;
        const int sl_col = ((my_block < a.kb) ? my_block : my_block - a.kb) * CB;
        const uint32_t stage_tx = uint32_t(nslots) * SLOT_BYTES;

        const int64_t r_begin = (split < a.nsplit_big) ? int64_t(split) * a.rows_big
                                                       : int64_t(a.nsplit_big) * a.rows_big + int64_t(split - a.nsplit_big) * a.rows_small;
        const int64_t r_end = min(a.n, r_begin + ((split < a.nsplit_big) ? a.rows_big : a.rows_small));
        const int nsteps = (r_end > r_begin) ? int((r_end - r_begin + KR - 1) / KR) : 0;

        auto issue = [&](int stage, int step) {  // one whole warp
            unsigned char* As = base + size_t(stage) * STAGE_BYTES;
            const int r0 = int(r_begin + int64_t(step) * KR);
            if (lane == 0) mbar_expect_tx(&full[stage], stage_tx);
            __syncwarp();
            constexpr int BOXES = KR / BOXR;   // lane -> (slot, 16-row box of the slot)
            const int sl = lane / BOXES, bx = lane % BOXES;
            if (sl < nslots)
                tma_load_2d(As + sl * SLOT_BYTES + bx * HALF_BYTES, sl_tm, r0 + bx * BOXR, sl_col, &full[stage]);
        };
        if (warp == 0)
            for (int i = 0; i < NST && i < nsteps; ++i) issue((s + i) % NST, i);

        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        if (active) {
            const int a_line = sa * SLOT_BYTES + g * 128;
            const int b_line = sb * SLOT_BYTES + g * 128;
            double wq[KQ];
#pragma unroll
            for (int q = 0; q < KQ; ++q) wq[q] = 1.0;
            auto load_w = [&](int step) {
#pragma unroll
                for (int q = 0; q < KQ; ++q) {
                    const int64_t r = r_begin + int64_t(step) * KR + BOXR * (q >> 2) + 2 * (q & 3) + (t & 1) + 8 * (t >> 1);
                    wq[q] = (r < r_end) ? __ldg(a.w + r) : 0.0;
                }
            };
            if (WEIGHTED && nsteps > 0) load_w(0);
            double wcur[KQ];
            // this warp's k4-steps of one stage on the first NTN 8-column groups of block b
            auto ksteps = [&](const unsigned char* As, auto ntn_c, auto diag_c) {
                constexpr int NTN = decltype(ntn_c)::value;
                constexpr bool DIAG = decltype(diag_c)::value;   // diagonal block of G: 8 x 8 sub-tiles with nt >= mt only
#pragma unroll
                for (int q = 0; q < KQ; ++q) {
                    if (!((qmask >> q) & 1)) continue;   // warp-uniform
                    double af[4], bf[NTN];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) af[mt] = *reinterpret_cast<const double*>(As + a_line + mt * 1024 + (q >> 2) * HALF_BYTES + off[q & 3]);
#pragma unroll
                    for (int nt = 0; nt < NTN; ++nt) bf[nt] = *reinterpret_cast<const double*>(As + b_line + nt * 1024 + (q >> 2) * HALF_BYTES + off[q & 3]);
                    if (WEIGHTED) {
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) af[mt] *= wcur[q];
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = DIAG ? mt : 0; nt < NTN; ++nt) dmma_nv(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
                }
            };
            bool ok = false;
            int ls = s;
            uint32_t lpar = par;
            for (int step = 0; step < nsteps; ++step) {
                if (!ok) mbar_wait(&full[ls], lpar);
                const unsigned char* As = base + size_t(ls) * STAGE_BYTES;
                const int lsn = (ls + 1 == NST) ? 0 : ls + 1;
                const uint32_t lparn = (ls + 1 == NST) ? (lpar ^ 1u) : lpar;
#pragma unroll
                for (int q = 0; q < KQ; ++q) wcur[q] = wq[q];
                if (WEIGHTED && step + 1 < nsteps) load_w(step + 1);
                switch (ntn) {  // warp-uniform: real branches, never predicated-off DMMAs
                    case 1: ksteps(As, std::integral_constant<int, 1>{}, std::false_type{}); break;
                    case 2: ksteps(As, std::integral_constant<int, 2>{}, std::false_type{}); break;
                    case 3: ksteps(As, std::integral_constant<int, 3>{}, std::false_type{}); break;
                    case 5: ksteps(As, std::integral_constant<int, 4>{}, std::true_type{}); break;   // diagonal block
                    default: ksteps(As, std::integral_constant<int, 4>{}, std::false_type{}); break;
                }
                if (step + 1 < nsteps) ok = mbar_test_wait(&full[lsn], lparn);
                __syncwarp();
                uint32_t last = 0;
                if (lane == 0) last = (smem_add_relaxed(&done[ls], 1u) == uint32_t(nactive - 1)) ? 1u : 0u;
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    if (lane == 0) done[ls] = 0;
                    if (step + NST < nsteps) issue(ls, step + NST);
                }
                ls = lsn;
                lpar = lparn;
            }
        }
        // all warps advance the ring identically
        {
            const int adv = nsteps % NST;
            const int wraps = nsteps / NST;
            int ns = s + adv;
            uint32_t np = par ^ (uint32_t(wraps) & 1u);
            if (ns >= NST) { ns -= NST; np ^= 1u; }
            s = ns;
            par = np;
        }
        if (active) {
            double* out = a.partial + ((size_t(split) * a.njobs + jid) * NWARPS + warp) * TILE_DOUBLES;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int row = mt * 8 + g;
                        const int col = nt * 8 + 2 * t + j;
                        out[col * CB + row] = acc[mt][nt][j];
                    }
        }
    }
}

// add the partial tiles in a fixed order and scatter into packed = [G (K x K) | C (K x T) | zz].
// One block per (job, tile, 64-element slice); its four thread groups each add every fourth (row split, warp)
// partial in order and the four sums are combined as (s0 + s1) + (s2 + s3): the same order on every run.
constexpr int RED_ELEMS = 64, RED_GROUPS = 4;
__global__ void __launch_bounds__(RED_ELEMS * RED_GROUPS)
gram_tma_reduce_kernel(const GramTmaArgs a, const double* __restrict__ zzpart, int nzz, double* __restrict__ packed) {
    constexpr int SLICES = TILE_DOUBLES / RED_ELEMS;
    const int slice = blockIdx.x % SLICES;
    const int tile = (blockIdx.x / SLICES) % NWARPS;
    const int jid = blockIdx.x / (SLICES * NWARPS);
    const GramJob* const job = a.jobs + jid;
    const int ntiles = job->ntiles;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < nzz; ++i) s += zzpart[i];
        packed[size_t(a.K) * a.K + size_t(a.K) * a.T] = s;
    }
    if (tile >= ntiles) return;
    const int ba = job->slot_block[job->slot_a[tile]], bb = job->slot_block[job->slot_b[tile]];
    const bool bz = bb >= a.kb;
    const int e = slice * RED_ELEMS + (threadIdx.x % RED_ELEMS), grp = threadIdx.x / RED_ELEMS;
    const int row = e % CB, col = e / CB;
    const int ga = ba * CB + row;
    const int gb = (bz ? bb - a.kb : bb) * CB + col;
    const bool wanted = ga < a.K && gb < (bz ? a.T : a.K) && !(!bz && ba == bb && row > col);  // lower triangle of a
                                                                                               // diagonal block is mirrored
    __shared__ double sh[RED_GROUPS][RED_ELEMS];
    double s = 0.0;
    if (wanted) {
        const int nw = job->tile_nw[tile];
        const int nparts = a.nsplit * nw;
        for (int p = grp; p < nparts; p += RED_GROUPS) {
            const int sp = p / nw, wv = job->tile_w[tile][p % nw];
            s += a.partial[((size_t(sp) * a.njobs + jid) * NWARPS + wv) * TILE_DOUBLES + e];
        }
    }
    sh[grp][threadIdx.x % RED_ELEMS] = s;
    __syncthreads();
    if (grp == 0 && wanted) {
        const int i = threadIdx.x;
        s = (sh[0][i] + sh[1][i]) + (sh[2][i] + sh[3][i]);
        double* G = packed;
        double* C = packed + size_t(a.K) * a.K;
        if (bz) {
            C[ga + size_t(gb) * a.K] = s;
        } else {
            G[ga + size_t(gb) * a.K] = s;
            G[gb + size_t(ga) * a.K] = s;
        }
    }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        FRK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p) fail(3, "cuTensorMapEncodeTiled is not available in this driver");
        fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

CUtensorMap make_map(const double* ptr, int64_t rows, int cols, int64_t ld) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {cuuint64_t(rows), cuuint64_t(cols)};
    cuuint64_t gstr[1] = {cuuint64_t(ld) * 8};
    cuuint32_t box[2] = {BOXR, CB};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(3, "cuTensorMapEncodeTiled failed (%d)", int(r));
    return tm;
}

// ---- host: pack the block pairs into jobs -------------------------------------------------------------------
struct Piece {
    std::vector<std::pair<int, int>> tiles;  // (block a, block b) in the combined F|Z block space
    std::vector<int> blocks;                 // distinct blocks touched, ascending
};

int union_size(const std::vector<int>& x, const std::vector<int>& y) {
    std::vector<int> u;
    std::set_union(x.begin(), x.end(), y.begin(), y.end(), std::back_inserter(u));
    return int(u.size());
}

}  // namespace

std::vector<GramJobHost> gram_pack_jobs(int K, int T) {
    const int kb = (K + CB - 1) / CB, zb = (T + CB - 1) / CB;
    const int ng = (kb + 3) / 4, nzg = (zb + 3) / 4;
    auto group = [](int gidx, int nblocks, int base) {
        std::vector<int> b;
        for (int i = 4 * gidx; i < std::min(nblocks, 4 * gidx + 4); ++i) b.push_back(base + i);
        return b;
    };
    std::vector<Piece> pieces;
    for (int ga = 0; ga < ng; ++ga) {
        const std::vector<int> A = group(ga, kb, 0);
        Piece d;  // diagonal macro tile: upper triangle only
        for (size_t i = 0; i < A.size(); ++i)
            for (size_t j = i; j < A.size(); ++j) d.tiles.push_back({A[i], A[j]});
        d.blocks = A;
        pieces.push_back(d);
        for (int gb = ga + 1; gb < ng; ++gb) {
            const std::vector<int> B = group(gb, kb, 0);
            Piece o;
            for (int x : A)
                for (int y : B) o.tiles.push_back({x, y});
            o.blocks = A;
            o.blocks.insert(o.blocks.end(), B.begin(), B.end());
            pieces.push_back(o);
        }
        for (int gz = 0; gz < nzg; ++gz) {
            const std::vector<int> Z = group(gz, zb, kb);
            Piece o;
            for (int y : Z)
                for (int x : A) o.tiles.push_back({x, y});
            o.blocks = A;
            o.blocks.insert(o.blocks.end(), Z.begin(), Z.end());
            pieces.push_back(o);
        }
    }
    std::stable_sort(pieces.begin(), pieces.end(), [](const Piece& x, const Piece& y) { return x.tiles.size() > y.tiles.size(); });
    // first-fit-decreasing merge, preferring the job that shares the most blocks
    std::vector<Piece> jobs;
    for (const Piece& p : pieces) {
        int best = -1, best_new = 1 << 30;
        for (size_t j = 0; j < jobs.size(); ++j) {
            if (jobs[j].tiles.size() + p.tiles.size() > size_t(NWARPS)) continue;
            const int u = union_size(jobs[j].blocks, p.blocks);
            if (u > MAXSLOTS) continue;
            const int added = u - int(jobs[j].blocks.size());
            if (added < best_new) { best_new = added; best = int(j); }
        }
        if (best < 0) {
            jobs.push_back(p);
        } else {
            Piece& q = jobs[best];
            q.tiles.insert(q.tiles.end(), p.tiles.begin(), p.tiles.end());
            std::vector<int> u;
            std::set_union(q.blocks.begin(), q.blocks.end(), p.blocks.begin(), p.blocks.end(), std::back_inserter(u));
            q.blocks = u;
        }
    }
    std::stable_sort(jobs.begin(), jobs.end(), [](const Piece& x, const Piece& y) { return x.tiles.size() > y.tiles.size(); });
    std::vector<GramJobHost> out;
    for (const Piece& p : jobs) {
        GramJobHost h;
        h.blocks = p.blocks;
        h.tiles = p.tiles;
        out.push_back(h);
    }
    return out;
}

bool gram_tma_usable(const double* dF, int64_t ldf, const double* dZ, int T, int64_t ldz, int64_t n) {
    if (reinterpret_cast<uintptr_t>(dF) % 16 || ldf % 2) return false;
    if (T > 0 && (reinterpret_cast<uintptr_t>(dZ) % 16 || ldz % 2)) return false;
    if (n >= (int64_t(1) << 31)) return false;  // TMA coordinates are 32-bit
    return true;
}

void gram_stats_tma_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                        const double* dw, double* dpacked, GramWorkspace& ws, int sms, int nzz, cudaStream_t st) {
    GramTmaArgs a;
    a.K = K; a.T = T; a.n = n; a.w = dw;
    a.kb = (K + CB - 1) / CB;
    const int kr = (a.kb + (T + CB - 1) / CB <= MAXSLOTS) ? 32 : 16;   // see stages_for()
    const int kq = kr / 4;
    a.kr = kr;
    // job table: cached per shape
    if (ws.jobs_K != K || ws.jobs_T != T) {
        const std::vector<GramJobHost> hj = gram_pack_jobs(K, T);
        std::vector<GramJob> dj(hj.size());
        for (size_t j = 0; j < hj.size(); ++j) {
            GramJob& d = dj[j];
            const GramJobHost& h = hj[j];
            d = GramJob{};
            d.nslots = int(h.blocks.size());
            d.ntiles = int(h.tiles.size());
            for (int i = 0; i < d.nslots; ++i) d.slot_block[i] = h.blocks[i];
            // Work units: (tile, set of k4-steps).  Warp i runs on SM sub-partition i % 4, at most four warps each.
            // Cost of a unit = valid 8-column groups of block b x k4-steps.  Whole tiles go out heaviest first to
            // the least loaded sub-partition; then, while a warp is free, part of a unit moves from the busiest
            // sub-partition to the idlest one (a tile split over warps is added up by the reduction kernel), so
            // e.g. the 10 tiles of a diagonal macro tile cost 2.5 tile-times per sub-partition instead of 3,3,2,2.
            struct Unit { int tile, cost, mask, sp, ntn; };   // cost in 8 x 8 sub-tiles per k4-step (16 = full tile)
            std::vector<Unit> units;
            std::vector<int> order(d.ntiles);
            for (int i = 0; i < d.ntiles; ++i) {
                const int bb = h.tiles[i].second;
                const int valid = std::min(CB, bb < a.kb ? K - bb * CB : T - (bb - a.kb) * CB);
                const int ntn = (valid + 7) / 8;
                const bool diag = (h.tiles[i].first == bb) && ntn == 4;   // upper-triangular sub-tiles only (10 of 16)
                units.push_back({i, diag ? 10 : 4 * ntn, (1 << kq) - 1, -1, diag ? 5 : ntn});
                order[i] = i;
                d.slot_a[i] = (unsigned char)(std::find(h.blocks.begin(), h.blocks.end(), h.tiles[i].first) - h.blocks.begin());
                d.slot_b[i] = (unsigned char)(std::find(h.blocks.begin(), h.blocks.end(), h.tiles[i].second) - h.blocks.begin());
            }
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return units[x].cost > units[y].cost; });
            int load[4] = {0, 0, 0, 0}, used[4] = {0, 0, 0, 0};
            auto nq = [](int mask) { return __builtin_popcount(unsigned(mask)); };
            for (int i : order) {
                int best = -1;
                for (int sp = 0; sp < 4; ++sp)
                    if (used[sp] < 4 && (best < 0 || load[sp] < load[best])) best = sp;
                units[i].sp = best;
                ++used[best];
                load[best] += units[i].cost * kq;
            }
            auto split = [&](int u, int k, int to) {   // move k of unit u's k4-steps to a new unit on sub-partition `to`
                int moved = 0, keep = units[u].mask;
                for (int q = kq - 1; q >= 0 && nq(moved) < k; --q)
                    if ((keep >> q) & 1) { moved |= 1 << q; keep &= ~(1 << q); }
                units[u].mask = keep;
                load[units[u].sp] -= units[u].cost * k;
                units.push_back({units[u].tile, units[u].cost, moved, to, units[u].ntn});
                ++used[to];
                load[to] += units[u].cost * k;
            };
            auto tile_units = [&](int tile) {
                int c = 0;
                for (const Unit& u : units) c += (u.tile == tile);
                return c;
            };
            static const int split_max_tiles = [] {   // tuning aid: jobs with more tiles are not rebalanced
                const char* e = std::getenv("FRK_GRAM_SPLIT_MAXTILES");
                return e ? std::atoi(e) : NWARPS;
            }();
            for (;;) {
                if (d.ntiles > split_max_tiles) break;
                if (int(units.size()) >= NWARPS) break;
                int hi = 0, lo = -1;
                for (int sp = 0; sp < 4; ++sp) {
                    if (load[sp] > load[hi]) hi = sp;
                    if (used[sp] < 4 && (lo < 0 || load[sp] < load[lo])) lo = sp;
                }
                if (lo < 0 || lo == hi) break;
                int best_u = -1, best_k = 0, best_max = load[hi];
                for (int u = 0; u < int(units.size()); ++u) {
                    if (units[u].sp != hi || nq(units[u].mask) < 2 || tile_units(units[u].tile) >= 4) continue;
                    for (int k = 1; k < nq(units[u].mask); ++k) {
                        const int nm = std::max(load[hi] - units[u].cost * k, load[lo] + units[u].cost * k);
                        if (nm < best_max) { best_max = nm; best_u = u; best_k = k; }
                    }
                }
                if (best_u < 0) break;
                split(best_u, best_k, lo);
            }
            // few units: give every sub-partition a second warp to hide latency (halves on the same sub-partition)
            while (int(units.size()) * 2 <= NWARPS) {
                const int cnt = int(units.size());
                bool any = false;
                for (int u = 0; u < cnt; ++u)
                    if (nq(units[u].mask) >= 2 && used[units[u].sp] < 4 && tile_units(units[u].tile) < 4) {
                        split(u, nq(units[u].mask) / 2, units[u].sp);
                        any = true;
                    }
                if (!any) break;
            }
            int next[4] = {0, 1, 2, 3};
            d.nactive = int(units.size());
            for (const Unit& u : units) {
                const int wv = next[u.sp];
                next[u.sp] += 4;
                d.w_tile[wv] = (unsigned char)u.tile;
                d.qmask[wv] = (unsigned char)u.mask;
                d.ntn[wv] = (unsigned char)u.ntn;
                d.tile_w[u.tile][d.tile_nw[u.tile]++] = (unsigned char)wv;
            }
        }
        const size_t bytes = sizeof(GramJob) * dj.size() + sizeof(unsigned);
        ws.guard.quiesce();
        ws.pair_order.alloc((bytes + sizeof(int) - 1) / sizeof(int));
        FRK_CUDA(cudaMemcpyAsync(ws.pair_order.p, dj.data(), sizeof(GramJob) * dj.size(), cudaMemcpyHostToDevice, st));
        FRK_CUDA(cudaStreamSynchronize(st));  // `dj` is a host temporary; only on a shape change
        ws.jobs_K = K;
        ws.jobs_T = T;
        ws.njobs = int(dj.size());
        int64_t units = 0;
        for (const GramJob& d : dj) units += d.ntiles;
        ws.job_tiles = int(units);
    }
    a.njobs = ws.njobs;
    a.jobs = reinterpret_cast<const GramJob*>(ws.pair_order.p);
    a.counter = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(ws.pair_order.p) + sizeof(GramJob) * a.njobs);
    FRK_CUDA(cudaMemsetAsync(a.counter, 0, sizeof(unsigned), st));
    // Row splits.  Items are handed out dynamically, split-major; the tail of the launch costs up to one item of
    // idle time per CTA, so the first 70% of the rows are cut into about four large items per CTA and the rest into
    // about six items per CTA a quarter of that size (measured: 63.3 -> 59 ms at n = 5e6, K = 500, T = 100 against
    // six equal items per CTA).  A single job makes all items equal: one per CTA.  At least 2048 rows per item.
    static const int items_env = [] {
        const char* e = std::getenv("FRK_GRAM_ITEMS");
        return e ? std::atoi(e) : 0;
    }();
    auto round_rows = [kr](int64_t r) { return std::max<int64_t>(kr, (r + kr - 1) / kr * kr); };
    const int64_t max_split = std::max<int64_t>(1, n / 2048);
    // Many jobs over many rows (K = 500: 13 jobs re-reading the same 20 column blocks): short items keep the CTAs that
    // run together within a few thousand rows of each other, so the re-reads hit L2 instead of HBM -- at n = 5e6,
    // K = 500, T = 100 dram__bytes_read falls from 82.8 GB (3.45 x algorithmic) to 37.4 GB at the same 56 ms
    // (profiles/gram_items_probe_r02.txt); the price is 1.7 GB of partial tiles.  With few jobs there is little
    // to re-read and short items only add epilogues (K = 200, T = 20: +5%), so those keep the two-level split.
    const bool short_items = a.njobs >= 8 && n >= (int64_t(1) << 20);
    if (a.njobs == 1 || items_env > 0 || max_split < 16 || short_items) {
        const int per_cta = (a.njobs == 1) ? 1 : (items_env > 0 ? items_env : (short_items ? 96 : 6));
        a.nsplit = int(std::max<int64_t>(1, std::min<int64_t>(max_split, (int64_t(sms) * per_cta + a.njobs - 1) / a.njobs)));
        a.nsplit_big = a.nsplit;
        a.rows_big = a.rows_small = round_rows((n + a.nsplit - 1) / a.nsplit);
    } else {
        const int64_t want_big = std::max<int64_t>(1, (int64_t(sms) * 4 + a.njobs - 1) / a.njobs);
        const int64_t want_small = std::max<int64_t>(1, (int64_t(sms) * 6 + a.njobs - 1) / a.njobs);
        const int64_t n_big = n / 10 * 7;
        a.nsplit_big = int(std::min<int64_t>(want_big, std::max<int64_t>(1, n_big / 2048)));
        a.rows_big = round_rows((n_big + a.nsplit_big - 1) / a.nsplit_big);
        const int64_t n_small = std::max<int64_t>(0, n - int64_t(a.nsplit_big) * a.rows_big);
        const int nsmall = int(std::min<int64_t>(want_small, std::max<int64_t>(1, n_small / 2048)));
        a.rows_small = round_rows((n_small + nsmall - 1) / nsmall);
        a.nsplit = a.nsplit_big + (n_small > 0 ? int((n_small + a.rows_small - 1) / a.rows_small) : 0);
    }
    a.nitems = a.nsplit * a.njobs;
    const size_t need = size_t(a.nitems) * NWARPS * TILE_DOUBLES;
    ws.guard.grow(ws.partial, need);
    a.partial = ws.partial.p;

    const CUtensorMap tmF = make_map(dF, n, K, ldf);
    const CUtensorMap tmZ = (T > 0) ? make_map(dZ, n, T, ldz) : tmF;
    static std::once_flag attr_once[64];  // the opt-in to large dynamic shared memory is per device; thread-safe
    int dev = 0;
    FRK_CUDA(cudaGetDevice(&dev));
    std::call_once(attr_once[dev & 63], [] {
        FRK_CUDA(cudaFuncSetAttribute(gram_tma_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes_for(16))));
        FRK_CUDA(cudaFuncSetAttribute(gram_tma_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes_for(16))));
        FRK_CUDA(cudaFuncSetAttribute(gram_tma_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes_for(32))));
        FRK_CUDA(cudaFuncSetAttribute(gram_tma_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes_for(32))));
    });
    const int grid = std::min(sms, a.nitems);
    if (kr == 16) {
        if (dw) gram_tma_kernel<true, 16><<<grid, NTHREADS, smem_bytes_for(16), st>>>(tmF, tmZ, a);
        else gram_tma_kernel<false, 16><<<grid, NTHREADS, smem_bytes_for(16), st>>>(tmF, tmZ, a);
    } else {
        if (dw) gram_tma_kernel<true, 32><<<grid, NTHREADS, smem_bytes_for(32), st>>>(tmF, tmZ, a);
        else gram_tma_kernel<false, 32><<<grid, NTHREADS, smem_bytes_for(32), st>>>(tmF, tmZ, a);
    }
    FRK_CUDA(cudaGetLastError());
    gram_tma_reduce_kernel<<<a.njobs * NWARPS * (TILE_DOUBLES / RED_ELEMS), RED_ELEMS * RED_GROUPS, 0, st>>>(a, ws.zzpart.p, nzz,
                                                                                                          dpacked);
    FRK_CUDA(cudaGetLastError());
    ws.last_nsplit = a.nsplit;
    ws.last_npairs = a.njobs;
}

}  // namespace frk
