// K4, TMA variant: G = F' W F, C = F' W Z over 128 x 128 output tiles, operands staged by 2-D TMA tensor copies.
//
// Same mathematics and the same workspace / reduction as gram.cu (which stays as the path for odd leading
// dimensions); what changes is the machinery around the DMMAs, applying what the K1 profiles taught:
//   * panels arrive by cp.async.bulk.tensor.2d (UTMALDG) with SWIZZLE_128B into dense [column][16 rows] boxes;
//     the MMA k index is permuted (rows 2q, 2q+1, 2q+8, 2q+9 form k4-step q) so that every fragment LDS.64 is
//     bank-conflict free under that swizzle; out-of-range rows / columns are zero-filled by the TMA unit;
//   * no CTA-wide barrier inside a work item: a 6-deep stage ring, `full` mbarriers, and the warp that is last
//     to finish with a stage (shared counter) re-arms it;
//   * persistent CTAs pull (row split, tile pair) items from a global counter, split-major so that CTAs running
//     together read the same rows of F through L2;
//   * diagonal tile pairs run only their 10 upper-triangular 32 x 32 warp tiles, on warps 0..9 (3,3,2,2 per SM
//     sub-partition), so a diagonal item costs 3/4 of an off-diagonal one instead of the same.
#include <algorithm>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include "frk_common.cuh"
#include "gram.cuh"

namespace frk {

namespace {

constexpr int TB = 128;
constexpr int KR = 16;
constexpr int NST = 6;
constexpr int NTHREADS = 512;
constexpr int NWARPS = NTHREADS / 32;
constexpr int PANEL_BYTES = TB * KR * 8;          // 16 KB, 1024-byte aligned (swizzle atom)
constexpr int STAGE_BYTES = 2 * PANEL_BYTES;
constexpr size_t SMEM_BYTES = size_t(NST) * STAGE_BYTES + 1024 /* alignment slack */ + 256;

struct GramTmaArgs {
    int K, T;
    int64_t n;
    const double* w;
    int ntf, ntz, npairs, nsplit;
    int64_t rows_per_split;
    double* partial;         // [nsplit][npairs][TB*TB]
    unsigned* counter;       // work-item counter (zeroed before the launch)
    const int* pair_order;   // pair ids, long items first
    int nitems;
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

// upper-triangular warp tiles of a diagonal 4 x 4 warp grid, in the order warps 0..9 take them
__device__ __constant__ int kDiagWr[10] = {0, 0, 0, 0, 1, 1, 1, 2, 2, 3};
__device__ __constant__ int kDiagWc[10] = {0, 1, 2, 3, 1, 2, 3, 2, 3, 3};

template <bool WEIGHTED>
__global__ void __launch_bounds__(NTHREADS, 1)
gram_tma_kernel(const __grid_constant__ CUtensorMap tmF, const __grid_constant__ CUtensorMap tmZ, const GramTmaArgs a) {
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
    const int nff = a.ntf * (a.ntf + 1) / 2;

    for (;;) {
        __syncthreads();  // every warp is done with the previous item (ring quiescent, done[] all zero)
        if (tid == 0) *s_item = int(atomicAdd(a.counter, 1u));
        __syncthreads();
        const int item = *s_item;
        if (item >= a.nitems) break;
        const int split = item / a.npairs;
        const int pair = a.pair_order[item % a.npairs];
        int ta = 0, tb = 0;
        bool bz = false;
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
        const bool diag = (!bz && ta == tb);
        const int nactive = diag ? 10 : NWARPS;
        const bool active = warp < nactive;
        const int wr = diag ? kDiagWr[active ? warp : 0] : (warp >> 2);
        const int wc = diag ? kDiagWc[active ? warp : 0] : (warp & 3);
        const CUtensorMap* tmB = bz ? &tmZ : &tmF;
        const uint32_t stage_tx = diag ? PANEL_BYTES : 2 * PANEL_BYTES;

        const int64_t r_begin = int64_t(split) * a.rows_per_split;
        const int64_t r_end = min(a.n, r_begin + a.rows_per_split);
        const int nsteps = (r_end > r_begin) ? int((r_end - r_begin + KR - 1) / KR) : 0;

        auto issue = [&](int stage, int step) {  // one thread
            unsigned char* As = base + size_t(stage) * STAGE_BYTES;
            const int r0 = int(r_begin + int64_t(step) * KR);
            mbar_expect_tx(&full[stage], stage_tx);
            tma_load_2d(As, &tmF, r0, ta * TB, &full[stage]);
            if (!diag) tma_load_2d(As + PANEL_BYTES, tmB, r0, tb * TB, &full[stage]);
        };
        if (tid == 0)
            for (int i = 0; i < NST && i < nsteps; ++i) issue((s + i) % NST, i);

        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        if (active) {
            const int a_line = (wr * 32 + g) * 128;
            const int b_line = (wc * 32 + g) * 128;
            double wq[4] = {1.0, 1.0, 1.0, 1.0};
            auto load_w = [&](int step) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int64_t r = r_begin + int64_t(step) * KR + 2 * q + (t & 1) + 8 * (t >> 1);
                    wq[q] = (r < r_end) ? __ldg(a.w + r) : 0.0;
                }
            };
            if (WEIGHTED && nsteps > 0) load_w(0);
            bool ok = false;
            int ls = s;
            uint32_t lpar = par;
            for (int step = 0; step < nsteps; ++step) {
                if (!ok) mbar_wait(&full[ls], lpar);
                const unsigned char* As = base + size_t(ls) * STAGE_BYTES;
                const unsigned char* Bs = diag ? As : As + PANEL_BYTES;
                const int lsn = (ls + 1 == NST) ? 0 : ls + 1;
                const uint32_t lparn = (ls + 1 == NST) ? (lpar ^ 1u) : lpar;
                double wcur[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) wcur[q] = wq[q];
                if (WEIGHTED && step + 1 < nsteps) load_w(step + 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double af[4], bf[4];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) af[mt] = *reinterpret_cast<const double*>(As + a_line + mt * 1024 + off[q]);
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) bf[nt] = *reinterpret_cast<const double*>(Bs + b_line + nt * 1024 + off[q]);
                    if (WEIGHTED) {
#pragma unroll
                        for (int mt = 0; mt < 4; ++mt) af[mt] *= wcur[q];
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) dmma_nv(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
                    if (q == 3 && step + 1 < nsteps) ok = mbar_test_wait(&full[lsn], lparn);
                }
                __syncwarp();
                if (lane == 0) {
                    if (smem_add_relaxed(&done[ls], 1u) == uint32_t(nactive - 1)) {
                        done[ls] = 0;
                        if (step + NST < nsteps) issue(ls, step + NST);
                    }
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
    cuuint32_t box[2] = {KR, TB};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) fail(3, "cuTensorMapEncodeTiled failed (%d)", int(r));
    return tm;
}

}  // namespace

bool gram_tma_usable(const double* dF, int64_t ldf, const double* dZ, int T, int64_t ldz, int64_t n) {
    if (reinterpret_cast<uintptr_t>(dF) % 16 || ldf % 2) return false;
    if (T > 0 && (reinterpret_cast<uintptr_t>(dZ) % 16 || ldz % 2)) return false;
    if (n >= (int64_t(1) << 31)) return false;  // TMA coordinates are 32-bit
    return true;
}

void gram_stats_tma_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                        const double* dw, GramWorkspace& ws, int sms, GramLayout& lay, cudaStream_t st) {
    GramTmaArgs a;
    a.K = K; a.T = T; a.n = n; a.w = dw;
    a.ntf = (K + TB - 1) / TB;
    a.ntz = (T + TB - 1) / TB;
    a.npairs = a.ntf * (a.ntf + 1) / 2 + a.ntf * a.ntz;
    // about six items per CTA for dynamic balance, at least 2048 rows per item
    const int64_t max_split = std::max<int64_t>(1, n / 2048);
    a.nsplit = int(std::max<int64_t>(1, std::min<int64_t>(max_split, (int64_t(sms) * 6 + a.npairs - 1) / a.npairs)));
    a.rows_per_split = ((n + a.nsplit - 1) / a.nsplit + KR - 1) / KR * KR;
    a.nitems = a.nsplit * a.npairs;
    const size_t need = size_t(a.nitems) * TB * TB;
    if (ws.partial.n < need) ws.partial.alloc(need);
    a.partial = ws.partial.p;
    // pair order: off-diagonal F x F, then F x Z, then the (cheaper) diagonal pairs; cached per tile shape
    if (ws.order_ntf != a.ntf || ws.order_ntz != a.ntz) {
        std::vector<int> order, diag_ids;
        const int nff = a.ntf * (a.ntf + 1) / 2;
        int id = 0;
        for (int ta = 0; ta < a.ntf; ++ta)
            for (int tb = ta; tb < a.ntf; ++tb, ++id) (ta == tb ? diag_ids : order).push_back(id);
        for (id = nff; id < a.npairs; ++id) order.push_back(id);
        order.insert(order.end(), diag_ids.begin(), diag_ids.end());
        ws.pair_order.alloc(order.size() + 1);
        FRK_CUDA(cudaMemcpyAsync(ws.pair_order.p, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice, st));
        FRK_CUDA(cudaStreamSynchronize(st));  // `order` is a host temporary; only on a shape change
        ws.order_ntf = a.ntf;
        ws.order_ntz = a.ntz;
    }
    unsigned* counter = reinterpret_cast<unsigned*>(ws.pair_order.p + a.npairs);
    FRK_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
    a.pair_order = ws.pair_order.p;
    a.counter = counter;

    const CUtensorMap tmF = make_map(dF, n, K, ldf);
    const CUtensorMap tmZ = (T > 0) ? make_map(dZ, n, T, ldz) : tmF;
    static bool attr_set = false;
    if (!attr_set) {
        FRK_CUDA(cudaFuncSetAttribute(gram_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
        FRK_CUDA(cudaFuncSetAttribute(gram_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(SMEM_BYTES)));
        attr_set = true;
    }
    const int grid = std::min(sms, a.nitems);
    if (dw) gram_tma_kernel<true><<<grid, NTHREADS, SMEM_BYTES, st>>>(tmF, tmZ, a);
    else gram_tma_kernel<false><<<grid, NTHREADS, SMEM_BYTES, st>>>(tmF, tmZ, a);
    FRK_CUDA(cudaGetLastError());
    lay.ntf = a.ntf; lay.ntz = a.ntz; lay.npairs = a.npairs; lay.nsplit = a.nsplit;
}

}  // namespace frk
