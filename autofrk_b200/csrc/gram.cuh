// K4: Gram / cross-product / sum-of-squares reductions (see gram.cu).
#pragma once
#include <utility>
#include <vector>
#include "frk_common.cuh"

namespace frk {

// TMA path: one job = the 32-column blocks it stages and the block pairs (a, b) it computes (gram_tma.cu)
struct GramJobHost {
    std::vector<int> blocks;                 // F blocks are < ceil(K/32); Z blocks follow
    std::vector<std::pair<int, int>> tiles;
};
std::vector<GramJobHost> gram_pack_jobs(int K, int T);

struct GramWorkspace {
    DevBuf<double> partial;  // row-split partial tiles
    DevBuf<double> zzpart;   // per-block partial sums of squares
    DevBuf<int> pair_order;  // TMA path: job table + work counter
    int jobs_K = -1, jobs_T = -1, njobs = 0, job_tiles = 0;  // shape the cached job table was built for
    int64_t launches = 0;
    int last_nsplit = 0, last_npairs = 0;
};

// number of doubles in the packed result [G (K x K, column-major, full symmetric) | C (K x T) | zz]
size_t gram_packed_len(int K, int T);

// dF n x K (ld ldf), dZ n x T (ld ldz; T may be 0) on the device -> dpacked on the device; asynchronous on st.
// dw: optional row weights (n doubles, device) or nullptr.
void gram_stats_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                    const double* dw, double* dpacked, GramWorkspace& ws, cudaStream_t st);

// TMA variant (gram_tma.cu): needs 16-byte aligned bases and even leading dimensions
bool gram_tma_usable(const double* dF, int64_t ldf, const double* dZ, int T, int64_t ldz, int64_t n);
void gram_stats_tma_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                        const double* dw, double* dpacked, GramWorkspace& ws, int sms, int nzz, cudaStream_t st);

// narrow variant (gram_narrow.cu) for K <= 48, T <= 16: returns false if it does not apply
bool gram_stats_narrow_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                           const double* dw, double* dpacked, GramWorkspace& ws, int sms, int nzz, cudaStream_t st);

}  // namespace frk
