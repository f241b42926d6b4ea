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

// register-resident variant for K + T <= 256 columns (gram_mid.cu): 16 warps x 2 units of 2 x 4 tiles of 8 x 8
constexpr int GRAM_MID_WARPS = 16;
constexpr int GRAM_MID_UNITS = 2;
struct GramMidUnit {
    int a_base;        // byte offset inside a stage of the first of <= 2 adjacent row groups (8 columns of F each)
    int b_base;        // ... of the first of <= 4 adjacent column groups (inside one 32-column block of F or of Z)
    int ga0, gb0, isz; // the same two positions as 8-column group indices (gb0 in F, or in Z when isz)
    int kind;          // 0 = empty; 1 + (nr-1)*4 + (nc-1): nr x nc rectangle; 9: 2 x 4 minus tile (1,0) (upper rows of a
                       // diagonal block); 10: 2 x 2 minus tile (1,0) (its lower rows).  Tile (a, b) -> accumulator a*4 + b
};
constexpr int GRAM_MID_TRI24 = 9, GRAM_MID_TRI22 = 10;
struct GramMidTile {
    short ga, gb;      // 8-column group of F (rows of the output) / of F or Z (columns)
    unsigned char isz, warp, slot, pad;   // logical warp and accumulator slot that hold it
};
struct GramMidPlan {
    // staged column blocks of a stage: F blocks then Z blocks; the last block of each matrix is trimmed to its valid
    // 8-column groups.  blk_off: byte offset inside a stage, blk_cols: columns staged (box width), stage_bytes: total
    int blk_off[8] = {0}, blk_cols[8] = {0}, stage_bytes = 0;
    GramMidUnit units[GRAM_MID_WARPS * GRAM_MID_UNITS];
    std::vector<GramMidTile> tiles;
    int nsets = 0, kb = 0, nblocks = 0, nunits = 0;
    int sp_load[4] = {0, 0, 0, 0};    // DMMAs per k4-step and round of sets on each SM sub-partition
};
bool gram_mid_plan(int K, int T, GramMidPlan& plan);

struct GramWorkspace {
    DevBuf<double> partial;  // row-split partial tiles
    DevBuf<double> zzpart;   // per-block partial sums of squares
    DevBuf<int> pair_order;  // TMA path: job table + work counter
    int jobs_K = -1, jobs_T = -1, njobs = 0, job_tiles = 0;  // shape the cached job table was built for
    int64_t launches = 0;
    int last_nsplit = 0, last_npairs = 0;
    DevBuf<double> mid_tab;  // mid path: unit table + tile table of the cached shape
    int mid_K = -1, mid_T = -1, mid_ntiles = 0, mid_nsets = 0, mid_kb = 0, mid_nblocks = 0;
    int mid_blk_off[8] = {0}, mid_blk_cols[8] = {0}, mid_stage_bytes = 0;
    bool mid_ok = false;
    const char* last_path = "";
    StreamGuard guard;       // orders calls that reuse this workspace from different streams
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

// register-resident variant (gram_mid.cu): computes zz itself; returns false if the shape / alignment does not fit
bool gram_stats_mid_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                        const double* dw, double* dpacked, GramWorkspace& ws, int sms, cudaStream_t st);

// narrow variant (gram_narrow.cu) for K <= 48, T <= 16: returns false if it does not apply
bool gram_stats_narrow_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                           const double* dw, double* dpacked, GramWorkspace& ws, int sms, int nzz, cudaStream_t st);

}  // namespace frk
