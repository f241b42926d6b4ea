// K4: Gram / cross-product / sum-of-squares reductions (see gram.cu).
#pragma once
#include "frk_common.cuh"

namespace frk {

struct GramLayout {
    int ntf = 0, ntz = 0, npairs = 0, nsplit = 0;
};

struct GramWorkspace {
    DevBuf<double> partial;  // row-split partial tiles
    DevBuf<double> zzpart;   // per-block partial sums of squares
    DevBuf<int> pair_order;  // TMA path: tile-pair order + work counter
    int order_ntf = -1, order_ntz = -1;  // shape the cached order was built for
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
                        const double* dw, GramWorkspace& ws, int sms, GramLayout& lay, cudaStream_t st);

// narrow variant (gram_narrow.cu) for K <= 48, T <= 16: returns false if it does not apply
bool gram_stats_narrow_dev(const double* dF, int64_t n, int K, int64_t ldf, const double* dZ, int T, int64_t ldz,
                           const double* dw, double* dpacked, GramWorkspace& ws, int sms, int nzz, cudaStream_t st);

}  // namespace frk
