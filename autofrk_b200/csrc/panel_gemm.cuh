// Y = F * B for a tall column-major panel F (rows x K) and a small matrix B (K x q): the shape of every n-sized product
// left on the fit / predict path once the statistics are formed (see panel_gemm.cu).
#pragma once
#include "frk_common.cuh"

namespace frk {

// dF rows x K (ld ldf), dB K x q (ld ldb), dY rows x q (ld ldy), all on the device; asynchronous on st.
void panel_times_small_dev(const double* dF, int64_t rows, int64_t ldf, int K, const double* dB, int64_t ldb, int q,
                           double* dY, int64_t ldy, cudaStream_t st);

}  // namespace frk
