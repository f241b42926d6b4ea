// Host <-> device copies of R's pageable matrices at the pinned-memory rate (see host_staging.cu).
#pragma once
#include "frk_common.cuh"

namespace frk {

// Column-major `rows x cols` doubles from host memory (leading dimension ld_h) to device memory (ld_d) on stream st.
// On return the host data has been consumed (the device-side copies may still be in flight on st).
void upload_matrix(double* d_dst, int64_t ld_d, const double* h_src, int64_t ld_h, int64_t rows, int64_t cols, cudaStream_t st);

bool host_pointer_is_pinned(const void* p);
int host_copy_threads();

}  // namespace frk
