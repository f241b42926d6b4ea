// K1 instantiations for d = 1 (one translation unit per dimension so they compile in parallel).
#include "mrts_predict_kernel.cuh"
namespace frk {
static const PredictKernelInfo k_table_d1[] = {FRK_PREDICT_CONFIG_TABLE(1)};
const PredictKernelInfo* predict_kernels_d1(int* count) {
    *count = int(sizeof(k_table_d1) / sizeof(k_table_d1[0]));
    return k_table_d1;
}
}  // namespace frk
