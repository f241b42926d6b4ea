// K1 instantiations for d = 2 (one translation unit per dimension so they compile in parallel).
#include "mrts_predict_kernel.cuh"
namespace frk {
static const PredictKernelInfo k_table_d2[] = {FRK_PREDICT_CONFIG_TABLE(2)};
const PredictKernelInfo* predict_kernels_d2(int* count) {
    *count = int(sizeof(k_table_d2) / sizeof(k_table_d2[0]));
    return k_table_d2;
}
}  // namespace frk
