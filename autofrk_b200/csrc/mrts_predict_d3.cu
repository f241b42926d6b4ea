// K1 instantiations for d = 3 (one translation unit per dimension so they compile in parallel).
#include "mrts_predict_kernel.cuh"
namespace frk {
static const PredictKernelInfo k_table_d3[] = {FRK_PREDICT_CONFIG_TABLE(3)};
const PredictKernelInfo* predict_kernels_d3(int* count) {
    *count = int(sizeof(k_table_d3) / sizeof(k_table_d3[0]));
    return k_table_d3;
}
}  // namespace frk
