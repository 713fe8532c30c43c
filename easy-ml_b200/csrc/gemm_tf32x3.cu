// gemm_tf32x3.cu -- placeholder until the tcgen05 kernel lands: reports "not handled" so f32 uses SIMT.
#include "common.h"
namespace eml {
int launch_gemm_tf32x3_f32(DeviceCtx*, const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t,
                           Strides3, Strides3, Strides3, bool, cudaStream_t, bool* handled) {
    *handled = false;
    return EML_OK;
}
}  // namespace eml
