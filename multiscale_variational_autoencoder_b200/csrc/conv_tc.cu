// tcgen05 / TMEM tensor-core implicit-GEMM path (MVAE_PREC_TF32).  Shapes not covered return MVAE_ERR_UNSUPPORTED and
// the caller takes the fp32 path (a precision choice, not a device fallback: both are sm_100a kernels).
#include "common.cuh"

namespace mvae {
struct ConvGeom;
int conv_fwd_tc(const ConvGeom&, const float*, const float*, const float*, const float*, const float*, int, float*,
                cudaStream_t) { return MVAE_ERR_UNSUPPORTED; }
int conv_dgrad_tc(const ConvGeom&, const float*, const float*, const float*, const float*, const float*, int, float*,
                  cudaStream_t) { return MVAE_ERR_UNSUPPORTED; }
int conv_wgrad_tc(const ConvGeom&, const float*, const float*, const float*, float*, float*, cudaStream_t) {
    return MVAE_ERR_UNSUPPORTED;
}
}  // namespace mvae
