// Depthwise column walker, fwd kernels for bf16 activations (see dwc_launch.cuh / dwc_core.cuh).
#include "dwc_launch.cuh"

namespace td3d {
int launch_dw_fwd_cw_bf16(const DwArgs& a, cudaStream_t st) { return dwc_fwd_dispatch<bf16>(a, st); }
}  // namespace td3d
