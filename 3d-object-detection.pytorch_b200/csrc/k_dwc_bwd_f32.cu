// Depthwise column walker, bwd kernels for float activations (see dwc_launch.cuh / dwc_core.cuh).
#include "dwc_launch.cuh"

namespace td3d {
int launch_dw_bwd_fused_f32(const DwBwdArgs& a, cudaStream_t st) { return dwc_dispatch<float>(a, st); }
}  // namespace td3d
