// Depthwise column walker, fwd kernels for float activations (see dwc_launch.cuh / dwc_core.cuh).
#include "dwc_launch.cuh"

namespace td3d {
int launch_dw_fwd_cw_f32(const DwArgs& a, cudaStream_t st) { return dwc_fwd_dispatch<float>(a, st); }
}  // namespace td3d
