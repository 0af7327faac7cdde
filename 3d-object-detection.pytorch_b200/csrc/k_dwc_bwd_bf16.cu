// Depthwise column walker, bwd kernels for bf16 activations (see dwc_launch.cuh / dwc_core.cuh).
#include "dwc_launch.cuh"

namespace td3d {
int launch_dw_bwd_fused_bf16(const DwBwdArgs& a, cudaStream_t st) { return dwc_dispatch<bf16>(a, st); }
}  // namespace td3d
