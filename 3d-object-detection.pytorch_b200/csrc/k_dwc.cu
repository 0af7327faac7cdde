// Entry points of the depthwise column walkers (kernels: k_dwc_{bwd,fwd}_{bf16,f32}.cu).
#include "td3d_kernels.h"

namespace td3d {

int launch_dw_bwd_fused_bf16(const DwBwdArgs& a, cudaStream_t st);
int launch_dw_bwd_fused_f32(const DwBwdArgs& a, cudaStream_t st);
int launch_dw_fwd_cw_bf16(const DwArgs& a, cudaStream_t st);
int launch_dw_fwd_cw_f32(const DwArgs& a, cudaStream_t st);

int launch_dw_bwd_fused(const DwBwdArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B > 0 && a.H > 0 && a.W > 0 && a.gx && a.dw, "dw_bwd_fused: bad arguments");
  TD3D_REQUIRE((double)a.H * a.W * a.C < 2147483648.0, "dw_bwd_fused: a sample plane of %d x %d x %d exceeds the 32-bit offset range", a.H, a.W, a.C);
  return dtype == TD3D_BF16 ? launch_dw_bwd_fused_bf16(a, st) : launch_dw_bwd_fused_f32(a, st);
}

int launch_dw_fwd_cw(const DwArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B > 0 && a.H > 0 && a.W > 0 && a.y, "dw_fwd_cw: bad arguments");
  TD3D_REQUIRE((double)a.H * a.W * a.C < 2147483648.0, "dw_fwd_cw: a sample plane of %d x %d x %d exceeds the 32-bit offset range", a.H, a.W, a.C);
  return dtype == TD3D_BF16 ? launch_dw_fwd_cw_bf16(a, st) : launch_dw_fwd_cw_f32(a, st);
}

}  // namespace td3d
