// Depthwise k x k convolution dispatch (reference op: nn.Conv2d(hidden, hidden, k, s, (k-1)//2, groups=hidden) inside
// InvertedResidual, torchdet3d/models/mobilenetv3.py:136,152).
//   forward : three kernels, chosen per layer from measurements (scripts/dw_bench.py, profiles/r02_v5_dw_bench.txt):
//             column walker (k_dwc.cu) for large planes, SiLU inputs and the inference epilogue; row walker (k_dww.cu)
//             for stride-1 planes whose rows fit a warp; tiled persistent kernels (k_dw2.cu) for the rest
//   backward: one-pass column walker (k_dwc.cu / dwc_core.cuh: both gradients + BatchNorm sums from a single read of g,
//             y_out, x) on every layer
#include "td3d_kernels.h"

#include <stdlib.h>

namespace td3d {

// TD3D_DW_FWD_CW=1 (A/B measurements) routes every forward depthwise conv through the column walker.
static int dw_fwd_cw() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TD3D_DW_FWD_CW");
    v = e ? atoi(e) : 0;
  }
  return v;
}

int launch_dw_fwd(const DwArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B <= 65535, "dw fwd: C=%d must be a multiple of 8", a.C);
  TD3D_REQUIRE((a.k == 3 || a.k == 5) && (a.stride == 1 || a.stride == 2), "dw fwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  // SiLU inputs (EfficientNet) and the bias + activation epilogue of inference exist only in the column walker
  if (a.xf.act == TD3D_ACT_SILU || a.out_bias || a.out_act != TD3D_ACT_NONE || dw_fwd_cw()) return launch_dw_fwd_cw(a, dtype, st);
  // batch 256 bf16, MobileNetV3-large: the column walker wins on the 112x112x64 s2 (188 vs 207 us), 56x56x72 (105 vs 174 us)
  // and 56x56x72 k5 s2 (121 vs 145 us) layers and loses on 112x112x16 (too few channel groups per pixel) and on every
  // plane of 28x28 and below (few columns to walk between window refills)
  if (a.H >= 56 && a.C >= 32) return launch_dw_fwd_cw(a, dtype, st);
  if (dw_walker_supported(a.H, a.W, a.C, a.k, a.stride)) return launch_dw_fwd_walker(a, dtype, st);
  return launch_dw_fwd_v2(a, dtype, st);
}

// One implementation: after the inner-loop clean-up (compile-time activations, 32-bit offsets) the one-pass column walker
// beats the two-kernel backward of round 1 on every MobileNetV3-large layer at batch 256 bf16 (e.g. 56x56x72: 237 vs 604 us,
// 14x14x672: 200 vs 263 us, 7x7x960 k5: 165 vs 172 us), so the data- / weight-gradient pair was deleted.
int launch_dw_bwd(const DwBwdArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B <= 65535, "dw bwd: C=%d must be a multiple of 8", a.C);
  TD3D_REQUIRE((a.k == 3 || a.k == 5) && (a.stride == 1 || a.stride == 2), "dw bwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  return launch_dw_bwd_fused(a, dtype, st);
}

}  // namespace td3d
