// Depthwise k x k convolution dispatch (reference op: nn.Conv2d(hidden, hidden, k, s, (k-1)//2, groups=hidden) inside
// InvertedResidual, torchdet3d/models/mobilenetv3.py:136,152).
//   forward : row walker (k_dww.cu) for small stride-1 planes, tiled persistent kernels (k_dw2.cu) otherwise
//   backward: one-pass column walker (k_dwc.cu / dwc_core.cuh: both gradients + BatchNorm sums from a single read of g,
//             y_out, x); the row-walker data- / weight-gradient pair of k_dww.cu only for 7x7 planes with 5x5 taps
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
  if (dw_walker_supported(a.H, a.W, a.C, a.k, a.stride)) return launch_dw_fwd_walker(a, dtype, st);
  return launch_dw_fwd_v2(a, dtype, st);
}

// Backward implementation choice, measured per layer on B200 (scripts/dw_bench.py, profiles/r02_dw_bench.txt, batch 256
// bf16): after the inner-loop clean-up the one-pass column walker beats the two-kernel backward of round 1 on every
// MobileNetV3-large layer (e.g. 112x112x64 s2: 407 vs 661 us, 56x56x72: 237 vs 604 us, 14x14x672 k5 s2: 209 vs 290 us)
// except the 7x7 5x5 layers (210 vs 177 us: 49-pixel planes give a thread only 7 columns to walk between flushes of its 25
// tap accumulators).  In the model: 12.65 vs 13.08 ms per step.  SiLU layers (EfficientNet) exist only in the one-pass
// kernel.  TD3D_DW_BWD_FUSED=0 / 1 forces the split / one-pass kernels everywhere (A/B measurements).
static int dw_bwd_force() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("TD3D_DW_BWD_FUSED");
    v = e ? atoi(e) : -1;
  }
  return v;
}

bool dw_bwd_is_split(const DwBwdArgs& a) {
  if (a.xf.act == TD3D_ACT_SILU || !dw_walker_supported(a.H, a.W, a.C, a.k, a.stride)) return false;   // split twins exist for small stride-1 planes only
  if (dw_bwd_force() >= 0) return dw_bwd_force() == 0;
  return a.k == 5 && a.H * a.W <= 64;
}

int launch_dw_bwd(const DwBwdArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B <= 65535, "dw bwd: C=%d must be a multiple of 8", a.C);
  TD3D_REQUIRE((a.k == 3 || a.k == 5) && (a.stride == 1 || a.stride == 2), "dw bwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  if (!dw_bwd_is_split(a)) return launch_dw_bwd_fused(a, dtype, st);
  return launch_dw_bwd_walker(a, dtype, st);
}

}  // namespace td3d
