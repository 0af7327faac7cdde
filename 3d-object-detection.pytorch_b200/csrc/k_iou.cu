// Batched 2D-keypoint based 3D IoU and EPnP lift: one sample (pair) per thread, double precision,
// per-thread body in iou_core.cuh.  Reference: torchdet3d/evaluation/metrics.py:70-89 (a Python loop over the batch with a
// numpy eigen-solve, a scipy Qhull call and a device -> host copy of the keypoints), torchdet3d/utils/geometry.py:51-108.
// The work per pair is ~10^5 flops of branchy fp64 on ~4 KB of thread-local state, so the kernel is latency-bound by design;
// what it buys is that the evaluation loop stays on the device and the batch runs in parallel.
#include "iou_core.cuh"
#include "td3d_kernels.h"

namespace td3d {

struct IouCam { double v[4]; };      // fx, fy, cx, cy of the NDC camera matrix

__global__ void __launch_bounds__(64) iou_2d_based_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int n,
                                                           int portrait, IouCam cam, double* __restrict__ out) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[18], g[18];
  for (int k = 0; k < 18; ++k) { p[k] = pred[(size_t)i * 18 + k]; g[k] = gt[(size_t)i * 18 + k]; }
  out[i] = iou3d::iou_from_keypoints(p, g, portrait, cam.v);
}

__global__ void __launch_bounds__(64) lift_2d_kernel(const float* __restrict__ kp, int n, int portrait, IouCam cam,
                                                      double* __restrict__ out) {
  pdl_entry();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[18];
  for (int k = 0; k < 18; ++k) p[k] = kp[(size_t)i * 18 + k];
  double l[9][3];
  iou3d::lift_2d(p, portrait, cam.v, l);
  for (int k = 0; k < 9; ++k)
    for (int c = 0; c < 3; ++c) out[(size_t)i * 27 + k * 3 + c] = l[k][c];
}

static IouCam make_cam(const double* cam_ndc) {
  IouCam c;
  // default: the reference's camera matrix [[1, 0, .5], [0, 1, .5], [0, 0, 1]] in NDC form (geometry.py:16-37)
  c.v[0] = cam_ndc ? cam_ndc[0] : 2.0; c.v[1] = cam_ndc ? cam_ndc[1] : 2.0;
  c.v[2] = cam_ndc ? cam_ndc[2] : 0.0; c.v[3] = cam_ndc ? cam_ndc[3] : 0.0;
  return c;
}

int launch_iou_2d_based(const float* pred_kp, const float* gt_kp, int n, int portrait, const double* cam_ndc, double* iou,
                        cudaStream_t st) {
  TD3D_REQUIRE(n >= 0 && (n == 0 || (pred_kp && gt_kp && iou)), "iou_2d_based: null argument");
  if (n == 0) return TD3D_OK;
  TD3D_CUDA(launch_kernel(iou_2d_based_kernel, ceil_div(n, 64), 64, 0, st, pred_kp, gt_kp, n, portrait, make_cam(cam_ndc), iou));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_lift_2d(const float* kp, int n, int portrait, const double* cam_ndc, double* out, cudaStream_t st) {
  TD3D_REQUIRE(n >= 0 && (n == 0 || (kp && out)), "lift_2d: null argument");
  if (n == 0) return TD3D_OK;
  TD3D_CUDA(launch_kernel(lift_2d_kernel, ceil_div(n, 64), 64, 0, st, kp, n, portrait, make_cam(cam_ndc), out));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
