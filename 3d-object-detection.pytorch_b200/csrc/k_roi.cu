// ROI front-end: the step in front of the regressor.  Crops the 2D-detector boxes out of uint8 frames,
// resizes each crop to the network input with OpenCV's uint8 INTER_LINEAR rule, swaps BGR->RGB, normalises and writes
// the float32 NCHW batch the regressor consumes -- one kernel, no host round trip, no intermediate uint8 crop.
// Replaces, per box: `Regressor.crop` + `IEModel._preprocess` (torchdet3d/utils/ie_wrappers.py:155-158,18-21:
// frame[y0:y1, x0:x1] -> cv.resize(img, (w, h)) -> transpose(2, 0, 1)), `ConvertColor` (utils/transforms.py:10-17) and
// the Normalize of the test-time pipeline with the constants of configs/default_config.py:9-10.
//
// Bit-exactness with OpenCV is part of the contract (oracle/roi_port.py pins the rule against cv2 itself):
//   fx = float((dx + 0.5) * (src_w / dst_w) - 0.5), sx = floor(fx), fx -= sx, clamped at both borders (fx = 0);
//   11-bit integer weights a = rint(w * 2048); horizontal pass r = S[sx]*a0 + S[sx+1]*a1;
//   vertical pass (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2 with rows clamped to the crop.
// One thread per output pixel (all three channels): the frame is read through L2 (a 1080p frame is 6 MB), the fp32
// output planes are written coalesced along x.
#include "td3d_kernels.h"

namespace td3d {

struct RoiArgs {
  const uint8_t* frames; int n_frames, fh, fw;
  const int32_t* boxes;          // [N][5]: frame index, x0, y0, x1, y1 (x1, y1 exclusive, as the numpy slice)
  int n_boxes, oh, ow;
  float mean[3], inv[3];         // mean*255, 1/(std*255) per OUTPUT channel
  int swap_rb;
  float* out;                    // [N][3][oh][ow]
};

__device__ __forceinline__ void roi_axis(int d, int src, int dst, int& i0, int& i1, int& w0, int& w1, bool clamp_weights) {
  const double scale = (double)src / (double)dst;
  float f = (float)(((double)d + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (clamp_weights) {                      // x axis: OpenCV pins the weight at the borders
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= src - 1) { f = 0.f; s = src - 1; }
  }
  w1 = __float2int_rn(f * 2048.f);
  w0 = __float2int_rn((1.f - f) * 2048.f);
  i0 = min(max(s, 0), src - 1);             // y axis: rows are clamped, weights kept
  i1 = min(max(s + 1, 0), src - 1);
}

__global__ void __launch_bounds__(256) roi_crop_resize_kernel(RoiArgs a) {
  pdl_entry();
  const int n = blockIdx.z;
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= a.ow) return;
  const int32_t* bx = a.boxes + (size_t)n * 5;
  const int f = min(max(bx[0], 0), a.n_frames - 1);
  const int x0 = min(max(bx[1], 0), a.fw), y0 = min(max(bx[2], 0), a.fh);
  const int x1 = min(max(bx[3], 0), a.fw), y1 = min(max(bx[4], 0), a.fh);
  const int cw = x1 - x0, ch = y1 - y0;
  float* o = a.out + (((size_t)n * 3) * a.oh + dy) * a.ow + dx;
  const size_t plane = (size_t)a.oh * a.ow;
  if (cw <= 0 || ch <= 0) {                 // empty box: a defined (zero-pixel) crop instead of garbage
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * plane] = (0.f - a.mean[c]) * a.inv[c];
    return;
  }
  int sx0, sx1, a0, a1, sy0, sy1, b0, b1;
  roi_axis(dx, cw, a.ow, sx0, sx1, a0, a1, true);
  roi_axis(dy, ch, a.oh, sy0, sy1, b0, b1, false);
  const uint8_t* fr = a.frames + (size_t)f * a.fh * a.fw * 3;
  const uint8_t* r0 = fr + ((size_t)(y0 + sy0) * a.fw + x0) * 3;
  const uint8_t* r1 = fr + ((size_t)(y0 + sy1) * a.fw + x0) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int sc = a.swap_rb ? 2 - c : c;
    const int h0 = (int)r0[sx0 * 3 + sc] * a0 + (int)r0[sx1 * 3 + sc] * a1;
    const int h1 = (int)r1[sx0 * 3 + sc] * a0 + (int)r1[sx1 * 3 + sc] * a1;
    int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    v = min(max(v, 0), 255);
    o[c * plane] = __fmul_rn(__fsub_rn((float)v, a.mean[c]), a.inv[c]);
  }
}

int launch_roi_crop_resize(const uint8_t* frames, int n_frames, int fh, int fw, const int32_t* boxes, int n_boxes, int oh,
                           int ow, const float* mean255, const float* inv_std255, int swap_rb, float* out, cudaStream_t st) {
  TD3D_REQUIRE(frames && boxes && out && mean255 && inv_std255, "roi: null argument");
  TD3D_REQUIRE(n_frames > 0 && fh > 0 && fw > 0 && n_boxes > 0 && n_boxes <= 65535 && oh > 0 && oh <= 65535 && ow > 0,
               "roi: bad sizes frames=%d %dx%d boxes=%d out=%dx%d", n_frames, fh, fw, n_boxes, oh, ow);
  RoiArgs a;
  a.frames = frames; a.n_frames = n_frames; a.fh = fh; a.fw = fw; a.boxes = boxes; a.n_boxes = n_boxes; a.oh = oh; a.ow = ow;
  for (int c = 0; c < 3; ++c) { a.mean[c] = mean255[c]; a.inv[c] = inv_std255[c]; }
  a.swap_rb = swap_rb; a.out = out;
  const int threads = ow >= 256 ? 256 : ((ow + 31) / 32) * 32;
  TD3D_CUDA(launch_kernel(roi_crop_resize_kernel, dim3(ceil_div(ow, threads), oh, n_boxes), threads, 0, st, a));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
