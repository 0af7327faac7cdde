// Stem 3x3/s2/p1 convolution, 3 -> C (C = 16) channels (reference conv_3x3_bn,
// torchdet3d/models/mobilenetv3.py:110-115, used :178).
//
// Reads the NCHW fp32 image exactly as the reference API receives it and emits NHWC in the compute
// dtype, plus the BatchNorm batch statistics (sum, sum of squares per (sample, channel)).
// K = 27 is far too short for a tensor-core pipeline to matter: the layer is bound by reading the
// 602 KB/crop image, so it is a shared-memory-staged CUDA-core kernel.
#include "td3d_kernels.h"

namespace td3d {

static const int ST_TH = 4, ST_TW = 32;             // output tile (rows x cols) = 128 threads
static const int ST_IH = 2 * ST_TH + 1, ST_IW = 2 * ST_TW + 1;
static const int ST_C = 16;

// lane l ends with the sum over the warp of v[l] (31 shuffles instead of 32*5)
__device__ __forceinline__ float warp_transpose_sum32(float v[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16, n = 16; o >= 1; o >>= 1, n >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      float keep = hi ? v[i + n] : v[i];
      float send = hi ? v[i] : v[i + n];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

template <typename T>
__global__ void __launch_bounds__(ST_TH * ST_TW)
stem_fwd_kernel(const float* __restrict__ img, const float* __restrict__ w, T* __restrict__ y,
                float* __restrict__ stats, int H, int W, int Ho, int Wo, int C, const float* __restrict__ out_bias,
                int out_act) {
  pdl_entry();
  __shared__ float s_in[3][ST_IH][ST_IW + 1];
  __shared__ float s_w[27 * ST_C];
  __shared__ float s_acc[2 * ST_C];
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
  const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
  for (int i = threadIdx.x; i < 3 * ST_IH * ST_IW; i += blockDim.x) {
    int ci = i / (ST_IH * ST_IW), r = (i / ST_IW) % ST_IH, c = i % ST_IW;
    int iy = iy0 + r, ix = ix0 + c;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((size_t)b * 3 + ci) * H + iy) * W + ix);
    s_in[ci][r][c] = v;
  }
  const int ty = threadIdx.x / ST_TW, tx = threadIdx.x % ST_TW;
  const int oy = oy0 + ty, ox = ox0 + tx;
  const bool valid = oy < Ho && ox < Wo;
  // the image tile is staged once; the C output channels (16 for MobileNetV3, 32 / 40 for EfficientNet-B0 / B3) are
  // produced in slices of ST_C
  for (int c0 = 0; c0 < C; c0 += ST_C) {
    __syncthreads();                         // s_in ready (first slice) / previous slice done with s_w, s_acc
    for (int i = threadIdx.x; i < 27 * ST_C; i += blockDim.x) {
      const int c = c0 + i % ST_C;
      s_w[i] = c < C ? w[(i / ST_C) * C + c] : 0.f;
    }
    if (threadIdx.x < 2 * ST_C) s_acc[threadIdx.x] = 0.f;
    __syncthreads();
    float acc[ST_C];
#pragma unroll
    for (int c = 0; c < ST_C; ++c) acc[c] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          float x = s_in[ci][2 * ty + ky][2 * tx + kx];
          const float* wr = s_w + ((ci * 3 + ky) * 3 + kx) * ST_C;
#pragma unroll
          for (int c = 0; c < ST_C; ++c) acc[c] = fmaf(x, wr[c], acc[c]);
        }
    if (out_bias) {      // inference: BatchNorm folded into w / out_bias, activation before the single store
#pragma unroll
      for (int c = 0; c < ST_C; ++c) acc[c] = c0 + c < C ? act_fwd(acc[c] + __ldg(out_bias + c0 + c), out_act) : 0.f;
    }
    float red[32];
    if (valid) {
      T* yo = y + (((size_t)b * Ho + oy) * Wo + ox) * C + c0;
      store8(yo, acc);
      if (c0 + 8 < C) store8(yo + 8, acc + 8);
      // statistics are taken on the values as stored (bf16-rounded in bf16 mode), i.e. exactly the
      // tensor the consumer will normalise
#pragma unroll
      for (int c = 0; c < ST_C; ++c) {
        float v = to_f(from_f<T>(acc[c]));
        red[c] = v;
        red[ST_C + c] = v * v;
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c) red[c] = 0.f;
    }
    if (stats) {
      float tot = warp_transpose_sum32(red);
      atomicAdd(&s_acc[threadIdx.x & 31], tot);
      __syncthreads();
      if (threadIdx.x < 2 * ST_C && c0 + (threadIdx.x % ST_C) < C)
        atomicAdd(&stats[((size_t)b * 2 + threadIdx.x / ST_C) * C + c0 + threadIdx.x % ST_C], s_acc[threadIdx.x]);
    }
  }
}

int launch_stem_fwd(const float* img, const float* w27xC, void* y, float* stats, int B, int H, int W, int C,
                    int dtype, cudaStream_t st, const float* out_bias, int out_act) {
  TD3D_REQUIRE(C % 8 == 0 && C >= 8 && C <= 64, "stem: output channels must be a multiple of 8 in 8..64 (got %d)", C);
  int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  dim3 grid(ceil_div(Wo, ST_TW), ceil_div(Ho, ST_TH), B);
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(stem_fwd_kernel<bf16>, grid, ST_TH * ST_TW, 0, st, img, w27xC, (bf16*)y, stats, H, W, Ho, Wo, C, out_bias, out_act));
  else
    TD3D_CUDA(launch_kernel(stem_fwd_kernel<float>, grid, ST_TH * ST_TW, 0, st, img, w27xC, (float*)y, stats, H, W, Ho, Wo, C, out_bias, out_act));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// dW[co][ci][ky][kx] += sum_{b,oy,ox} gy[b,oy,ox,co] * img[b,ci,2oy+ky-1,2ox+kx-1]
//   gy = alpha[b,co]*g + beta[co]*y + gamma[b,co]   (lazy BN backward of the stem BN)
// Persistent CTAs walk contiguous ranges of 4x32-pixel tiles.  Warp w owns the filter row
// (ci, ky) = (w/3, w%3) with its three kx taps, lane = output column: per pixel 3 image values and
// the 16 gy channels feed 48 register accumulators (4 LDS.128 + 3 LDS per 48 FFMA).  The next tile's
// image patch and (g, y) vectors are prefetched into registers while the current tile is consumed.
static const int SW_WARPS = 9;
static const int SW_THREADS = 32 * SW_WARPS;                   // 288
static const int SW_NIMG = (3 * ST_IH * ST_IW + SW_THREADS - 1) / SW_THREADS;   // image values per thread and tile
static const int SW_GS = 20;                                   // s_gy pixel stride (floats): conflict-free LDS.128

template <typename T> struct SwRaw;
template <> struct SwRaw<bf16> {
  uint4 g, y;
  __device__ __forceinline__ void load(const bf16* gp, const bf16* yp) {
    g = __ldg(reinterpret_cast<const uint4*>(gp));
    y = __ldg(reinterpret_cast<const uint4*>(yp));
  }
  __device__ __forceinline__ void get(float gv[8], float yv[8]) const {
    const uint32_t a[4] = {g.x, g.y, g.z, g.w}, b[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gv[2 * i] = __uint_as_float(a[i] << 16); gv[2 * i + 1] = __uint_as_float(a[i] & 0xffff0000u);
      yv[2 * i] = __uint_as_float(b[i] << 16); yv[2 * i + 1] = __uint_as_float(b[i] & 0xffff0000u);
    }
  }
};
template <> struct SwRaw<float> {
  float4 g0, g1, y0, y1;
  __device__ __forceinline__ void load(const float* gp, const float* yp) {
    g0 = __ldg(reinterpret_cast<const float4*>(gp)); g1 = __ldg(reinterpret_cast<const float4*>(gp) + 1);
    y0 = __ldg(reinterpret_cast<const float4*>(yp)); y1 = __ldg(reinterpret_cast<const float4*>(yp) + 1);
  }
  __device__ __forceinline__ void get(float gv[8], float yv[8]) const {
    gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w; gv[4] = g1.x; gv[5] = g1.y; gv[6] = g1.z; gv[7] = g1.w;
    yv[0] = y0.x; yv[1] = y0.y; yv[2] = y0.z; yv[3] = y0.w; yv[4] = y1.x; yv[5] = y1.y; yv[6] = y1.z; yv[7] = y1.w;
  }
};

template <typename T>
__global__ void __launch_bounds__(SW_THREADS, 2)
stem_wgrad_kernel(const float* __restrict__ img, const T* __restrict__ g, const T* __restrict__ y,
                  const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ gamma,
                  float* __restrict__ dw, int B, int H, int W, int Ho, int Wo, int tiles_per_cta, int C) {
  pdl_entry();
  __shared__ float s_in[3 * ST_IH * ST_IW];      // [ci][row][col], odd row stride
  __shared__ __align__(16) float s_gy[ST_TH * ST_TW][SW_GS];
  __shared__ float s_al[ST_C], s_be[ST_C], s_ga[ST_C];
  const int tiles_x = (Wo + ST_TW - 1) / ST_TW, tiles_y = (Ho + ST_TH - 1) / ST_TH;
  const int n_tiles = B * tiles_x * tiles_y;
  const int t0 = blockIdx.x * tiles_per_cta, t1 = min(n_tiles, t0 + tiles_per_cta);
  if (t0 >= t1) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ci = warp / 3, ky = warp % 3;
  const bool has_gy = tid < ST_TH * ST_TW * 2;
  const int gp = tid >> 1, ghalf = tid & 1;      // (pixel, 8-channel half) of the staged gy vector
  const int c0 = blockIdx.y * ST_C;              // this CTA's 16-channel slice of the C output channels
  const bool half_ok = c0 + ghalf * 8 < C;
  if (tid < ST_C) s_be[tid] = c0 + tid < C ? beta[c0 + tid] : 0.f;
  float acc[3][ST_C];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < ST_C; ++c) acc[k][c] = 0.f;

  float r_img[SW_NIMG];
  SwRaw<T> r_gy;
  bool r_live = false;
  auto prefetch = [&](int tile) {
    const int b = tile / (tiles_x * tiles_y);
    const int oy0 = ((tile / tiles_x) % tiles_y) * ST_TH, ox0 = (tile % tiles_x) * ST_TW;
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    const float* ib = img + (size_t)b * 3 * H * W;
#pragma unroll
    for (int u = 0; u < SW_NIMG; ++u) {
      float v = 0.f;
      const int i = tid + u * SW_THREADS;            // flat index into s_in[3][ST_IH][ST_IW]
      if (i < 3 * ST_IH * ST_IW) {
        const int row = i / ST_IW;                    // ci * ST_IH + r
        const int iy = iy0 + row % ST_IH, ix = ix0 + i % ST_IW;
        if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W)
          v = __ldg(ib + ((size_t)(row / ST_IH) * H + iy) * W + ix);
      }
      r_img[u] = v;
    }
    r_live = false;
    if (has_gy) {
      const int oy = oy0 + gp / ST_TW, ox = ox0 + gp % ST_TW;
      if (oy < Ho && ox < Wo && half_ok) {
        const size_t off = (((size_t)b * Ho + oy) * Wo + ox) * C + c0 + ghalf * 8;
        r_gy.load(g + off, y + off);
        r_live = true;
      }
    }
  };

  prefetch(t0);
  int cur_b = -1;
  for (int tile = t0; tile < t1; ++tile) {
    const int b = tile / (tiles_x * tiles_y);
    if (b != cur_b) {           // CTA-uniform; the previous readers of s_al/s_ga are two barriers behind
      if (tid < ST_C) {
        s_al[tid] = c0 + tid < C ? alpha[(size_t)b * C + c0 + tid] : 0.f;
        s_ga[tid] = c0 + tid < C ? gamma[(size_t)b * C + c0 + tid] : 0.f;
      }
      cur_b = b;
      __syncthreads();
    }
    // ---- staged registers -> shared memory ----
#pragma unroll
    for (int u = 0; u < SW_NIMG; ++u)
      if (tid + u * SW_THREADS < 3 * ST_IH * ST_IW) s_in[tid + u * SW_THREADS] = r_img[u];
    if (has_gy) {
      float v[8];
      if (r_live) {
        float gv[8], yv[8];
        r_gy.get(gv, yv);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          v[j] = fmaf(s_al[ghalf * 8 + j], gv[j], fmaf(s_be[ghalf * 8 + j], yv[j], s_ga[ghalf * 8 + j]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(&s_gy[gp][ghalf * 8]);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();
    if (tile + 1 < t1) prefetch(tile + 1);
    // ---- accumulate ----
#pragma unroll
    for (int it = 0; it < ST_TH; ++it) {
      const float* xr = &s_in[(ci * ST_IH + 2 * it + ky) * ST_IW + 2 * lane];
      const float x0 = xr[0], x1 = xr[1], x2 = xr[2];
      const float4* gq = reinterpret_cast<const float4*>(&s_gy[it * ST_TW + lane][0]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 gv = gq[q];
        acc[0][4 * q + 0] = fmaf(x0, gv.x, acc[0][4 * q + 0]); acc[0][4 * q + 1] = fmaf(x0, gv.y, acc[0][4 * q + 1]);
        acc[0][4 * q + 2] = fmaf(x0, gv.z, acc[0][4 * q + 2]); acc[0][4 * q + 3] = fmaf(x0, gv.w, acc[0][4 * q + 3]);
        acc[1][4 * q + 0] = fmaf(x1, gv.x, acc[1][4 * q + 0]); acc[1][4 * q + 1] = fmaf(x1, gv.y, acc[1][4 * q + 1]);
        acc[1][4 * q + 2] = fmaf(x1, gv.z, acc[1][4 * q + 2]); acc[1][4 * q + 3] = fmaf(x1, gv.w, acc[1][4 * q + 3]);
        acc[2][4 * q + 0] = fmaf(x2, gv.x, acc[2][4 * q + 0]); acc[2][4 * q + 1] = fmaf(x2, gv.y, acc[2][4 * q + 1]);
        acc[2][4 * q + 2] = fmaf(x2, gv.z, acc[2][4 * q + 2]); acc[2][4 * q + 3] = fmaf(x2, gv.w, acc[2][4 * q + 3]);
      }
    }
    __syncthreads();
  }
  // ---- warp reduction, one atomic per (tap, channel) and warp; reference layout [co][ci][ky][kx] ----
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int c = 0; c < ST_C; ++c) {
      const float r = warp_sum(acc[k][c]);
      if (lane == 0 && c0 + c < C) atomicAdd(&dw[(c0 + c) * 27 + ci * 9 + ky * 3 + k], r);
    }
}

int launch_stem_wgrad(const float* img, const void* g, const void* y, const float* alpha, const float* beta,
                      const float* gamma, float* dw, int B, int H, int W, int C, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(C % 8 == 0 && C >= 8 && C <= 64, "stem wgrad: output channels must be a multiple of 8 in 8..64 (got %d)", C);
  int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  int n_tiles = B * ceil_div(Wo, ST_TW) * ceil_div(Ho, ST_TH);
  int ctas = n_tiles < 148 * 2 ? n_tiles : 148 * 2;
  int per = ceil_div(n_tiles, ctas);
  dim3 grid(ceil_div(n_tiles, per), ceil_div(C, ST_C));
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(stem_wgrad_kernel<bf16>, grid, SW_THREADS, 0, st, img, (const bf16*)g, (const bf16*)y, alpha, beta, gamma, dw, B, H, W, Ho, Wo, per, C));
  else
    TD3D_CUDA(launch_kernel(stem_wgrad_kernel<float>, grid, SW_THREADS, 0, st, img, (const float*)g, (const float*)y, alpha, beta, gamma, dw, B, H, W, Ho, Wo, per, C));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
