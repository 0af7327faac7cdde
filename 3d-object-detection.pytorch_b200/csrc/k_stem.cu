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
                float* __restrict__ stats, int H, int W, int Ho, int Wo) {
  __shared__ float s_in[3][ST_IH][ST_IW + 1];
  __shared__ float s_w[27 * ST_C];
  __shared__ float s_acc[2 * ST_C];
  const int b = blockIdx.z;
  const int oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
  const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
  for (int i = threadIdx.x; i < 27 * ST_C; i += blockDim.x) s_w[i] = w[i];
  if (threadIdx.x < 2 * ST_C) s_acc[threadIdx.x] = 0.f;
  for (int i = threadIdx.x; i < 3 * ST_IH * ST_IW; i += blockDim.x) {
    int ci = i / (ST_IH * ST_IW), r = (i / ST_IW) % ST_IH, c = i % ST_IW;
    int iy = iy0 + r, ix = ix0 + c;
    float v = 0.f;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((size_t)b * 3 + ci) * H + iy) * W + ix);
    s_in[ci][r][c] = v;
  }
  __syncthreads();
  const int ty = threadIdx.x / ST_TW, tx = threadIdx.x % ST_TW;
  const int oy = oy0 + ty, ox = ox0 + tx;
  const bool valid = oy < Ho && ox < Wo;
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
  float red[32];
  if (valid) {
    T* yo = y + (((size_t)b * Ho + oy) * Wo + ox) * ST_C;
    store8(yo, acc);
    store8(yo + 8, acc + 8);
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
    if (threadIdx.x < 2 * ST_C) atomicAdd(&stats[(size_t)b * 2 * ST_C + threadIdx.x], s_acc[threadIdx.x]);
  }
}

int launch_stem_fwd(const float* img, const float* w27xC, void* y, float* stats, int B, int H, int W, int C,
                    int dtype, cudaStream_t st) {
  TD3D_REQUIRE(C == ST_C, "stem: only %d output channels supported (got %d)", ST_C, C);
  int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  dim3 grid(ceil_div(Wo, ST_TW), ceil_div(Ho, ST_TH), B);
  if (dtype == TD3D_BF16)
    stem_fwd_kernel<bf16><<<grid, ST_TH * ST_TW, 0, st>>>(img, w27xC, (bf16*)y, stats, H, W, Ho, Wo);
  else
    stem_fwd_kernel<float><<<grid, ST_TH * ST_TW, 0, st>>>(img, w27xC, (float*)y, stats, H, W, Ho, Wo);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// dW[co][ci][ky][kx] += sum_{b,oy,ox} gy[b,oy,ox,co] * img[b,ci,2oy+ky-1,2ox+kx-1]
//   gy = alpha[b,co]*g + beta[co]*y + gamma[b,co]   (lazy BN backward of the stem BN)
// Persistent blocks: thread (tap t, pixel lane pl) keeps 16 accumulators across all its tiles and
// flushes once (432 atomics per block).
static const int SW_PL = 8;                 // pixel lanes per tap
static const int SW_THREADS = 27 * SW_PL;   // 216

template <typename T>
__global__ void __launch_bounds__(SW_THREADS)
stem_wgrad_kernel(const float* __restrict__ img, const T* __restrict__ g, const T* __restrict__ y,
                  const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ gamma,
                  float* __restrict__ dw, int B, int H, int W, int Ho, int Wo) {
  __shared__ float s_in[3][ST_IH][ST_IW + 1];
  __shared__ __align__(16) float s_gy[ST_TH * ST_TW][ST_C];
  __shared__ float s_dw[27 * ST_C];
  const int tiles_x = (Wo + ST_TW - 1) / ST_TW, tiles_y = (Ho + ST_TH - 1) / ST_TH;
  const int n_tiles = B * tiles_x * tiles_y;
  const int t = threadIdx.x / SW_PL, pl = threadIdx.x % SW_PL;
  const int ci = t / 9, ky = (t / 3) % 3, kx = t % 3;
  float acc[ST_C];
#pragma unroll
  for (int c = 0; c < ST_C; ++c) acc[c] = 0.f;
  for (int i = threadIdx.x; i < 27 * ST_C; i += blockDim.x) s_dw[i] = 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int b = tile / (tiles_x * tiles_y);
    const int oy0 = ((tile / tiles_x) % tiles_y) * ST_TH, ox0 = (tile % tiles_x) * ST_TW;
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * ST_IH * ST_IW; i += blockDim.x) {
      int c3 = i / (ST_IH * ST_IW), r = (i / ST_IW) % ST_IH, c = i % ST_IW;
      int iy = iy0 + r, ix = ix0 + c;
      float v = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(img + (((size_t)b * 3 + c3) * H + iy) * W + ix);
      s_in[c3][r][c] = v;
    }
    for (int i = threadIdx.x; i < ST_TH * ST_TW * 2; i += blockDim.x) {
      int p = i >> 1, half = i & 1;
      int oy = oy0 + p / ST_TW, ox = ox0 + p % ST_TW;
      float v[8];
      if (oy < Ho && ox < Wo) {
        size_t off = (((size_t)b * Ho + oy) * Wo + ox) * ST_C + half * 8;
        float gv[8], yv[8], al[8], be[8], ga[8];
        load8(g + off, gv);
        load8(y + off, yv);
        loadf8(alpha + (size_t)b * ST_C + half * 8, al);
        loadf8(beta + half * 8, be);
        loadf8(gamma + (size_t)b * ST_C + half * 8, ga);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(al[j], gv[j], fmaf(be[j], yv[j], ga[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s_gy[p][half * 8 + j] = v[j];
    }
    __syncthreads();
    if (t < 27) {
      for (int p = pl; p < ST_TH * ST_TW; p += SW_PL) {
        int ty = p / ST_TW, tx = p % ST_TW;
        float x = s_in[ci][2 * ty + ky][2 * tx + kx];
        const float4* gy4 = reinterpret_cast<const float4*>(s_gy[p]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 gq = gy4[q];
          acc[4 * q + 0] = fmaf(x, gq.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(x, gq.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(x, gq.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(x, gq.w, acc[4 * q + 3]);
        }
      }
    }
  }
  __syncthreads();
  if (t < 27) {
#pragma unroll
    for (int c = 0; c < ST_C; ++c) atomicAdd(&s_dw[c * 27 + t], acc[c]);   // reference layout [co][ci][ky][kx]
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 27 * ST_C; i += blockDim.x) atomicAdd(&dw[i], s_dw[i]);
}

int launch_stem_wgrad(const float* img, const void* g, const void* y, const float* alpha, const float* beta,
                      const float* gamma, float* dw, int B, int H, int W, int C, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(C == ST_C, "stem wgrad: only %d output channels supported", ST_C);
  int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  int n_tiles = B * ceil_div(Wo, ST_TW) * ceil_div(Ho, ST_TH);
  int grid = n_tiles < 148 * 4 ? n_tiles : 148 * 4;
  if (dtype == TD3D_BF16)
    stem_wgrad_kernel<bf16><<<grid, SW_THREADS, 0, st>>>(img, (const bf16*)g, (const bf16*)y, alpha, beta, gamma, dw, B, H, W, Ho, Wo);
  else
    stem_wgrad_kernel<float><<<grid, SW_THREADS, 0, st>>>(img, (const float*)g, (const float*)y, alpha, beta, gamma, dw, B, H, W, Ho, Wo);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
