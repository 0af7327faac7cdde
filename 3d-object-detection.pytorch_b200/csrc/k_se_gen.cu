// Squeeze-and-Excite FCs, general flavour (torchvision SqueezeExcitation inside EfficientNet MBConv --
// BASELINE configs 3 / 5; oracle = torchvision .features through the reference's model_wrapper):
//   z = mean_HW(x);  a1 = W1 z + b1;  hid = silu(a1);  gate = sigmoid(W2 hid + b2);  out = x * gate
// Squeeze widths here are max(1, C_in // 4) = 4, 6, 10, 34, 58 ...: not multiples of 4, so the vectorised kernels
// of k_se.cu (128-bit weight rows) do not apply.  The work is tiny (<= 2304 x 96 MACs per sample), so one CTA per
// sample with warp-per-hidden-unit dot products is enough; weight gradients take one thread per matrix element
// over the batch (no atomics).
#include "td3d_kernels.h"

namespace td3d {

__device__ __forceinline__ float seg_sigmoid(float u) { return __fdividef(1.f, 1.f + __expf(-u)); }

// grid = B, block = 256, smem = (C + Ch) floats
__global__ void __launch_bounds__(256) se_gen_fwd_kernel(SeArgs a) {
  pdl_entry();
  extern __shared__ float sm[];
  float* z = sm;
  float* hs = sm + a.C;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float v = a.pool_stats[((size_t)b * 2) * a.C + c] * a.inv_hw;
    if (a.scale) v = fmaf(v, a.scale[c], a.shift[c]);
    z[c] = v;
    a.zbar[(size_t)b * a.C + c] = v;
  }
  __syncthreads();
  for (int h = warp; h < a.Ch; h += 8) {
    const float* w = a.w1 + (size_t)h * a.C;
    float acc = 0.f;
    for (int c = lane; c < a.C; c += 32) acc = fmaf(__ldg(w + c), z[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float a1 = acc + a.b1[h];
      a.hid[(size_t)b * a.Ch + h] = a1;                 // pre-activation: backward needs silu'(a1)
      hs[h] = a1 * seg_sigmoid(a1);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float acc = a.b2[c];
    if (a.w2t) {                                         // W2^T [Ch, C]: lanes read consecutive channels
      for (int h = 0; h < a.Ch; ++h) acc = fmaf(__ldg(a.w2t + (size_t)h * a.C + c), hs[h], acc);
    } else {
      const float* w = a.w2 + (size_t)c * a.Ch;
      for (int h = 0; h < a.Ch; ++h) acc = fmaf(__ldg(w + h), hs[h], acc);
    }
    a.pre[(size_t)b * a.C + c] = acc;
    a.gate[(size_t)b * a.C + c] = seg_sigmoid(acc);
  }
}

// grid = B, block = 256, smem = (C + Ch) floats.  g_gate[b,c] = sum_HW g_out * x arrives as the second statistic.
__global__ void __launch_bounds__(256) se_gen_bwd_kernel(SeBwdArgs a) {
  pdl_entry();
  extern __shared__ float sm[];
  float* gp = sm;
  float* ga = sm + a.C;
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const float p1 = a.bwd_stats[((size_t)b * 2) * a.C + c], p2 = a.bwd_stats[((size_t)b * 2 + 1) * a.C + c];
    const float gs = a.scale ? fmaf(a.scale[c], p2, a.shift[c] * p1) : p2;
    const float g = seg_sigmoid(a.pre[(size_t)b * a.C + c]);
    const float v = gs * g * (1.f - g);
    gp[c] = v;
    a.g_pre[(size_t)b * a.C + c] = v;
  }
  __syncthreads();
  for (int h = warp; h < a.Ch; h += 8) {
    const float* w = a.w2t + (size_t)h * a.C;            // W2^T [Ch, C]
    float acc = 0.f;
    for (int c = lane; c < a.C; c += 32) acc = fmaf(__ldg(w + c), gp[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      const float a1 = a.hid[(size_t)b * a.Ch + h];
      const float s = seg_sigmoid(a1);
      const float v = acc * s * fmaf(a1, 1.f - s, 1.f);
      ga[h] = v;
      a.g_hid[(size_t)b * a.Ch + h] = v;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float acc = 0.f;
    for (int h = 0; h < a.Ch; ++h) acc = fmaf(__ldg(a.w1 + (size_t)h * a.C + c), ga[h], acc);
    a.g_pool[(size_t)b * a.C + c] = acc;
  }
}

// one thread per (h, c) and batch chunk (blockIdx.y): dW1[h,c] += sum_b g_a1[b,h] zbar[b,c];  dW2[c,h] += sum_b g_pre[b,c] silu(a1[b,h]).
// The first version looped one thread over the whole batch: 225-320 us per launch at batch 512 whatever the layer size
// (a serial chain of L2 round trips); the batch is now split over up to 32 chunks that meet in fp32 atomics.
__global__ void __launch_bounds__(256) se_gen_wgrad_kernel(SeBwdArgs a, int b_chunk) {
  pdl_entry();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.C * a.Ch) return;
  const int h = idx / a.C, c = idx - h * a.C;
  const int b0 = blockIdx.y * b_chunk, b1 = min(a.B, b0 + b_chunk);
  float acc1 = 0.f, acc2 = 0.f, sb1 = 0.f, sb2 = 0.f;
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const float ga = __ldg(a.g_hid + (size_t)b * a.Ch + h), a1 = __ldg(a.hid + (size_t)b * a.Ch + h);
    const float zb = __ldg(a.zbar + (size_t)b * a.C + c), gpv = __ldg(a.g_pre + (size_t)b * a.C + c);
    acc1 = fmaf(ga, zb, acc1);
    acc2 = fmaf(gpv, a1 * seg_sigmoid(a1), acc2);
    sb1 += ga;
    sb2 += gpv;
  }
  atomicAdd(&a.dw1[(size_t)h * a.C + c], acc1);
  atomicAdd(&a.dw2[(size_t)c * a.Ch + h], acc2);
  if (c == 0) atomicAdd(&a.db1[h], sb1);
  if (h == 0) atomicAdd(&a.db2[c], sb2);
}

int launch_se_gen_fwd(const SeArgs& a, cudaStream_t st) {
  TD3D_REQUIRE(a.B > 0 && a.C > 0 && a.Ch > 0 && (a.C + a.Ch) * sizeof(float) <= 48 * 1024, "se_gen fwd: bad shape C=%d Ch=%d", a.C, a.Ch);
  TD3D_CUDA(launch_kernel(se_gen_fwd_kernel, a.B, 256, sizeof(float) * (a.C + a.Ch), st, a));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_se_gen_bwd(const SeBwdArgs& a, cudaStream_t st) {
  TD3D_REQUIRE(a.B > 0 && a.C > 0 && a.Ch > 0 && a.w1 && a.w2t, "se_gen bwd: bad arguments");
  TD3D_CUDA(launch_kernel(se_gen_bwd_kernel, a.B, 256, sizeof(float) * (a.C + a.Ch), st, a));
  TD3D_LAUNCH_CHECK();
  const int chunks = a.B < 32 ? a.B : 32;
  const int b_chunk = ceil_div(a.B, chunks);
  TD3D_CUDA(launch_kernel(se_gen_wgrad_kernel, dim3(ceil_div((int64_t)a.C * a.Ch, 256), ceil_div(a.B, b_chunk)), 256, 0, st, a, b_chunk));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
