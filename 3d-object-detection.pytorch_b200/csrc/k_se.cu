// Squeeze-and-Excite FCs (reference SELayer, torchdet3d/models/mobilenetv3.py:92-107):
//   zbar = mean_HW(z);  hid = relu(W1 zbar + b1);  gate = h_sigmoid(W2 hid + b2);  x = z * gate
// The squeeze (per-(b,c) pixel sums) is produced by the depthwise-conv epilogue; the gate is
// applied by the consumer's load transform. These kernels are the tiny per-sample FCs only.
#include "td3d_kernels.h"

namespace td3d {

static const int SE_SB = 4;        // samples per block (weight rows are reused across them)
static const int SE_THREADS = 256;

__global__ void __launch_bounds__(SE_THREADS) se_fwd_kernel(SeArgs a) {
  extern __shared__ float sm[];
  float* s_z = sm;                     // [SE_SB][C]
  float* s_h = sm + SE_SB * a.C;       // [SE_SB][Ch]
  const int b0 = blockIdx.x * SE_SB;
  const int nb = min(SE_SB, a.B - b0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int i = threadIdx.x; i < nb * a.C; i += blockDim.x) {
    int s = i / a.C, c = i % a.C;
    float z = a.pool_stats[((size_t)(b0 + s) * 2 + 0) * a.C + c] * a.inv_hw;
    if (a.scale) z = fmaf(z, a.scale[c], a.shift[c]);
    s_z[s * a.C + c] = z;
    a.zbar[(size_t)(b0 + s) * a.C + c] = z;
  }
  __syncthreads();
  for (int j = warp; j < a.Ch; j += nwarp) {
    float acc[SE_SB];
#pragma unroll
    for (int s = 0; s < SE_SB; ++s) acc[s] = 0.f;
    const float* wr = a.w1 + (size_t)j * a.C;
    for (int c = lane; c < a.C; c += 32) {
      float w = wr[c];
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) acc[s] = fmaf(w, s_z[s * a.C + c], acc[s]);
    }
#pragma unroll
    for (int s = 0; s < SE_SB; ++s) acc[s] = warp_sum(acc[s]);
    if (lane == 0) {
      for (int s = 0; s < nb; ++s) {
        float h = fmaxf(acc[s] + a.b1[j], 0.f);
        s_h[s * a.Ch + j] = h;
        a.hid[(size_t)(b0 + s) * a.Ch + j] = h;
      }
    }
  }
  __syncthreads();
  for (int c = warp; c < a.C; c += nwarp) {
    float acc[SE_SB];
#pragma unroll
    for (int s = 0; s < SE_SB; ++s) acc[s] = 0.f;
    const float* wr = a.w2 + (size_t)c * a.Ch;
    for (int j = lane; j < a.Ch; j += 32) {
      float w = wr[j];
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) acc[s] = fmaf(w, s_h[s * a.Ch + j], acc[s]);
    }
#pragma unroll
    for (int s = 0; s < SE_SB; ++s) acc[s] = warp_sum(acc[s]);
    if (lane == 0) {
      for (int s = 0; s < nb; ++s) {
        float p = acc[s] + a.b2[c];
        a.pre[(size_t)(b0 + s) * a.C + c] = p;
        a.gate[(size_t)(b0 + s) * a.C + c] = hsigmoid(p);
      }
    }
  }
}

int launch_se_fwd(const SeArgs& a, cudaStream_t st) {
  size_t smem = sizeof(float) * SE_SB * (a.C + a.Ch);
  se_fwd_kernel<<<ceil_div(a.B, SE_SB), SE_THREADS, smem, st>>>(a);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// per-sample chain: g_gate -> g_pre -> g_hid -> g_zbar (=g_pool)
__global__ void __launch_bounds__(SE_THREADS) se_bwd_chain_kernel(SeBwdArgs a) {
  extern __shared__ float sm[];
  float* s_gp = sm;                    // [SE_SB][C]  g_pre
  float* s_gh = sm + SE_SB * a.C;      // [SE_SB][Ch] g_hid
  const int b0 = blockIdx.x * SE_SB;
  const int nb = min(SE_SB, a.B - b0);
  for (int i = threadIdx.x; i < nb * a.C; i += blockDim.x) {
    int s = i / a.C, c = i % a.C;
    size_t bc = (size_t)(b0 + s) * a.C + c;
    float p1 = a.bwd_stats[((size_t)(b0 + s) * 2 + 0) * a.C + c];
    float p2 = a.bwd_stats[((size_t)(b0 + s) * 2 + 1) * a.C + c];
    float gs = a.scale ? fmaf(a.scale[c], p2, a.shift[c] * p1) : p2;   // sum_HW g_u * z
    float gp = gs * hsigmoid_bwd(a.pre[bc]);
    s_gp[s * a.C + c] = gp;
    a.g_pre[bc] = gp;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < a.Ch; j += blockDim.x) {
    float acc[SE_SB];
#pragma unroll
    for (int s = 0; s < SE_SB; ++s) acc[s] = 0.f;
    for (int c = 0; c < a.C; ++c) {
      float w = a.w2[(size_t)c * a.Ch + j];
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) acc[s] = fmaf(w, s_gp[s * a.C + c], acc[s]);
    }
    for (int s = 0; s < nb; ++s) {
      size_t bj = (size_t)(b0 + s) * a.Ch + j;
      float gh = a.hid[bj] > 0.f ? acc[s] : 0.f;
      s_gh[s * a.Ch + j] = gh;
      a.g_hid[bj] = gh;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float acc[SE_SB];
#pragma unroll
    for (int s = 0; s < SE_SB; ++s) acc[s] = 0.f;
    for (int j = 0; j < a.Ch; ++j) {
      float w = a.w1[(size_t)j * a.C + c];
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) acc[s] = fmaf(w, s_gh[s * a.Ch + j], acc[s]);
    }
    for (int s = 0; s < nb; ++s) a.g_pool[(size_t)(b0 + s) * a.C + c] = acc[s];
  }
}

// weight gradients: dW2[c,j] = sum_b g_pre[b,c]*hid[b,j]; dW1[j,c] = sum_b g_hid[b,j]*zbar[b,c]
__global__ void se_bwd_wgrad_kernel(SeBwdArgs a) {
  const int n2 = a.C * a.Ch;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) {                               // dW2 [C,Ch], j fastest
    int c = i / a.Ch, j = i % a.Ch;
    float acc = 0.f;
    for (int b = 0; b < a.B; ++b) acc = fmaf(a.g_pre[(size_t)b * a.C + c], a.hid[(size_t)b * a.Ch + j], acc);
    a.dw2[i] = acc;
  } else if (i < 2 * n2) {                    // dW1 [Ch,C], c fastest
    int k = i - n2;
    int j = k / a.C, c = k % a.C;
    float acc = 0.f;
    for (int b = 0; b < a.B; ++b) acc = fmaf(a.g_hid[(size_t)b * a.Ch + j], a.zbar[(size_t)b * a.C + c], acc);
    a.dw1[k] = acc;
  } else if (i < 2 * n2 + a.C) {              // db2
    int c = i - 2 * n2;
    float acc = 0.f;
    for (int b = 0; b < a.B; ++b) acc += a.g_pre[(size_t)b * a.C + c];
    a.db2[c] = acc;
  } else if (i < 2 * n2 + a.C + a.Ch) {       // db1
    int j = i - 2 * n2 - a.C;
    float acc = 0.f;
    for (int b = 0; b < a.B; ++b) acc += a.g_hid[(size_t)b * a.Ch + j];
    a.db1[j] = acc;
  }
}

int launch_se_bwd(const SeBwdArgs& a, cudaStream_t st) {
  size_t smem = sizeof(float) * SE_SB * (a.C + a.Ch);
  se_bwd_chain_kernel<<<ceil_div(a.B, SE_SB), SE_THREADS, smem, st>>>(a);
  TD3D_LAUNCH_CHECK();
  int n = 2 * a.C * a.Ch + a.C + a.Ch;
  se_bwd_wgrad_kernel<<<ceil_div(n, 256), 256, 0, st>>>(a);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
