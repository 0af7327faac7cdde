// Squeeze-and-Excite FCs (reference SELayer, torchdet3d/models/mobilenetv3.py:92-107):
//   zbar = mean_HW(z);  hid = relu(W1 zbar + b1);  gate = h_sigmoid(W2 hid + b2);  x = z * gate
// The squeeze (per-(b,c) pixel sums) is produced by the depthwise-conv epilogue and the gate is
// applied by the consumer's load transform, so only the tiny [B,C]x[C,C/4] products remain.  They
// are batched over the whole mini-batch as fp32 GEMMs (the exact FFMA kernels of
// k_gemm_simple.cu) with small elementwise glue kernels -- everything stays fp32.
#include "td3d_kernels.h"

namespace td3d {

__global__ void se_zbar_kernel(SeArgs a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.C) return;
  int b = i / a.C, c = i % a.C;
  float z = a.pool_stats[((size_t)b * 2 + 0) * a.C + c] * a.inv_hw;
  if (a.scale) z = fmaf(z, a.scale[c], a.shift[c]);
  a.zbar[i] = z;
}
__global__ void se_gate_kernel(const float* __restrict__ pre, float* __restrict__ gate, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gate[i] = hsigmoid(pre[i]);
}

static GemmNT nt(const float* a, const float* w, float* y, const float* bias, int M, int N, int K, int relu) {
  GemmNT g = {};
  g.a = a; g.w = w; g.y = y; g.bias = bias; g.M = M; g.N = N; g.K = K; g.relu = relu;
  return g;
}

int launch_se_fwd(const SeArgs& a, cudaStream_t st) {
  const int n = a.B * a.C;
  se_zbar_kernel<<<ceil_div(n, 256), 256, 0, st>>>(a);
  TD3D_LAUNCH_CHECK();
  TD3D_TRY(launch_gemm_nt_simt(nt(a.zbar, a.w1, a.hid, a.b1, a.B, a.Ch, a.C, 1), TD3D_F32, st));
  TD3D_TRY(launch_gemm_nt_simt(nt(a.hid, a.w2, a.pre, a.b2, a.B, a.C, a.Ch, 0), TD3D_F32, st));
  se_gate_kernel<<<ceil_div(n, 256), 256, 0, st>>>(a.pre, a.gate, n);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// g_pre[b,c] = (sum_HW g_u * z)[b,c] * h_sigmoid'(pre)
__global__ void se_gpre_kernel(SeBwdArgs a) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.B * a.C) return;
  int b = i / a.C, c = i % a.C;
  float p1 = a.bwd_stats[((size_t)b * 2 + 0) * a.C + c];
  float p2 = a.bwd_stats[((size_t)b * 2 + 1) * a.C + c];
  float gs = a.scale ? fmaf(a.scale[c], p2, a.shift[c] * p1) : p2;
  a.g_pre[i] = gs * hsigmoid_bwd(a.pre[i]);
}
__global__ void se_relu_mask_kernel(const float* __restrict__ hid, float* __restrict__ g, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && !(hid[i] > 0.f)) g[i] = 0.f;
}
// out[c] = sum_b x[b,c]
__global__ void se_colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int C) {
  __shared__ float s[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float acc = 0.f;
  if (c < C)
    for (int b = w; b < B; b += 8) acc += x[(size_t)b * C + c];
  s[w][lane] = acc;
  __syncthreads();
  if (w == 0 && c < C) {
    for (int i = 1; i < 8; ++i) acc += s[i][lane];
    out[c] = acc;
  }
}

int launch_se_bwd(const SeBwdArgs& a, cudaStream_t st) {
  const int n = a.B * a.C, nh = a.B * a.Ch;
  se_gpre_kernel<<<ceil_div(n, 256), 256, 0, st>>>(a);
  TD3D_LAUNCH_CHECK();
  // g_hid = (g_pre W2) * [hid > 0]   (W operand = W2^T [Ch, C])
  TD3D_TRY(launch_gemm_nt_simt(nt(a.g_pre, a.w2t, a.g_hid, nullptr, a.B, a.Ch, a.C, 0), TD3D_F32, st));
  se_relu_mask_kernel<<<ceil_div(nh, 256), 256, 0, st>>>(a.hid, a.g_hid, nh);
  TD3D_LAUNCH_CHECK();
  // g_zbar = g_hid W1            (W operand = W1^T [C, Ch])
  TD3D_TRY(launch_gemm_nt_simt(nt(a.g_hid, a.w1t, a.g_pool, nullptr, a.B, a.C, a.Ch, 0), TD3D_F32, st));
  // weight gradients (the gradient arena is zeroed at the start of backward; TN accumulates)
  GemmTN t2 = {a.g_pre, a.hid, a.dw2, a.B, a.C, a.Ch};
  TD3D_TRY(launch_gemm_tn_simt(t2, TD3D_F32, st));
  GemmTN t1 = {a.g_hid, a.zbar, a.dw1, a.B, a.Ch, a.C};
  TD3D_TRY(launch_gemm_tn_simt(t1, TD3D_F32, st));
  se_colsum_kernel<<<ceil_div(a.C, 32), 256, 0, st>>>(a.g_pre, a.db2, a.B, a.C);
  TD3D_LAUNCH_CHECK();
  se_colsum_kernel<<<ceil_div(a.Ch, 32), 256, 0, st>>>(a.g_hid, a.db1, a.B, a.Ch);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
