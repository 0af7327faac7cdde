// BatchNorm finalize kernels (reference: nn.BatchNorm2d/1d after every conv / classifier linear,
// torchdet3d/models/mobilenetv3.py:113,121,137,143,149,153,159,193).
//
// The conv kernels accumulate per-slot float partial sums [slots][2][C]; these kernels reduce the
// slots in double and emit the folded per-channel constants that consumers apply lazily:
//   forward : scale = gamma*invstd, shift = beta - mean*scale      (x_hat*gamma+beta == scale*y+shift)
//   backward: g_y = alpha[b,c]*g_u + beta[c]*y + gammac[b,c]        (BN backward incl. SE pooled path)
#include "td3d_kernels.h"

namespace td3d {

static const int BN_WARPS = 16;   // block = 16 warps x 32 channels; warps stride over the slots

// block: 32 consecutive channels (lane) x BN_WARPS slot-lanes (warp). Coalesced slot reads, double sums.
__global__ void __launch_bounds__(32 * BN_WARPS) bn_finalize_fwd_kernel(BnFwdArgs a) {
  pdl_entry();
  __shared__ double s_sum[2][BN_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  double s1 = 0.0, s2 = 0.0;
  if (c < a.C) {
    // independent loads first (the kernel is pure latency: a handful of CTAs, 2*slots/BN_WARPS loads each)
    const float* __restrict__ st = a.stats + c;
    const size_t C2 = 2 * (size_t)a.C;
    int s = w;
    for (; s + 7 * BN_WARPS < a.slots; s += 8 * BN_WARPS) {
      float v1[8], v2[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v1[u] = __ldg(st + (size_t)(s + u * BN_WARPS) * C2);
        v2[u] = __ldg(st + (size_t)(s + u * BN_WARPS) * C2 + a.C);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { s1 += (double)v1[u]; s2 += (double)v2[u]; }
    }
    for (; s < a.slots; s += BN_WARPS) {
      s1 += (double)__ldg(st + (size_t)s * C2);
      s2 += (double)__ldg(st + (size_t)s * C2 + a.C);
    }
  }
  s_sum[0][w][lane] = s1;
  s_sum[1][w][lane] = s2;
  __syncthreads();
  if (w != 0 || c >= a.C) return;
  s1 = 0.0; s2 = 0.0;
  for (int i = 0; i < BN_WARPS; ++i) { s1 += s_sum[0][i][lane]; s2 += s_sum[1][i][lane]; }
  double mean = s1 / a.count;
  double var = s2 / a.count - mean * mean;
  if (var < 0.0) var = 0.0;
  double invstd = 1.0 / sqrt(var + (double)a.eps);
  a.scale[c] = (float)((double)a.gamma[c] * invstd);
  a.shift[c] = (float)((double)a.beta[c] - mean * (double)a.gamma[c] * invstd);
  a.mean[c] = (float)mean;
  a.invstd[c] = (float)invstd;
  if (a.running_mean) {
    double unb = a.count > 1.0 ? var * (a.count / (a.count - 1.0)) : var;
    a.running_mean[c] = (float)((1.0 - a.momentum) * (double)a.running_mean[c] + a.momentum * mean);
    a.running_var[c] = (float)((1.0 - a.momentum) * (double)a.running_var[c] + a.momentum * unb);
    if (c == 0 && a.nbt) *a.nbt += 1;
  }
}

int launch_bn_finalize_fwd(const BnFwdArgs& a, cudaStream_t st) {
  TD3D_CUDA(launch_kernel(bn_finalize_fwd_kernel, ceil_div(a.C, 32), 32 * BN_WARPS, 0, st, a));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

__global__ void bn_eval_fold_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                                    float* scale, float* shift, int C, float eps) {
  pdl_entry();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // same operation order as ATen's eval-mode batch_norm: (x - mean) * (gamma / sqrt(var + eps)) + beta
  float invstd = 1.f / sqrtf(rv[c] + eps);
  float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - rm[c] * sc;
}

int launch_bn_eval_fold(const float* gamma, const float* beta, const float* rm, const float* rv, float* scale,
                        float* shift, int C, float eps, cudaStream_t st) {
  TD3D_CUDA(launch_kernel(bn_eval_fold_kernel, ceil_div(C, 128), 128, 0, st, gamma, beta, rm, rv, scale, shift, C, eps));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// z = a*y + b (a = gamma*invstd), u = se*z, x = act(u). Given per-(slot,c) sums P1 = sum g_u,
// P2 = sum g_u*y, the SE gate se[b,c], the gradient g_pool[b,c] w.r.t. the pooled mean of z and
// the forward pooled sums P0 = sum_HW y:
//   g_z = se*g_u + g_pool/HW
//   g_y = a*(g_z - mean_M(g_z) - x_hat*mean_M(g_z*x_hat))
//       = alpha[b,c]*g_u + beta[c]*y + gammac[b,c]
__global__ void __launch_bounds__(32 * BN_WARPS) bn_bwd_finalize_kernel(BnBwdArgs a) {
  pdl_entry();
  __shared__ double s_sum[2][BN_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const bool on = c < a.C;
  const double mu = on ? a.mean[c] : 0.0, is = on ? a.invstd[c] : 0.0;
  const double M = (double)a.B * (double)a.HW;
  const double inv_hw = 1.0 / (double)a.HW;         // no divisions inside the slot / sample loops
  const bool se_mode = a.se != nullptr;
  double S1 = 0.0, S2 = 0.0;
  if (on) {
    const float* __restrict__ st = a.stats + c;
    const float* __restrict__ sep = a.se ? a.se + c : nullptr;
    const float* __restrict__ gpp = a.g_pool ? a.g_pool + c : nullptr;
    const float* __restrict__ fpp = a.fwd_pool ? a.fwd_pool + c : nullptr;
    const size_t C1 = (size_t)a.C, C2 = 2 * C1;
    for (int s0 = w; s0 < a.slots; s0 += 8 * BN_WARPS) {
      float p1[8], p2[8], gate[8], gp[8], p0[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {                   // all loads of 8 slots in flight together
        const int s = s0 + u * BN_WARPS;
        const bool live = s < a.slots;
        p1[u] = live ? __ldg(st + (size_t)s * C2) : 0.f;
        p2[u] = live ? __ldg(st + (size_t)s * C2 + C1) : 0.f;
        gate[u] = 1.f; gp[u] = 0.f; p0[u] = 0.f;
        if (se_mode && live) {  // slots == B by construction
          gate[u] = __ldg(sep + (size_t)s * C1);
          gp[u] = __ldg(gpp + (size_t)s * C1);
          p0[u] = __ldg(fpp + (size_t)s * C2);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        S1 += (double)gate[u] * (double)p1[u] + (double)gp[u];
        S2 += ((double)gate[u] * ((double)p2[u] - mu * (double)p1[u]) + (double)gp[u] * ((double)p0[u] * inv_hw - mu)) * is;
      }
    }
  }
  s_sum[0][w][lane] = S1;
  s_sum[1][w][lane] = S2;
  __syncthreads();
  if (!on) return;
  S1 = 0.0; S2 = 0.0;
  for (int i = 0; i < BN_WARPS; ++i) { S1 += s_sum[0][i][lane]; S2 += s_sum[1][i][lane]; }
  const double c1 = S1 / M, c2 = S2 / M;
  const double aa = (double)a.gamma[c] * is;
  // gridDim.y CTAs repeat the (L2-resident) slot reduction and split the B per-sample writes
  const int b_chunk = (a.B + (int)gridDim.y - 1) / (int)gridDim.y;
  const int b_begin = (int)blockIdx.y * b_chunk, b_end = min(a.B, b_begin + b_chunk);
  if (w == 0 && blockIdx.y == 0) {
    a.beta[c] = (float)(-aa * c2 * is);
    a.dgamma[c] = (float)S2;
    a.dbeta[c] = (float)S1;
  }
  // per-sample constants in fp32 from fp64 per-channel terms (B*C cheap FMAs instead of B*C divisions)
  const float aa_f = (float)aa, g0_f = (float)(aa * (c2 * mu * is - c1)), ahw_f = (float)(aa * inv_hw);
  float* __restrict__ alpha_o = a.alpha + c;
  float* __restrict__ gammac_o = a.gammac + c;
  if (!se_mode) {
#pragma unroll 4
    for (int b = b_begin + w; b < b_end; b += BN_WARPS) {
      alpha_o[(size_t)b * a.C] = aa_f;
      gammac_o[(size_t)b * a.C] = g0_f;
    }
  } else {
    const float* __restrict__ sep = a.se + c;
    const float* __restrict__ gpp = a.g_pool + c;
    for (int b0 = b_begin + w; b0 < b_end; b0 += 8 * BN_WARPS) {
      float gate[8], gp[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + u * BN_WARPS;
        gate[u] = b < b_end ? __ldg(sep + (size_t)b * a.C) : 0.f;
        gp[u] = b < b_end ? __ldg(gpp + (size_t)b * a.C) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + u * BN_WARPS;
        if (b < b_end) {
          alpha_o[(size_t)b * a.C] = aa_f * gate[u];
          gammac_o[(size_t)b * a.C] = fmaf(ahw_f, gp[u], g0_f);
        }
      }
    }
  }
}

int launch_bn_bwd_finalize(const BnBwdArgs& a, cudaStream_t st) {
  TD3D_REQUIRE(!a.se || a.slots == a.B, "bn_bwd_finalize: SE mode needs slots == B");
  const int gy = a.B >= 64 ? 8 : 1;
  TD3D_CUDA(launch_kernel(bn_bwd_finalize_kernel, dim3(ceil_div(a.C, 32), gy), 32 * BN_WARPS, 0, st, a));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
