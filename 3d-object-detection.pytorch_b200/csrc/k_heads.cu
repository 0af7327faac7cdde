// Regressor heads + classification head (reference ModelWrapper, torchdet3d/builders/
// model_builder.py:76-87 ctor, :126-146 forward, :112-124 forward_to_onnx).
//
//   kp[b]     = sigmoid(W_reg[cats[b]] f_b + b_reg[cats[b]])          (head picked per sample, :137)
//   logits[b] = W_cls (dropout(f_b)) + b_cls                          (Dropout(0.5) + Linear, :82-85)
//
// The reference runs B tiny GEMVs in a Python loop with B device->host syncs; here one block per
// sample does all 18+nc dot products, `cats` never leaves the device. Head parameters are read in
// fp32 straight from the parameter arena (reference [out,in] layout).
#include "td3d_kernels.h"

namespace td3d {

// Philox4x32-10 (counter = (b, c/4), key = seed) -> keep bit for Dropout(p=0.5)
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// seed may be advanced by a device-side step counter so that a CUDA-graph replay of the training
// step draws a fresh mask every iteration (the counter is the optimizer's global step)
__device__ __forceinline__ uint64_t effective_seed(const HeadsArgs& a) {
  return a.step_ptr ? a.seed + 0x9E3779B97F4A7C15ull * (uint64_t)(uint32_t)(*a.step_ptr) : a.seed;
}
__device__ __forceinline__ float dropout_keep(const float* keep, uint64_t seed, int training, int b, int c, int C) {
  if (!training) return 1.f;                    // eval: identity
  if (keep) return keep[(size_t)b * C + c] * 2.f;   // injected mask, scaled by 1/(1-p)
  uint4 r = philox4x32_10(make_uint4((uint32_t)b, (uint32_t)(c >> 2), 0u, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t w = (c & 3) == 0 ? r.x : ((c & 3) == 1 ? r.y : ((c & 3) == 2 ? r.z : r.w));
  return (w & 0x80000000u) ? 2.f : 0.f;
}

template <typename T>
__global__ void __launch_bounds__(128) heads_fwd_kernel(HeadsArgs a, const T* __restrict__ feat) {
  pdl_entry();
  extern __shared__ float s_f[];   // [C] features, [C] dropped features
  float* s_d = s_f + a.C;
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    float f = to_f(feat[(size_t)b * a.C + c]);
    s_f[c] = f;
    s_d[c] = f * dropout_keep(a.keep, effective_seed(a), a.training, b, c, a.C);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  int cat = (int)a.cats[b];
  cat = min(max(cat, 0), a.max_classes - 1);
  const float* wh = a.w_reg + (size_t)cat * a.reg_stride;
  for (int o = warp; o < a.P + a.nc; o += nwarp) {
    const bool is_kp = o < a.P;
    const float* wr = is_kp ? wh + (size_t)o * a.C : a.w_cls + (size_t)(o - a.P) * a.C;
    const float* x = is_kp ? s_f : s_d;
    float acc = 0.f;
    for (int c = lane; c < a.C; c += 32) acc = fmaf(wr[c], x[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (is_kp) {
        float z = acc + wh[(size_t)a.P * a.C + o];
        a.kp[(size_t)b * a.P + o] = 1.f / (1.f + expf(-z));
      } else {
        a.logits[(size_t)b * a.nc + (o - a.P)] = acc + a.b_cls[o - a.P];
      }
    }
  }
}

// export mode: all heads, kp_all[k][b][o]
template <typename T>
__global__ void __launch_bounds__(128) heads_all_kernel(HeadsArgs a, const T* __restrict__ feat, float* kp_all) {
  pdl_entry();
  extern __shared__ float s_f[];
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) s_f[c] = to_f(feat[(size_t)b * a.C + c]);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int n_kp = a.max_classes * a.P;
  for (int o = warp; o < n_kp + a.nc; o += nwarp) {
    const bool is_kp = o < n_kp;
    int k = o / a.P, j = o % a.P;
    const float* wh = a.w_reg + (size_t)k * a.reg_stride;
    const float* wr = is_kp ? wh + (size_t)j * a.C : a.w_cls + (size_t)(o - n_kp) * a.C;
    float acc = 0.f;
    for (int c = lane; c < a.C; c += 32) acc = fmaf(wr[c], s_f[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (is_kp) {
        float z = acc + wh[(size_t)a.P * a.C + j];
        kp_all[((size_t)k * a.B + b) * a.P + j] = 1.f / (1.f + expf(-z));
      } else {
        a.logits[(size_t)b * a.nc + (o - n_kp)] = acc + a.b_cls[o - n_kp];
      }
    }
  }
}

__global__ void select_argmax_kernel(const float* kp_all, const float* logits, float* kp_sel, int64_t* labels,
                                     int B, int P, int nc, int max_classes) {
  pdl_entry();
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int best = 0;
  float bv = logits[(size_t)b * nc];
  for (int n = 1; n < nc; ++n) {
    float v = logits[(size_t)b * nc + n];
    if (v > bv) { bv = v; best = n; }           // first maximum wins (torch.argmax / np.argmax)
  }
  labels[b] = best;
  int k = min(best, max_classes - 1);
  for (int j = 0; j < P; ++j) kp_sel[(size_t)b * P + j] = kp_all[((size_t)k * B + b) * P + j];
}

// ---- backward ---------------------------------------------------------------------------------
// g_pre[b,o] = d_kp[b,o]*kp*(1-kp);  g_feat[b,c] = W_reg[cat]^T g_pre + keep*(W_cls^T d_logits)
template <typename T>
__global__ void __launch_bounds__(128) heads_bwd_feat_kernel(HeadsBwdArgs a) {
  pdl_entry();
  __shared__ float s_g[64];   // [P] g_pre, then [nc] d_logits
  const HeadsArgs& f = a.f;
  const int b = blockIdx.x;
  if (threadIdx.x < f.P) {
    float k = f.kp[(size_t)b * f.P + threadIdx.x];
    float g = a.d_kp[(size_t)b * f.P + threadIdx.x] * k * (1.f - k);
    s_g[threadIdx.x] = g;
    a.g_pre[(size_t)b * f.P + threadIdx.x] = g;
  } else if (threadIdx.x < f.P + f.nc) {
    s_g[threadIdx.x] = a.d_logits[(size_t)b * f.nc + (threadIdx.x - f.P)];
  }
  __syncthreads();
  int cat = min(max((int)f.cats[b], 0), f.max_classes - 1);
  const float* wh = f.w_reg + (size_t)cat * f.reg_stride;
  for (int c = threadIdx.x; c < f.C; c += blockDim.x) {
    float acc = 0.f;
    for (int o = 0; o < f.P; ++o) acc = fmaf(wh[(size_t)o * f.C + c], s_g[o], acc);
    float accl = 0.f;
    for (int n = 0; n < f.nc; ++n) accl = fmaf(f.w_cls[(size_t)n * f.C + c], s_g[f.P + n], accl);
    a.g_feat[(size_t)b * f.C + c] = acc + accl * dropout_keep(f.keep, effective_seed(f), f.training, b, c, f.C);
  }
}

// weight gradients. grid = (C/32, max_classes + 1); blockIdx.y == max_classes is the cls head.
// lane = channel, the 8 warps split the samples (fixed partition -> deterministic sums), partials
// meet in shared memory.  A head whose class is absent from the batch gets present[k] = 0 and zero
// gradients.
static const int HW_WARPS = 8;
template <typename T>
__global__ void __launch_bounds__(32 * HW_WARPS) heads_bwd_wgrad_kernel(HeadsBwdArgs a, const T* __restrict__ feat) {
  pdl_entry();
  __shared__ float s_red[HW_WARPS][33][32];     // [warp][output row; 32 = bias][lane]
  __shared__ int s_tot[HW_WARPS];
  const HeadsArgs& f = a.f;
  const int k = blockIdx.y;
  const bool is_cls = k == f.max_classes;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int n_out = is_cls ? f.nc : f.P;
  const uint64_t seed = effective_seed(f);
  float acc[32];
  float accb = 0.f;
#pragma unroll
  for (int o = 0; o < 32; ++o) acc[o] = 0.f;
  int total = 0;
  for (int b = warp; b < f.B; b += HW_WARPS) {
    if (!is_cls && (int)f.cats[b] != k) continue;      // warp-uniform branch
    ++total;
    const float* g = is_cls ? a.d_logits + (size_t)b * f.nc : a.g_pre + (size_t)b * f.P;
    const float gl = lane < n_out ? g[lane] : 0.f;
    float x = 0.f;
    if (c < f.C) {
      x = to_f(feat[(size_t)b * f.C + c]);
      if (is_cls) x *= dropout_keep(f.keep, seed, f.training, b, c, f.C);
    }
#pragma unroll
    for (int o = 0; o < 32; ++o)
      if (o < n_out) acc[o] = fmaf(__shfl_sync(0xffffffffu, gl, o), x, acc[o]);
    accb += gl;
  }
#pragma unroll
  for (int o = 0; o < 32; ++o) s_red[warp][o][lane] = acc[o];
  s_red[warp][32][lane] = accb;
  if (lane == 0) s_tot[warp] = total;
  __syncthreads();
  float* dw = is_cls ? a.dw_cls : a.dw_reg + (size_t)k * f.reg_stride;
  for (int i = threadIdx.x; i < 33 * 32; i += 32 * HW_WARPS) {
    const int o = i >> 5, l = i & 31;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < HW_WARPS; ++w) s += s_red[w][o][l];
    if (o < 32) {
      const int cc = blockIdx.x * 32 + l;
      if (o < n_out && cc < f.C) dw[(size_t)o * f.C + cc] = s;
    } else if (blockIdx.x == 0 && l < n_out) {
      float* db = is_cls ? a.db_cls : dw + (size_t)f.P * f.C;
      db[l] = s;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && !is_cls) {
    int t = 0;
    for (int w = 0; w < HW_WARPS; ++w) t += s_tot[w];
    a.present[k] = t > 0 ? 1 : 0;
  }
}

int launch_heads_fwd(const HeadsArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.P + a.nc <= 64 && a.P <= 32 && a.nc <= 32, "heads: too many outputs (P=%d nc=%d)", a.P, a.nc);
  size_t smem = sizeof(float) * 2 * a.C;
  if (dtype == TD3D_BF16) TD3D_CUDA(launch_kernel(heads_fwd_kernel<bf16>, a.B, 128, smem, st, a, (const bf16*)a.feat));
  else TD3D_CUDA(launch_kernel(heads_fwd_kernel<float>, a.B, 128, smem, st, a, (const float*)a.feat));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_heads_all(const HeadsArgs& a, float* kp_all, int dtype, cudaStream_t st) {
  size_t smem = sizeof(float) * a.C;
  if (dtype == TD3D_BF16) TD3D_CUDA(launch_kernel(heads_all_kernel<bf16>, a.B, 128, smem, st, a, (const bf16*)a.feat, kp_all));
  else TD3D_CUDA(launch_kernel(heads_all_kernel<float>, a.B, 128, smem, st, a, (const float*)a.feat, kp_all));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_select_argmax(const float* kp_all, const float* logits, float* kp_sel, int64_t* labels, int B, int P,
                         int nc, int max_classes, cudaStream_t st) {
  TD3D_CUDA(launch_kernel(select_argmax_kernel, ceil_div(B, 128), 128, 0, st, kp_all, logits, kp_sel, labels, B, P, nc, max_classes));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_heads_bwd(const HeadsBwdArgs& a, int dtype, cudaStream_t st) {
  const HeadsArgs& f = a.f;
  TD3D_REQUIRE(f.P + f.nc <= 64 && f.P <= 32 && f.nc <= 32, "heads bwd: too many outputs");
  dim3 grid(ceil_div(f.C, 32), f.max_classes + 1);
  if (dtype == TD3D_BF16) {
    TD3D_CUDA(launch_kernel(heads_bwd_feat_kernel<bf16>, f.B, 128, 0, st, a));
    TD3D_LAUNCH_CHECK();
    TD3D_CUDA(launch_kernel(heads_bwd_wgrad_kernel<bf16>, grid, 32 * HW_WARPS, 0, st, a, (const bf16*)f.feat));
  } else {
    TD3D_CUDA(launch_kernel(heads_bwd_feat_kernel<float>, f.B, 128, 0, st, a));
    TD3D_LAUNCH_CHECK();
    TD3D_CUDA(launch_kernel(heads_bwd_wgrad_kernel<float>, grid, 32 * HW_WARPS, 0, st, a, (const float*)f.feat));
  }
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
