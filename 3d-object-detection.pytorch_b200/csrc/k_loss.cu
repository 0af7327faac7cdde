// Fused loss forward+backward and metric accumulation.
//   losses  : torchdet3d/losses/regression_losses.py:8-58 (Diag, ADD, Wing), loss_builder.py:13-26
//             (L1 / SmoothL1 / MSE / CrossEntropy), weighted sum regression_losses.py:84-92
//   metrics : torchdet3d/evaluation/metrics.py:10-37 (ADD, symmetric ADD, accuracy) and the
//             per-class sums of :39-68
// The reference launches ~15 micro-kernels per loss and ~250 for the 9x9 metric loop (plus host
// syncs); here one warp owns one sample, everything stays in registers, one launch each.
#include "td3d_kernels.h"
#include <math_constants.h>

namespace td3d {

static const int LOSS_THREADS = 1024;
static const int NPT = 18;   // 9 keypoints x (x,y)   (model_builder.py:73 num_points=18)

__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

// min (or max) over lanes of equal parity with first-index tie break; inactive lanes carry +-inf
__device__ __forceinline__ void parity_argext(float& v, int& idx, bool want_max) {
#pragma unroll
  for (int o = 2; o <= 16; o <<= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    bool take = want_max ? (ov > v || (ov == v && oi < idx)) : (ov < v || (ov == v && oi < idx));
    if (take) { v = ov; idx = oi; }
  }
}

__global__ void __launch_bounds__(LOSS_THREADS)
loss_kernel(td3d_loss_desc d, const float* __restrict__ kp, const float* __restrict__ gt,
            const float* __restrict__ logits, const int64_t* __restrict__ cats, int B, int nc,
            float* __restrict__ loss_out, float* __restrict__ d_kp, float* __restrict__ d_logits) {
  pdl_entry();
  __shared__ double s_part[32][7];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const float inv_elems = 1.f / ((float)B * NPT), inv_b = 1.f / (float)B;
  const float wing_const = d.wing_w - d.wing_w * logf(1.f + d.wing_w / d.wing_eps);
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};   // l1, smoothl1, mse, add, diag, wing, ce (unweighted sums)
  for (int b = warp; b < B; b += nwarp) {
    const bool on = lane < NPT;
    float p = on ? kp[(size_t)b * NPT + lane] : 0.f;
    float g = on ? gt[(size_t)b * NPT + lane] : 0.f;
    float df = p - g, ad = fabsf(df), s = sgn(df);
    float grad = 0.f;
    float t_l1 = 0.f, t_sl1 = 0.f, t_mse = 0.f, t_add = 0.f, t_wing = 0.f;
    if (d.w_l1 != 0.f) { t_l1 = ad; grad += d.w_l1 * s * inv_elems; }
    if (d.w_mse != 0.f) { t_mse = df * df; grad += d.w_mse * 2.f * df * inv_elems; }
    if (d.w_smoothl1 != 0.f) {
      float be = d.smoothl1_beta;
      if (ad < be) { t_sl1 = 0.5f * df * df / be; grad += d.w_smoothl1 * (df / be) * inv_elems; }
      else { t_sl1 = ad - 0.5f * be; grad += d.w_smoothl1 * s * inv_elems; }
    }
    if (d.w_wing != 0.f) {
      float first, gcore;
      if (ad < d.wing_w) { first = d.wing_w * logf(1.f + ad / d.wing_eps); gcore = d.wing_w / (d.wing_eps + ad); }
      else { first = ad; gcore = 1.f; }
      t_wing = first >= d.wing_w ? first - wing_const : first;   // second masked update sees the rewritten value
      grad += d.w_wing * gcore * s * inv_elems;
    }
    float d_other = __shfl_xor_sync(0xffffffffu, df, 1);
    if (d.w_add != 0.f) {
      float n = sqrtf(df * df + d_other * d_other);
      t_add = (lane & 1) ? 0.f : n;                               // one norm per keypoint
      if (n > 0.f) grad += d.w_add * (df / n) * inv_b;
    }
    float t_diag = 0.f;
    if (d.w_diag != 0.f) {
      // bbox diagonal of pred and gt (compute_diag, regression_losses.py:51-58)
      float vmin = on ? p : CUDART_INF_F, vmax = on ? p : -CUDART_INF_F;
      int imin = lane, imax = lane;
      parity_argext(vmin, imin, false);
      parity_argext(vmax, imax, true);
      float gmin = on ? g : CUDART_INF_F, gmax = on ? g : -CUDART_INF_F;
      int j0 = lane, j1 = lane;
      parity_argext(gmin, j0, false);
      parity_argext(gmax, j1, true);
      float ext = vmax - vmin, gext = gmax - gmin;                 // even lanes: x extent, odd: y extent
      float ext_o = __shfl_xor_sync(0xffffffffu, ext, 1), gext_o = __shfl_xor_sync(0xffffffffu, gext, 1);
      float diag_p = sqrtf(ext * ext + ext_o * ext_o), diag_g = sqrtf(gext * gext + gext_o * gext_o);
      float dd = diag_p - diag_g, add_ = fabsf(dd);
      float gdd;
      if (add_ < 0.4f) { t_diag = 0.5f * dd * dd / 0.4f; gdd = dd / 0.4f; }
      else { t_diag = add_ - 0.2f; gdd = sgn(dd); }
      if (lane != 0) t_diag = 0.f;
      if (on && diag_p > 0.f) {
        float ge = d.w_diag * gdd * inv_b * (ext / diag_p);         // d loss / d extent of this lane's axis
        if (lane == imax) grad += ge;
        if (lane == imin) grad -= ge;
      }
    }
    if (d_kp && on) d_kp[(size_t)b * NPT + lane] = grad;
    float t_ce = 0.f;
    if (d.w_ce != 0.f) {
      const bool lon = lane < nc;
      float l = lon ? logits[(size_t)b * nc + lane] : -CUDART_INF_F;
      float m = warp_max(l);
      float e = lon ? expf(l - m) : 0.f;
      float se = warp_sum(e);
      int cat = (int)cats[b];
      float lcat = __shfl_sync(0xffffffffu, l, cat & 31);
      t_ce = lane == 0 ? (m + logf(se) - lcat) : 0.f;
      if (d_logits && lon) d_logits[(size_t)b * nc + lane] = d.w_ce * inv_b * (e / se - (lane == cat ? 1.f : 0.f));
    } else if (d_logits && lane < nc) {
      d_logits[(size_t)b * nc + lane] = 0.f;
    }
    float t[7] = {t_l1, t_sl1, t_mse, t_add, t_diag, t_wing, t_ce};
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      float v = warp_sum(t[i]);
      if (lane == 0) acc[i] += (double)v;
    }
  }
  if (lane == 0)
    for (int i = 0; i < 7; ++i) s_part[warp][i] = acc[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int w = 0; w < nwarp; ++w)
      for (int i = 0; i < 7; ++i) tot[i] += s_part[w][i];
    const double ne = (double)B * NPT, nb = (double)B;
    float l1 = (float)(tot[0] / ne) * d.w_l1, sl1 = (float)(tot[1] / ne) * d.w_smoothl1;
    float mse = (float)(tot[2] / ne) * d.w_mse, add = (float)(tot[3] / nb) * d.w_add;
    float diag = (float)(tot[4] / nb) * d.w_diag, wing = (float)(tot[5] / ne) * d.w_wing;
    float ce = (float)(tot[6] / nb) * d.w_ce;
    loss_out[1] = l1; loss_out[2] = sl1; loss_out[3] = mse; loss_out[4] = add;
    loss_out[5] = diag; loss_out[6] = wing; loss_out[7] = ce;
    loss_out[0] = l1 + sl1 + mse + add + diag + wing + ce;
  }
}

int launch_loss(const td3d_loss_desc& d, const float* kp, const float* gt, const float* logits, const int64_t* cats,
                int B, int nc, float* loss_out, float* d_kp, float* d_logits, cudaStream_t st) {
  TD3D_REQUIRE(B > 0 && nc >= 1 && nc <= 32, "loss: bad shape B=%d nc=%d", B, nc);
  TD3D_REQUIRE(d.w_ce == 0.f || (logits && cats), "loss: cross entropy needs logits and cats");
  TD3D_CUDA(launch_kernel(loss_kernel, 1, LOSS_THREADS, 0, st, d, kp, gt, logits, cats, B, nc, loss_out, d_kp, d_logits));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// acc[0..3] += sum_b ADD_b, sum_b SADD_b, hits, count ; acc[4+4k..] the same restricted to cats==k
__global__ void __launch_bounds__(256)
metrics_kernel(const float* __restrict__ kp, const float* __restrict__ gt, const float* __restrict__ logits,
               const int64_t* __restrict__ cats, int B, int nc, int max_classes, double* __restrict__ acc) {
  pdl_entry();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  float pv = lane < NPT ? kp[(size_t)b * NPT + lane] : 0.f;
  float gv = lane < NPT ? gt[(size_t)b * NPT + lane] : 0.f;
  const int i = lane < 9 ? lane : 0;
  float px = __shfl_sync(0xffffffffu, pv, 2 * i), py = __shfl_sync(0xffffffffu, pv, 2 * i + 1);
  float d_same = 0.f, d_min = CUDART_INF_F;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    float gx = __shfl_sync(0xffffffffu, gv, 2 * j), gy = __shfl_sync(0xffffffffu, gv, 2 * j + 1);
    float dx = px - gx, dy = py - gy;
    float dist = sqrtf(dx * dx + dy * dy);
    if (j == i) d_same = dist;
    d_min = fminf(d_min, dist);
  }
  if (lane >= 9) { d_same = 0.f; d_min = 0.f; }
  float add = warp_sum(d_same) * (1.f / 9.f), sadd = warp_sum(d_min) * (1.f / 9.f);
  if (lane == 0) {
    int cat = (int)cats[b];
    float hit = 0.f;
    if (logits) {
      int best = 0;
      float bv = logits[(size_t)b * nc];
      for (int n = 1; n < nc; ++n) {
        float v = logits[(size_t)b * nc + n];
        if (v > bv) { bv = v; best = n; }
      }
      hit = best == cat ? 1.f : 0.f;
    }
    atomicAdd(&acc[0], (double)add); atomicAdd(&acc[1], (double)sadd);
    atomicAdd(&acc[2], (double)hit); atomicAdd(&acc[3], 1.0);
    if (cat >= 0 && cat < max_classes) {
      double* a = acc + 4 + 4 * cat;
      atomicAdd(&a[0], (double)add); atomicAdd(&a[1], (double)sadd);
      atomicAdd(&a[2], (double)hit); atomicAdd(&a[3], 1.0);
    }
  }
}

int launch_metrics(const float* kp, const float* gt, const float* logits, const int64_t* cats, int B, int nc,
                   int max_classes, double* acc, cudaStream_t st) {
  TD3D_REQUIRE(B > 0, "metrics: empty batch");
  TD3D_CUDA(launch_kernel(metrics_kernel, ceil_div(B, 8), 256, 0, st, kp, gt, logits, cats, B, nc, max_classes, acc));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
