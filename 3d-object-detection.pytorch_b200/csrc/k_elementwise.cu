// Bandwidth-bound elementwise kernels on NHWC activations (rows = pixels, 8-channel vectors).
//
// Thread mapping shared by all kernels here ("row loop"): a block works inside ONE sample b
// (blockIdx.y) on a chunk of its HW pixels; a thread owns one fixed 8-channel vector cv and walks
// pixels pl, pl+PL, ...  Per-channel constants (BN scale/shift, SE gate, alpha/beta/gamma) are
// loaded once per thread, global accesses of a warp are contiguous, and per-(b,c) reductions are
// register -> shared atomics -> one global atomic per channel per block.
//
// stats layout (all kernels): float [slots][2][C]; slot = b here.
#include "td3d_kernels.h"

namespace td3d {

static const int EW_THREADS = 256;
static const int EW_ITERS = 8;
static const int EW_U = 4;             // pixels per thread whose loads are in flight together

// 8 channels as loaded (kept packed until used: a batch of EW_U pixels x up to 3 tensors stays in registers)
template <typename T> struct RawV8;
template <> struct RawV8<bf16> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float v[8]) const {
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
    v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
  }
};
template <> struct RawV8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void get(float v[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// out = act(se*(scale*y+shift)) (+res);  pool (optional): stats[b][0][c] += sum_pixels out
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
apply_xform_kernel(const T* __restrict__ y, XForm xf, const T* __restrict__ res, T* __restrict__ out,
                   float* __restrict__ stats, int HW, int C, int pix_per_block) {
  pdl_entry();
  extern __shared__ float s_acc[];  // [C] when stats
  const int b = blockIdx.y;
  const ActK ak = make_actk(xf.act);
  if (stats) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
  }
  // channel vectors beyond blockDim (C/8 > 256) are covered by an outer loop
  for (int cv0 = 0; cv0 < (C >> 3); cv0 += blockDim.x) {
    int CV = min((int)blockDim.x, (C >> 3) - cv0);
    int PL = max(1, (int)blockDim.x / CV);
    int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    if (pl >= PL) continue;
    const int c = (cv0 + cv) << 3;
    float sc[8], sh[8], se[8], acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; se[i] = 1.f; acc[i] = 0.f; }
    if (xf.scale) { loadf8(xf.scale + c, sc); loadf8(xf.shift + c, sh); }
    if (xf.se) loadf8(xf.se + (size_t)b * C + c, se);
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(HW, p0 + pix_per_block);
    const size_t base = (size_t)b * HW * C + c;          // per-thread base; pixel offsets p*C fit 32 bits
    const T* yb = y + base;
    const T* rb = res ? res + base : nullptr;
    T* ob = out ? out + base : nullptr;
    for (int pb = p0 + pl; pb < p1; pb += PL * EW_U) {
      RawV8<T> rv[EW_U], rr[EW_U];
#pragma unroll
      for (int u = 0; u < EW_U; ++u) {
        const int p = pb + u * PL;
        if (p < p1) {
          const uint32_t off = (uint32_t)p * (uint32_t)C;
          rv[u].load(yb + off);
          if (res) rr[u].load(rb + off);
        }
      }
#pragma unroll
      for (int u = 0; u < EW_U; ++u) {
        const int p = pb + u * PL;
        if (p >= p1) continue;
        const uint32_t off = (uint32_t)p * (uint32_t)C;
        float v[8];
        rv[u].get(v);
        // branch-free activation from per-kernel constants: the `act` switch cost two uniform compare+branch
        // pairs per element (ncu: 30 % of the issued instructions)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[i] = xf.se_post ? actk_fwd(fmaf(v[i], sc[i], sh[i]), ak) * se[i] : actk_fwd(se[i] * fmaf(v[i], sc[i], sh[i]), ak);
        if (res) {
          float r[8];
          rr[u].get(r);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += r[i];
        }
        if (out) store8(ob + off, v);
        if (stats) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += out ? to_f(from_f<T>(v[i])) : v[i];   // pooled sums of what the consumer reads
        }
      }
    }
    if (stats) {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_acc[c + i], acc[i]);
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x)
      atomicAdd(&stats[((size_t)b * 2 + 0) * C + i], s_acc[i]);
  }
}

// out = alpha[b,c]*g + beta[c]*y + gamma[b,c]   (BatchNorm backward applied lazily)
template <typename T>
__global__ void __launch_bounds__(EW_THREADS)
affine2_kernel(const T* g, const T* __restrict__ y, const float* __restrict__ alpha,
               const float* __restrict__ beta, const float* __restrict__ gamma, T* out,
               int HW, int C, int pix_per_block) {
  pdl_entry();
  const int b = blockIdx.y;
  for (int cv0 = 0; cv0 < (C >> 3); cv0 += blockDim.x) {
    int CV = min((int)blockDim.x, (C >> 3) - cv0);
    int PL = max(1, (int)blockDim.x / CV);
    int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    if (pl >= PL) continue;
    const int c = (cv0 + cv) << 3;
    float al[8], be[8], ga[8];
    loadf8(alpha + (size_t)b * C + c, al);
    loadf8(beta + c, be);
    loadf8(gamma + (size_t)b * C + c, ga);
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(HW, p0 + pix_per_block);
    for (int p = p0 + pl; p < p1; p += PL) {
      const size_t off = ((size_t)b * HW + p) * C + c;
      float gv[8], yv[8];
      load8(g + off, gv);
      load8(y + off, yv);
#pragma unroll
      for (int i = 0; i < 8; ++i) gv[i] = fmaf(al[i], gv[i], fmaf(be[i], yv[i], ga[i]));
      store8(out + off, gv);
    }
  }
}

// gu = g * act'(u(y));  stats[b][0][c] += sum gu ; stats[b][1][c] += sum gu*y
// g_pooled != nullptr, g == nullptr: the incoming gradient is g_pooled[b,c] * g_scale for every pixel (avg-pool bwd)
// SE after the activation (xf.se_post, x = act(z)*se with z = scale*y+shift):
//   gu == nullptr            : statistics only, second sum = sum g*act(z)  (= d loss / d gate, what the SE backward needs)
//   gu != nullptr, g_pooled  : gu = (se*g + g_pooled*g_scale) * act'(z)    (gate path + squeeze path in one pass)
template <typename T>
__global__ void __launch_bounds__(EW_THREADS, 2)
act_bwd_stats_kernel(const T* g, const float* __restrict__ g_pooled, float g_scale,
                     const T* __restrict__ y, XForm xf, T* gu, float* __restrict__ stats,
                     const T* __restrict__ addend, int HW, int C, int pix_per_block) {
  pdl_entry();
  extern __shared__ float s_acc[];  // [2][C]
  const int b = blockIdx.y;
  const ActK ak = make_actk(xf.act);
  if (stats) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
  }
  for (int cv0 = 0; cv0 < (C >> 3); cv0 += blockDim.x) {
    int CV = min((int)blockDim.x, (C >> 3) - cv0);
    int PL = max(1, (int)blockDim.x / CV);
    int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    if (pl >= PL) continue;
    const int c = (cv0 + cv) << 3;
    float sc[8], sh[8], se[8], a1[8], a2[8], gp[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; se[i] = 1.f; a1[i] = 0.f; a2[i] = 0.f; gp[i] = 0.f; }
    if (xf.scale) { loadf8(xf.scale + c, sc); loadf8(xf.shift + c, sh); }
    if (xf.se) loadf8(xf.se + (size_t)b * C + c, se);
    if (g_pooled) {
      loadf8(g_pooled + (size_t)b * C + c, gp);
#pragma unroll
      for (int i = 0; i < 8; ++i) gp[i] *= g_scale;
    }
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(HW, p0 + pix_per_block);
    const size_t base = (size_t)b * HW * C + c;          // per-thread base; pixel offsets p*C fit 32 bits
    const T* yb = y + base;
    const T* gb = (g_pooled && !xf.se_post) ? nullptr : g + base;
    const T* ab = addend ? addend + base : nullptr;
    T* ub = gu ? gu + base : nullptr;
    // batches of EW_U pixels: all loads of a batch are issued before the first use (memory-level parallelism)
    for (int pb = p0 + pl; pb < p1; pb += PL * EW_U) {
      RawV8<T> rg[EW_U], ry[EW_U], ra[EW_U];
#pragma unroll
      for (int u = 0; u < EW_U; ++u) {
        const int p = pb + u * PL;
        if (p < p1) {
          const uint32_t off = (uint32_t)p * (uint32_t)C;
          ry[u].load(yb + off);
          if (gb) rg[u].load(gb + off);
          if (addend) ra[u].load(ab + off);
        }
      }
#pragma unroll
      for (int u = 0; u < EW_U; ++u) {
        const int p = pb + u * PL;
        if (p >= p1) continue;
        const uint32_t off = (uint32_t)p * (uint32_t)C;
        float gv[8], yv[8];
        ry[u].get(yv);
        if (!gb) {
#pragma unroll
          for (int i = 0; i < 8; ++i) gv[i] = gp[i];
        } else {
          rg[u].get(gv);
        }
        if (addend) {
          float ad[8];
          ra[u].get(ad);
#pragma unroll
          for (int i = 0; i < 8; ++i) gv[i] += ad[i];
        }
        if (!xf.se_post) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float uu = se[i] * fmaf(yv[i], sc[i], sh[i]);
            gv[i] *= actk_bwd(uu, ak);
          }
        } else if (ub) {
#pragma unroll
          for (int i = 0; i < 8; ++i) gv[i] = fmaf(se[i], gv[i], gp[i]) * actk_bwd(fmaf(yv[i], sc[i], sh[i]), ak);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) yv[i] = actk_fwd(fmaf(yv[i], sc[i], sh[i]), ak);   // second sum pairs g with act(z)
        }
        if (ub) store8(ub + off, gv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float r = ub ? to_f(from_f<T>(gv[i])) : gv[i];      // statistics of the stored (rounded) gradient
          a1[i] += r;
          a2[i] = fmaf(r, yv[i], a2[i]);
        }
      }
    }
    if (stats) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&s_acc[c + i], a1[i]);
        atomicAdd(&s_acc[C + c + i], a2[i]);
      }
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x)
      atomicAdd(&stats[(size_t)b * 2 * C + i], s_acc[i]);
  }
}

// out[b,c] = T(stats[b][0][c] * scale)   (avg-pool finalize)
template <typename T>
__global__ void pool_finalize_kernel(const float* __restrict__ stats, float scale, T* __restrict__ out,
                                     int B, int C) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int b = i / C, c = i % C;
  out[i] = from_f<T>(stats[((size_t)b * 2) * C + c] * scale);
}

// Pixels per block: long per-thread pixel loops amortise the per-block reduction (shared + global
// atomics per channel), but the grid must still fill the machine (>= ~4 blocks per SM).
static void ew_grid(int B, int HW, int C, dim3* grid, int* pix_per_block) {
  int CV = C >> 3;
  if (CV > EW_THREADS) CV = EW_THREADS;
  int PL = EW_THREADS / CV;
  if (PL < 1) PL = 1;
  int iters = 64;
  while (iters > EW_ITERS && (long)ceil_div(HW, PL * iters) * B < 148 * 2) iters >>= 1;
  int ppb = PL * iters;
  if (ppb > HW) ppb = HW;
  *pix_per_block = ppb;
  *grid = dim3((unsigned)ceil_div(HW, ppb), (unsigned)B);
}

int launch_apply_xform(const void* y, const XForm& xf, const void* res, void* out, float* pool_stats,
                       int B, int HW, int C, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(C % 8 == 0 && B > 0 && HW > 0 && (double)HW * C < 2147483648.0, "apply_xform: bad shape B=%d HW=%d C=%d", B, HW, C);
  dim3 grid; int ppb;
  ew_grid(B, HW, C, &grid, &ppb);
  size_t smem = pool_stats ? sizeof(float) * C : 0;
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(apply_xform_kernel<bf16>, grid, EW_THREADS, smem, st, (const bf16*)y, xf, (const bf16*)res, (bf16*)out, pool_stats, HW, C, ppb));
  else
    TD3D_CUDA(launch_kernel(apply_xform_kernel<float>, grid, EW_THREADS, smem, st, (const float*)y, xf, (const float*)res, (float*)out, pool_stats, HW, C, ppb));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_affine2(const void* g, const void* y, const float* alpha, const float* beta, const float* gamma,
                   void* out, int B, int HW, int C, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(C % 8 == 0 && B > 0 && HW > 0, "affine2: bad shape");
  dim3 grid; int ppb;
  ew_grid(B, HW, C, &grid, &ppb);
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(affine2_kernel<bf16>, grid, EW_THREADS, 0, st, (const bf16*)g, (const bf16*)y, alpha, beta, gamma, (bf16*)out, HW, C, ppb));
  else
    TD3D_CUDA(launch_kernel(affine2_kernel<float>, grid, EW_THREADS, 0, st, (const float*)g, (const float*)y, alpha, beta, gamma, (float*)out, HW, C, ppb));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_act_bwd_stats(const void* g, const float* g_pooled, float g_scale, const void* y, const XForm& xf,
                         void* gu, float* stats, int B, int HW, int C, int dtype, cudaStream_t st, const void* addend) {
  TD3D_REQUIRE(C % 8 == 0 && B > 0 && HW > 0 && (double)HW * C < 2147483648.0, "act_bwd_stats: bad shape");
  dim3 grid; int ppb;
  ew_grid(B, HW, C, &grid, &ppb);
  size_t smem = stats ? sizeof(float) * 2 * C : 0;
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(act_bwd_stats_kernel<bf16>, grid, EW_THREADS, smem, st, (const bf16*)g, g_pooled, g_scale, (const bf16*)y, xf, (bf16*)gu, stats, (const bf16*)addend, HW, C, ppb));
  else
    TD3D_CUDA(launch_kernel(act_bwd_stats_kernel<float>, grid, EW_THREADS, smem, st, (const float*)g, g_pooled, g_scale, (const float*)y, xf, (float*)gu, stats, (const float*)addend, HW, C, ppb));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_pool_finalize(const float* stats, float scale, void* out, int B, int C, int dtype, cudaStream_t st) {
  int n = B * C;
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(pool_finalize_kernel<bf16>, ceil_div(n, 256), 256, 0, st, stats, scale, (bf16*)out, B, C));
  else
    TD3D_CUDA(launch_kernel(pool_finalize_kernel<float>, ceil_div(n, 256), 256, 0, st, stats, scale, (float*)out, B, C));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
