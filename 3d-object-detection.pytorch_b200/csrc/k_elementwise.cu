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
// pixels per thread whose loads are in flight together.  4 (two CTAs per SM at ~125 registers) beats 2 (three CTAs per SM):
// act_bwd_stats 0.99 vs 1.21 ms / step, apply_xform 0.71 vs 0.76 (MobileNetV3-large, B=256)
static const int EW_U = 4;

// ---- compile-time activations, 8 channels as four float2 (FFMA2 / FMUL2 / FADD2 issue two fp32 operations each) ----
// The activation used to be a run-time field applied through branch-free constants (ActK); with every mode of these
// kernels behind run-time flags the hot loop still carried ~130 instructions per 8-channel vector and ran at 60-65 % issue
// utilisation (profiles/r01_v3_ncu_full_elementwise.txt) -- issue-bound at 2.6-2.9 TB/s where the plain affine2 loop
// reaches 4.3.  Activation and SE placement are template parameters now; the bf16 rounding for the statistics reuses the
// packed store value instead of converting twice.
__device__ __forceinline__ float2 ew_fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 ew_mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 ew_add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
template <int ACT> __device__ __forceinline__ float ew_act(float u) {
  if (ACT == TD3D_ACT_RELU) return fmaxf(u, 0.f);
  if (ACT == TD3D_ACT_HSWISH) return u * __saturatef(fmaf(u, 1.f / 6.f, 0.5f));
  if (ACT == TD3D_ACT_SILU) return u * __fdividef(1.f, 1.f + __expf(-u));
  return u;
}
// d act(u) / du, matching autograd of x*relu6(x+3)/6 (hardtanh grad is 0 at both clamp points)
template <int ACT> __device__ __forceinline__ float ew_actd(float u) {
  if (ACT == TD3D_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (ACT == TD3D_ACT_HSWISH) return u <= -3.f ? 0.f : (u >= 3.f ? 1.f : fmaf(u, 1.f / 3.f, 0.5f));
  if (ACT == TD3D_ACT_SILU) { const float sg = __fdividef(1.f, 1.f + __expf(-u)); return sg * fmaf(u, 1.f - sg, 1.f); }
  return 1.f;
}
template <int ACT> __device__ __forceinline__ float2 ew_act2(float2 u) { return make_float2(ew_act<ACT>(u.x), ew_act<ACT>(u.y)); }
template <int ACT> __device__ __forceinline__ float2 ew_actd2(float2 u) { return make_float2(ew_actd<ACT>(u.x), ew_actd<ACT>(u.y)); }

template <typename T> struct Vec8;            // 8 channels in the activation dtype <-> four float2
template <> struct Vec8<bf16> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float2 v[4]) const {
    v[0] = make_float2(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u));
    v[1] = make_float2(__uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
    v[2] = make_float2(__uint_as_float(r.z << 16), __uint_as_float(r.z & 0xffff0000u));
    v[3] = make_float2(__uint_as_float(r.w << 16), __uint_as_float(r.w & 0xffff0000u));
  }
  // round v to the storage dtype (v keeps the rounded values: the statistics describe what the consumer reads)
  __device__ __forceinline__ void set(float2 v[4]) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[i].x, v[i].y);
    get(v);
  }
  __device__ __forceinline__ void store(bf16* p) const { *reinterpret_cast<uint4*>(p) = r; }
};
template <> struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void get(float2 v[4]) const {
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
  }
  __device__ __forceinline__ void set(float2 v[4]) {
    a = make_float4(v[0].x, v[0].y, v[1].x, v[1].y); b = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
  }
  __device__ __forceinline__ void store(float* p) const {
    reinterpret_cast<float4*>(p)[0] = a; reinterpret_cast<float4*>(p)[1] = b;
  }
};
__device__ __forceinline__ void ew_ldc8(const float* p, float2 v[4]) {      // 8 fp32 per-channel constants
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w); v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}

// out = act(se*(scale*y+shift)) (+res), or with SE_POST act(scale*y+shift)*se;  pool (optional): stats[b][0][c] += sum_pixels out
template <typename T, int ACT, bool SE_POST>
__global__ void __launch_bounds__(EW_THREADS, 2)
apply_xform_kernel(const T* __restrict__ y, XForm xf, const T* __restrict__ res, T* __restrict__ out,
                   float* __restrict__ stats, int HW, int C, int pix_per_block) {
  pdl_entry();
  extern __shared__ float s_acc[];  // [C] when stats
  const int b = blockIdx.y;
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
    float2 sc[4], sh[4], se[4], acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { sc[i] = make_float2(1.f, 1.f); sh[i] = make_float2(0.f, 0.f); se[i] = sc[i]; acc[i] = sh[i]; }
    if (xf.scale) { ew_ldc8(xf.scale + c, sc); ew_ldc8(xf.shift + c, sh); }
    if (xf.se) ew_ldc8(xf.se + (size_t)b * C + c, se);
    if (!SE_POST) {                          // se*(scale*y+shift) = (se*scale)*y + se*shift
#pragma unroll
      for (int i = 0; i < 4; ++i) { sc[i] = ew_mul2(sc[i], se[i]); sh[i] = ew_mul2(sh[i], se[i]); }
    }
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(HW, p0 + pix_per_block);
    const size_t base = (size_t)b * HW * C + c;          // per-thread base; pixel offsets p*C fit 32 bits
    const T* yb = y + base;
    const T* rb = res ? res + base : nullptr;
    T* ob = out ? out + base : nullptr;
    for (int pb = p0 + pl; pb < p1; pb += PL * EW_U) {
      Vec8<T> rv[EW_U], rr[EW_U];
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
        float2 v[4];
        rv[u].get(v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = ew_act2<ACT>(ew_fma2(v[i], sc[i], sh[i]));
          if (SE_POST) v[i] = ew_mul2(v[i], se[i]);
        }
        if (res) {
          float2 r[4];
          rr[u].get(r);
#pragma unroll
          for (int i = 0; i < 4; ++i) v[i] = ew_add2(v[i], r[i]);
        }
        if (out) {
          Vec8<T> o;
          o.set(v);                          // pooled sums of what the consumer reads
          o.store(ob + off);
        }
        if (stats) {
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = ew_add2(acc[i], v[i]);
        }
      }
    }
    if (stats) {
#pragma unroll
      for (int i = 0; i < 4; ++i) { atomicAdd(&s_acc[c + 2 * i], acc[i].x); atomicAdd(&s_acc[c + 2 * i + 1], acc[i].y); }
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = threadIdx.x * 4; i < C; i += blockDim.x * 4)      // C % 8 == 0: 16-byte vector reductions
      atomicAdd(reinterpret_cast<float4*>(&stats[((size_t)b * 2 + 0) * C + i]), make_float4(s_acc[i], s_acc[i + 1], s_acc[i + 2], s_acc[i + 3]));
  }
}

// out = alpha[b,c]*g + beta[c]*y + gamma[b,c]   (BatchNorm backward applied lazily)
template <typename T>
__global__ void __launch_bounds__(EW_THREADS, 3)
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
    float2 al[4], be[4], ga[4];
    ew_ldc8(alpha + (size_t)b * C + c, al);
    ew_ldc8(beta + c, be);
    ew_ldc8(gamma + (size_t)b * C + c, ga);
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(HW, p0 + pix_per_block);
    const size_t base = (size_t)b * HW * C + c;          // per-thread base; pixel offsets p*C fit 32 bits
    const T* gb = g + base;
    const T* yb = y + base;
    T* ob = out + base;
    for (int pb = p0 + pl; pb < p1; pb += PL * EW_U) {
      Vec8<T> rg[EW_U], ry[EW_U];
#pragma unroll
      for (int u = 0; u < EW_U; ++u) {
        const int p = pb + u * PL;
        if (p < p1) {
          const uint32_t off = (uint32_t)p * (uint32_t)C;
          rg[u].load(gb + off);
          ry[u].load(yb + off);
        }
      }
#pragma unroll
      for (int u = 0; u < EW_U; ++u) {
        const int p = pb + u * PL;
        if (p >= p1) continue;
        float2 gv[4], yv[4];
        rg[u].get(gv);
        ry[u].get(yv);
#pragma unroll
        for (int i = 0; i < 4; ++i) gv[i] = ew_fma2(al[i], gv[i], ew_fma2(be[i], yv[i], ga[i]));
        Vec8<T> o;
        o.set(gv);
        o.store(ob + (uint32_t)p * (uint32_t)C);
      }
    }
  }
}

// gu = g * act'(u(y));  stats[b][0][c] += sum gu ; stats[b][1][c] += sum gu*y
// g_pooled != nullptr, g == nullptr: the incoming gradient is g_pooled[b,c] * g_scale for every pixel (avg-pool bwd)
// MODE (SE after the activation, xf.se_post: x = act(z)*se with z = scale*y+shift):
//   EW_PLAIN      : u = se*(scale*y+shift), gu = g*act'(u)
//   EW_POST_GRAD  : gu = (se*g + g_pooled*g_scale) * act'(z)               (gate path + squeeze path in one pass)
//   EW_POST_STATS : gu == nullptr, statistics only, second sum = sum g*act(z)  (= d loss / d gate, what the SE backward needs)
enum { EW_PLAIN = 0, EW_POST_GRAD = 1, EW_POST_STATS = 2 };
template <typename T, int ACT, int MODE>
__global__ void __launch_bounds__(EW_THREADS, 2)
act_bwd_stats_kernel(const T* g, const float* __restrict__ g_pooled, float g_scale,
                     const T* __restrict__ y, XForm xf, T* gu, float* __restrict__ stats,
                     const T* __restrict__ addend, int HW, int C, int pix_per_block) {
  pdl_entry();
  extern __shared__ float s_acc[];  // [2][C]
  const int b = blockIdx.y;
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
    float2 sc[4], sh[4], se[4], a1[4], a2[4], gp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sc[i] = make_float2(1.f, 1.f); sh[i] = make_float2(0.f, 0.f); se[i] = sc[i]; a1[i] = sh[i]; a2[i] = sh[i]; gp[i] = sh[i];
    }
    if (xf.scale) { ew_ldc8(xf.scale + c, sc); ew_ldc8(xf.shift + c, sh); }
    if (xf.se) ew_ldc8(xf.se + (size_t)b * C + c, se);
    if (MODE == EW_PLAIN) {                  // se*(scale*y+shift) = (se*scale)*y + se*shift
#pragma unroll
      for (int i = 0; i < 4; ++i) { sc[i] = ew_mul2(sc[i], se[i]); sh[i] = ew_mul2(sh[i], se[i]); }
    }
    if (g_pooled) {
      ew_ldc8(g_pooled + (size_t)b * C + c, gp);
#pragma unroll
      for (int i = 0; i < 4; ++i) gp[i] = ew_mul2(gp[i], make_float2(g_scale, g_scale));
    }
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(HW, p0 + pix_per_block);
    const size_t base = (size_t)b * HW * C + c;          // per-thread base; pixel offsets p*C fit 32 bits
    const T* yb = y + base;
    const T* gb = (g_pooled && MODE == EW_PLAIN) ? nullptr : g + base;
    const T* ab = addend ? addend + base : nullptr;
    T* ub = gu ? gu + base : nullptr;
    // batches of EW_U pixels: all loads of a batch are issued before the first use (memory-level parallelism)
    for (int pb = p0 + pl; pb < p1; pb += PL * EW_U) {
      Vec8<T> rg[EW_U], ry[EW_U], ra[EW_U];
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
        float2 gv[4], yv[4];
        ry[u].get(yv);
        if (!gb) {
#pragma unroll
          for (int i = 0; i < 4; ++i) gv[i] = gp[i];
        } else {
          rg[u].get(gv);
        }
        if (addend) {
          float2 ad[4];
          ra[u].get(ad);
#pragma unroll
          for (int i = 0; i < 4; ++i) gv[i] = ew_add2(gv[i], ad[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 z = ew_fma2(yv[i], sc[i], sh[i]);
          if (MODE == EW_PLAIN) gv[i] = ew_mul2(gv[i], ew_actd2<ACT>(z));
          else if (MODE == EW_POST_GRAD) gv[i] = ew_mul2(ew_fma2(se[i], gv[i], gp[i]), ew_actd2<ACT>(z));
          else yv[i] = ew_act2<ACT>(z);                  // second sum pairs g with act(z)
        }
        if (ub) {
          Vec8<T> o;
          o.set(gv);                                     // statistics of the stored (rounded) gradient
          o.store(ub + off);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a1[i] = ew_add2(a1[i], gv[i]);
          a2[i] = ew_fma2(gv[i], yv[i], a2[i]);
        }
      }
    }
    if (stats) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        atomicAdd(&s_acc[c + 2 * i], a1[i].x);
        atomicAdd(&s_acc[c + 2 * i + 1], a1[i].y);
        atomicAdd(&s_acc[C + c + 2 * i], a2[i].x);
        atomicAdd(&s_acc[C + c + 2 * i + 1], a2[i].y);
      }
    }
  }
  if (stats) {
    __syncthreads();
    for (int i = threadIdx.x * 4; i < 2 * C; i += blockDim.x * 4)      // C % 8 == 0: 16-byte vector reductions
      atomicAdd(reinterpret_cast<float4*>(&stats[(size_t)b * 2 * C + i]), make_float4(s_acc[i], s_acc[i + 1], s_acc[i + 2], s_acc[i + 3]));
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

template <typename T, int ACT>
static int apply_xform_launch(const void* y, const XForm& xf, const void* res, void* out, float* pool_stats, int B, int HW, int C,
                              cudaStream_t st) {
  dim3 grid; int ppb;
  ew_grid(B, HW, C, &grid, &ppb);
  const size_t smem = pool_stats ? sizeof(float) * C : 0;
#define TD3D_AX_LAUNCH(POST)                                                                                          \
  TD3D_CUDA(launch_kernel(apply_xform_kernel<T, ACT, POST>, grid, EW_THREADS, smem, st, (const T*)y, xf, (const T*)res, \
                          (T*)out, pool_stats, HW, C, ppb))
  if (xf.se_post && xf.se) TD3D_AX_LAUNCH(true);
  else TD3D_AX_LAUNCH(false);
#undef TD3D_AX_LAUNCH
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}
template <typename T>
static int apply_xform_act(const void* y, const XForm& xf, const void* res, void* out, float* pool_stats, int B, int HW, int C,
                           cudaStream_t st) {
  switch (xf.act) {
    case TD3D_ACT_NONE: return apply_xform_launch<T, TD3D_ACT_NONE>(y, xf, res, out, pool_stats, B, HW, C, st);
    case TD3D_ACT_RELU: return apply_xform_launch<T, TD3D_ACT_RELU>(y, xf, res, out, pool_stats, B, HW, C, st);
    case TD3D_ACT_HSWISH: return apply_xform_launch<T, TD3D_ACT_HSWISH>(y, xf, res, out, pool_stats, B, HW, C, st);
    case TD3D_ACT_SILU: return apply_xform_launch<T, TD3D_ACT_SILU>(y, xf, res, out, pool_stats, B, HW, C, st);
  }
  set_last_error("apply_xform: unknown activation %d", xf.act);
  return TD3D_EINVAL;
}

int launch_apply_xform(const void* y, const XForm& xf, const void* res, void* out, float* pool_stats,
                       int B, int HW, int C, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(C % 8 == 0 && B > 0 && HW > 0 && (double)HW * C < 2147483648.0, "apply_xform: bad shape B=%d HW=%d C=%d", B, HW, C);
  return dtype == TD3D_BF16 ? apply_xform_act<bf16>(y, xf, res, out, pool_stats, B, HW, C, st)
                            : apply_xform_act<float>(y, xf, res, out, pool_stats, B, HW, C, st);
}

int launch_affine2(const void* g, const void* y, const float* alpha, const float* beta, const float* gamma,
                   void* out, int B, int HW, int C, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(C % 8 == 0 && B > 0 && HW > 0 && (double)HW * C < 2147483648.0, "affine2: bad shape B=%d HW=%d C=%d", B, HW, C);
  dim3 grid; int ppb;
  ew_grid(B, HW, C, &grid, &ppb);
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(affine2_kernel<bf16>, grid, EW_THREADS, 0, st, (const bf16*)g, (const bf16*)y, alpha, beta, gamma, (bf16*)out, HW, C, ppb));
  else
    TD3D_CUDA(launch_kernel(affine2_kernel<float>, grid, EW_THREADS, 0, st, (const float*)g, (const float*)y, alpha, beta, gamma, (float*)out, HW, C, ppb));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

template <typename T, int ACT>
static int act_bwd_launch(const void* g, const float* g_pooled, float g_scale, const void* y, const XForm& xf, void* gu, float* stats,
                          int B, int HW, int C, cudaStream_t st, const void* addend) {
  dim3 grid; int ppb;
  ew_grid(B, HW, C, &grid, &ppb);
  const size_t smem = stats ? sizeof(float) * 2 * C : 0;
  const int mode = !xf.se_post ? EW_PLAIN : (gu ? EW_POST_GRAD : EW_POST_STATS);
  TD3D_REQUIRE(mode == EW_PLAIN || g, "act_bwd_stats: SE-after-activation modes need the gradient tensor");
#define TD3D_ABS_LAUNCH(MODE)                                                                                               \
  TD3D_CUDA(launch_kernel(act_bwd_stats_kernel<T, ACT, MODE>, grid, EW_THREADS, smem, st, (const T*)g, g_pooled, g_scale, \
                          (const T*)y, xf, (T*)gu, stats, (const T*)addend, HW, C, ppb))
  if (mode == EW_PLAIN) TD3D_ABS_LAUNCH(EW_PLAIN);
  else if (mode == EW_POST_GRAD) TD3D_ABS_LAUNCH(EW_POST_GRAD);
  else TD3D_ABS_LAUNCH(EW_POST_STATS);
#undef TD3D_ABS_LAUNCH
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}
template <typename T>
static int act_bwd_act(const void* g, const float* g_pooled, float g_scale, const void* y, const XForm& xf, void* gu, float* stats,
                       int B, int HW, int C, cudaStream_t st, const void* addend) {
  switch (xf.act) {
    case TD3D_ACT_NONE: return act_bwd_launch<T, TD3D_ACT_NONE>(g, g_pooled, g_scale, y, xf, gu, stats, B, HW, C, st, addend);
    case TD3D_ACT_RELU: return act_bwd_launch<T, TD3D_ACT_RELU>(g, g_pooled, g_scale, y, xf, gu, stats, B, HW, C, st, addend);
    case TD3D_ACT_HSWISH: return act_bwd_launch<T, TD3D_ACT_HSWISH>(g, g_pooled, g_scale, y, xf, gu, stats, B, HW, C, st, addend);
    case TD3D_ACT_SILU: return act_bwd_launch<T, TD3D_ACT_SILU>(g, g_pooled, g_scale, y, xf, gu, stats, B, HW, C, st, addend);
  }
  set_last_error("act_bwd_stats: unknown activation %d", xf.act);
  return TD3D_EINVAL;
}

int launch_act_bwd_stats(const void* g, const float* g_pooled, float g_scale, const void* y, const XForm& xf,
                         void* gu, float* stats, int B, int HW, int C, int dtype, cudaStream_t st, const void* addend) {
  TD3D_REQUIRE(C % 8 == 0 && B > 0 && HW > 0 && (double)HW * C < 2147483648.0, "act_bwd_stats: bad shape");
  return dtype == TD3D_BF16 ? act_bwd_act<bf16>(g, g_pooled, g_scale, y, xf, gu, stats, B, HW, C, st, addend)
                            : act_bwd_act<float>(g, g_pooled, g_scale, y, xf, gu, stats, B, HW, C, st, addend);
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
