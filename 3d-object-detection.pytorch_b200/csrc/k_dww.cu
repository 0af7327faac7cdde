// Depthwise k x k convolution, "row walker" generation for SMALL planes (W <= 32, stride 1, NHWC).
// Reference op: nn.Conv2d(hidden, hidden, k, 1, (k-1)//2, groups=hidden) inside InvertedResidual,
// torchdet3d/models/mobilenetv3.py:136,152, and its autograd backward.
//
// ncu on the tiled kernels of k_dw2.cu showed the late layers (14x14 / 7x7 planes, 184..960
// channels) running at 0.3-0.8 TB/s: a 14x14 plane gives a 256-thread CTA only ~24 elements per
// thread and tile, so the three CTA barriers, the halo (2.5x staged pixels for a 5x5 on 7x7) and
// the index arithmetic of every tile dominate (64-216 issued instructions per element at an IPC
// of 0.25-0.4 per scheduler).  For planes this small a whole image row fits the lanes of a warp:
//   * lane = (x, v): x = pixel column (XL = 8/16/32 lanes), v = 8-channel (16 B) vector; a warp owns
//     one (sample, 32/XL vectors) plane at a time and WALKS ITS ROWS top to bottom;
//   * each input element is loaded ONCE with a 16-byte load, transformed once in registers
//     (BatchNorm fold + SE gate + activation);
//     horizontal neighbours come from warp shuffles, vertical reuse from K rolling accumulator
//     rows held in registers -- no shared-memory tile, no CTA barrier in the main loop;
//   * rows are prefetched 4-8 deep with cp.async into a per-warp shared-memory ring (each lane reads
//     back only what it copied itself: no barrier), and the stream runs on into the next plane; the
//     first version prefetched into registers and measured 2850 cycles per row step because the
//     ring's register moves waited on the loads in flight;
//   * statistics (BatchNorm sums / SE squeeze) stay in registers for the whole plane and leave the
//     warp as one atomic per (sample, channel).
//
//   forward   y  = dw(act(se*(scale*x+shift)))                      + sum y,  sum y^2   per (b,c)
//
// Forward only: the backward twins of this kernel (round 1) lost to the one-pass column walker
// (dwc_core.cuh) on every layer once that kernel's inner loop was cleaned up (profiles/r02_v5_dw_bench.txt)
// and were removed.
#include "td3d_kernels.h"

#include <stdlib.h>

namespace td3d {

namespace {

constexpr int WW_WARPS = 4;
constexpr int WW_THREADS = 32 * WW_WARPS;
struct WwArgs {
  const void* s0;            // x (raw forward input)
  XForm xf;                  // lazily applied transform of x
  const float* w;            // [K*K][C] fp32 taps
  void* out;                 // y
  float* stats;              // [B][2][C] or null
  int B, H, W, C;
  int xl_log2;               // lanes per image row = 1 << xl_log2 (>= W)
};

// ---- NV channels of one pixel as loaded from global memory (NV = 8: 16 B bf16 / 32 B fp32) ----
__device__ __forceinline__ uint32_t ww_s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ww_cp16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ww_cp8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void ww_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ww_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T, int NV> struct WwRaw;
template <> struct WwRaw<bf16, 8> {
  uint4 r;
  __device__ __forceinline__ void zero() { r = make_uint4(0u, 0u, 0u, 0u); }
  static constexpr int BYTES = 16;
  static __device__ __forceinline__ void fetch(uint32_t dst, const bf16* p) { ww_cp16(dst, p); }
  __device__ __forceinline__ void lds(const uint8_t* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void get(float2 (&v)[4]) const {
    v[0] = make_float2(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u));
    v[1] = make_float2(__uint_as_float(r.y << 16), __uint_as_float(r.y & 0xffff0000u));
    v[2] = make_float2(__uint_as_float(r.z << 16), __uint_as_float(r.z & 0xffff0000u));
    v[3] = make_float2(__uint_as_float(r.w << 16), __uint_as_float(r.w & 0xffff0000u));
  }
};
template <> struct WwRaw<float, 8> {
  float4 a, b;
  __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
  static constexpr int BYTES = 32;
  static __device__ __forceinline__ void fetch(uint32_t dst, const float* p) { ww_cp16(dst, p); ww_cp16(dst + 16, p + 4); }
  __device__ __forceinline__ void lds(const uint8_t* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 16);
  }
  __device__ __forceinline__ void get(float2 (&v)[4]) const {
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
    v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
  }
};
// store 8 channels in the activation dtype; r = the values as stored (rounded) for the statistics
__device__ __forceinline__ void ww_store8(bf16* p, const float2 (&v)[4], float2 (&r)[4]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[i].x, v[i].y);
  *reinterpret_cast<uint4*>(p) = raw;
  r[0] = make_float2(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u));
  r[1] = make_float2(__uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
  r[2] = make_float2(__uint_as_float(raw.z << 16), __uint_as_float(raw.z & 0xffff0000u));
  r[3] = make_float2(__uint_as_float(raw.w << 16), __uint_as_float(raw.w & 0xffff0000u));
}
__device__ __forceinline__ void ww_store8(float* p, const float2 (&v)[4], float2 (&r)[4]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
#pragma unroll
  for (int i = 0; i < 4; ++i) r[i] = v[i];
}

// value of lane (lane + delta) of the SAME image row, 0 outside the row.  Without SEL the caller
// guarantees W + P <= XL, so the lanes the rotation wraps into are padding lanes holding zeros.
template <bool SEL>
__device__ __forceinline__ float2 ww_shift(const float2& v, int delta, int x, int W) {
  const int src = ((int)threadIdx.x + delta) & 31;
  float2 t;
  t.x = __shfl_sync(0xffffffffu, v.x, src);
  t.y = __shfl_sync(0xffffffffu, v.y, src);
  if (SEL) {
    const bool ok = (unsigned)(x + delta) < (unsigned)W;
    t.x = ok ? t.x : 0.f;
    t.y = ok ? t.y : 0.f;
  }
  return t;
}

// per-channel constants in shared memory (written once per CTA): [which][channel of the warp's span]
enum { WK_SC = 0, WK_SH, WK_N };
// per-(sample, channel) constants, double buffered per warp and prefetched one plane ahead with cp.async
enum { WP_SE = 0, WP_N };

template <typename T> struct WwDepth {          // rows in flight per warp
  static constexpr int FWD = sizeof(T) == 2 ? 8 : 4;
};

// prefetch the per-plane constants of sample b into kp[WP_N][NC] (NC floats per kind and warp);
// `mine` = this lane copies the 16 bytes at channel offset `off` (floats) of the warp's span
__device__ __forceinline__ void ww_fetch_plane_consts(float* kp, int NC, const WwArgs& a, int b, int c_span0, int off,
                                                       bool mine) {
  if (mine && b < a.B) {
    const size_t g = (size_t)b * a.C + c_span0 + off;
    if (a.xf.se) ww_cp16(ww_s32(kp + WP_SE * NC + off), a.xf.se + g);
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int K, bool SEL>
__global__ void __launch_bounds__(WW_THREADS, K == 3 ? 4 : 3)
ww_conv_kernel(WwArgs a) {
  pdl_entry();
  constexpr int P = (K - 1) / 2;
  constexpr int D = WwDepth<T>::FWD;
  constexpr int NS = D + 1;                                        // ring slots: the one consumed last step is refilled
  constexpr int VB = WwRaw<T, 8>::BYTES;
  extern __shared__ __align__(16) uint8_t ww_dyn[];                // [warp][slot][lane] x VB bytes
  __shared__ __align__(16) float s_w[K * K * WW_WARPS * 32];       // [tap][channel of the CTA span]
  __shared__ __align__(16) float s_k[WW_WARPS][WK_N][32];
  __shared__ __align__(16) float s_p[WW_WARPS][2][WP_N][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int XL = 1 << a.xl_log2, VL = 32 >> a.xl_log2;
  const int x = lane & (XL - 1), v = lane >> a.xl_log2;
  const int span = WW_WARPS * VL * 8;
  const int c_cta = blockIdx.x * span;
  const int cl = (warp * VL + v) * 8;            // channel offset inside the CTA span
  const int c = c_cta + cl;
  const int H = a.H, W = a.W, C = a.C;
  for (int i = threadIdx.x; i < K * K * span; i += WW_THREADS) {
    const int tap = i / span, cc = c_cta + i % span;
    s_w[i] = cc < C ? a.w[(size_t)tap * C + cc] : 0.f;
  }
  const bool c_ok = c < C;
  const bool lane_ok = c_ok && x < W;
  float* kw = &s_k[warp][0][0];
  if (x == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      kw[WK_SC * 32 + v * 8 + i] = (c_ok && a.xf.scale) ? a.xf.scale[c + i] : 1.f;
      kw[WK_SH * 32 + v * 8 + i] = (c_ok && a.xf.scale) ? a.xf.shift[c + i] : 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) s_p[warp][u][WP_SE][v * 8 + i] = 1.f;   // default gate (kept when xf.se is null)
    }
  }
  __syncthreads();
  const float* wq = s_w + cl;
  const T* s0 = reinterpret_cast<const T*>(a.s0);
  T* out = reinterpret_cast<T*>(a.out);
  const int act = a.xf.act;
  const ActK ak = make_actk(act);
  const int steps = H + P;
  const int bstride = gridDim.y;
  const int c_warp = c_cta + warp * VL * 8;      // first channel of the warp's span
  const bool kc_mine = c_ok && x < 2;            // this lane copies 16 B (4 floats) of its vector's per-plane constants
  const int kc_off = v * 8 + x * 4;

  uint8_t* ring = ww_dyn + (size_t)warp * NS * 32 * VB + (size_t)lane * VB;
  const uint32_t ring_s = ww_s32(ring);
  constexpr uint32_t SLOT = 32 * VB, TSTRIDE = NS * SLOT;          // bytes per ring slot / per ring

  // ---- prefetch stream: 32-bit element offsets, advanced incrementally ----
  const uint32_t row_e = (uint32_t)W * (uint32_t)C;
  const uint32_t lane_e = (uint32_t)x * (uint32_t)C + (uint32_t)c;
  const uint32_t plane_jump = ((uint32_t)bstride * (uint32_t)H - (uint32_t)steps) * row_e;
  int pb = blockIdx.y, piy = 0;
  uint32_t foff = (uint32_t)pb * (uint32_t)H * row_e + lane_e;     // element offset of row (pb, piy) for this lane
  uint32_t fslot = 0;                                              // byte offset of the slot to fill
  auto fetch = [&]() {
    if (pb < a.B && lane_ok && piy < H) WwRaw<T, 8>::fetch(ring_s + fslot, s0 + foff);
    ww_commit();
    foff += row_e;
    if (++piy == steps) { piy = 0; pb += bstride; foff += plane_jump; }
    fslot += SLOT;
    if (fslot == TSTRIDE) fslot = 0;
  };
  // constants of the first plane, then the first D rows
  ww_fetch_plane_consts(&s_p[warp][0][0][0], 32, a, (int)blockIdx.y, c_warp, kc_off, kc_mine);
  ww_commit();
  ww_wait<0>();
  __syncwarp();
#pragma unroll 1
  for (int d = 0; d < D; ++d) fetch();
  uint32_t cslot = 0;
  int pbuf = 0;

  for (int b = blockIdx.y; b < a.B; b += bstride) {
    // ---- per-plane constants: this plane's are in s_p[pbuf] (prefetched a plane ago), start the next plane's ----
    if (steps <= D) ww_wait<0>();                // short planes: the constants' group may be younger than the ring depth
    __syncwarp();
    const float* kp = &s_p[warp][pbuf][0][0];
    ww_fetch_plane_consts(&s_p[warp][pbuf ^ 1][0][0], 32, a, b + bstride, c_warp, kc_off, kc_mine);
    // S[k] = partial sums of output row (iy - P + k) before row iy is processed; row iy adds filter row 2P - k
    float2 S[K - 1][4];
#pragma unroll
    for (int k = 0; k < K - 1; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) S[k][i] = make_float2(0.f, 0.f);
    float2 st1[4], st2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { st1[i] = make_float2(0.f, 0.f); st2[i] = st1[i]; }
    uint32_t ooff = (uint32_t)b * (uint32_t)H * row_e + lane_e;    // next output row of this lane

#pragma unroll 1
    for (int iy = 0; iy < steps; ++iy) {
      ww_wait<D - 1>();                                   // this step's rows have landed (own copies only)
      WwRaw<T, 8> c0;
      c0.lds(ring + cslot);
      cslot += SLOT;
      if (cslot == TSTRIDE) cslot = 0;
      fetch();                                            // refills the slot consumed one step ago
      float2 o[4];
      if (iy < H) {                                       // warp-uniform
        float2 vals[K][4];
        if (lane_ok) {
          c0.get(vals[P]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 sc = *reinterpret_cast<const float2*>(kw + WK_SC * 32 + v * 8 + 2 * i);
            const float2 sh = *reinterpret_cast<const float2*>(kw + WK_SH * 32 + v * 8 + 2 * i);
            const float2 e = *reinterpret_cast<const float2*>(kp + WP_SE * 32 + v * 8 + 2 * i);
            float2 u = __ffma2_rn(vals[P][i], sc, sh);
            u = __fmul2_rn(u, e);
            vals[P][i] = make_float2(actk_fwd(u.x, ak), actk_fwd(u.y, ak));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) vals[P][i] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int d = 1; d <= P; ++d)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            vals[P - d][i] = ww_shift<SEL>(vals[P][i], -d, x, W);
            vals[P + d][i] = ww_shift<SEL>(vals[P][i], d, x, W);
          }
        // slot k receives filter row ky = 2P - k; the shift of the slots is folded into the accumulation
#pragma unroll
        for (int k = 0; k < K; ++k) {
          float2 t[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) t[i] = k < K - 1 ? S[k < K - 1 ? k : 0][i] : make_float2(0.f, 0.f);
          const int ky = 2 * P - k;
#pragma unroll
          for (int kx = 0; kx < K; ++kx) {
            const float4 w0 = *reinterpret_cast<const float4*>(wq + (ky * K + kx) * span);
            const float4 w1 = *reinterpret_cast<const float4*>(wq + (ky * K + kx) * span + 4);
            t[0] = __ffma2_rn(vals[kx][0], make_float2(w0.x, w0.y), t[0]);
            t[1] = __ffma2_rn(vals[kx][1], make_float2(w0.z, w0.w), t[1]);
            t[2] = __ffma2_rn(vals[kx][2], make_float2(w1.x, w1.y), t[2]);
            t[3] = __ffma2_rn(vals[kx][3], make_float2(w1.z, w1.w), t[3]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (k == 0) o[i] = t[i];
            else S[k - 1][i] = t[i];
          }
        }
      } else {                                            // zero rows below the plane: only drain the slots
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[i] = S[0][i];
#pragma unroll
          for (int k = 0; k + 2 < K; ++k) S[k][i] = S[k + 1][i];
          S[K - 2][i] = make_float2(0.f, 0.f);
        }
      }
      // ---- output row iy - P is complete ----
      if (iy >= P) {
        if (lane_ok) {
          float2 r[4];
          ww_store8(out + ooff, o, r);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            st1[i] = __fadd2_rn(st1[i], r[i]);
            st2[i] = __ffma2_rn(r[i], r[i], st2[i]);
          }
        }
        ooff += row_e;
      }
    }
    pbuf ^= 1;
    // ---- per-(sample, channel) statistics: sum over the image row lanes, one atomic per channel ----
    if (a.stats) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        for (int o2 = 1; o2 < XL; o2 <<= 1) {
          st1[i].x += __shfl_xor_sync(0xffffffffu, st1[i].x, o2);
          st1[i].y += __shfl_xor_sync(0xffffffffu, st1[i].y, o2);
          st2[i].x += __shfl_xor_sync(0xffffffffu, st2[i].x, o2);
          st2[i].y += __shfl_xor_sync(0xffffffffu, st2[i].y, o2);
        }
      }
      if (x == 0 && c_ok) {
        float* sp = a.stats + (size_t)b * 2 * C + c;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          atomicAdd(sp + 2 * i, st1[i].x);
          atomicAdd(sp + 2 * i + 1, st1[i].y);
          atomicAdd(sp + C + 2 * i, st2[i].x);
          atomicAdd(sp + C + 2 * i + 1, st2[i].y);
        }
      }
    }
  }
  ww_wait<0>();
}

// ------------------------------------------------------------------------------------------------
int ww_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int ww_xl_log2(int W) { return W <= 8 ? 3 : (W <= 16 ? 4 : 5); }

// grid: x = channel spans, y = sample lanes (each CTA walks samples blockIdx.y, + gridDim.y, ...)
dim3 ww_grid(int B, int C, int span, int ctas_per_sm) {
  const int gx = ceil_div(C, span);
  int gy = ceil_div(ww_num_sms() * ctas_per_sm * 2, gx);    // ~2 resident waves worth of CTAs
  if (gy > B) gy = B;
  if (gy < 1) gy = 1;
  // even out the samples per CTA
  gy = ceil_div(B, ceil_div(B, gy));
  return dim3(gx, gy, 1);
}

template <typename T, int K>
int ww_conv_launch(const WwArgs& a, cudaStream_t st) {
  const int XL = 1 << a.xl_log2;
  const int span = WW_WARPS * (32 >> a.xl_log2) * 8;
  const dim3 grid = ww_grid(a.B, a.C, span, K == 3 ? 4 : 3);
  const bool sel = a.W + (K - 1) / 2 > XL;
  const size_t smem = (size_t)WW_WARPS * (WwDepth<T>::FWD + 1) * 32 * WwRaw<T, 8>::BYTES;
  static bool once = false;
  if (!once) {
    TD3D_CUDA(cudaFuncSetAttribute(ww_conv_kernel<T, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    TD3D_CUDA(cudaFuncSetAttribute(ww_conv_kernel<T, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    once = true;
  }
  if (sel) TD3D_CUDA(launch_kernel(ww_conv_kernel<T, K, true>, grid, WW_THREADS, smem, st, a));
  else TD3D_CUDA(launch_kernel(ww_conv_kernel<T, K, false>, grid, WW_THREADS, smem, st, a));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace

// the walker handles stride-1 layers whose image rows fit one warp
bool dw_walker_supported(int H, int W, int C, int k, int stride) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("TD3D_DW_WALKER");
    enabled = (e && atoi(e) == 0) ? 0 : 1;
  }
  return enabled && stride == 1 && (k == 3 || k == 5) && W <= 32 && H >= 1 && C % 8 == 0;
}

int launch_dw_fwd_walker(const DwArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE((double)a.B * a.H * a.W * a.C < 4294967296.0, "dw walker: tensor exceeds the 32-bit offset range");
  WwArgs w = {};
  w.s0 = a.x; w.xf = a.xf; w.w = a.w_taps; w.out = a.y; w.stats = a.stats;
  w.B = a.B; w.H = a.H; w.W = a.W; w.C = a.C; w.xl_log2 = ww_xl_log2(a.W);
  if (dtype == TD3D_BF16) return a.k == 3 ? ww_conv_launch<bf16, 3>(w, st) : ww_conv_launch<bf16, 5>(w, st);
  return a.k == 3 ? ww_conv_launch<float, 3>(w, st) : ww_conv_launch<float, 5>(w, st);
}

}  // namespace td3d
