// Depthwise k x k convolution (k in {3,5}, stride in {1,2}, pad (k-1)/2) on NHWC, shared-memory
// tiled (reference: nn.Conv2d(groups=hidden_dim) in InvertedResidual,
// torchdet3d/models/mobilenetv3.py:136,152).  The depthwise tensors are the widest of the network
// (35-45 % of all bytes), so these kernels are pure bandwidth work:
//
//   * a CTA owns one sample, an 8x8 tile of output pixels and a group of 32 channels; the input
//     tile (with halo) is loaded ONCE with 16-byte channel vectors, the producer's lazily applied
//     transform (BatchNorm fold + SE gate + activation, or the BatchNorm-backward affine) is
//     evaluated ONCE per element while staging it to shared memory as fp32, and the k*k taps then
//     run out of shared memory with 128-bit loads (conflict free: a quarter-warp reads one pixel's
//     32 consecutive channels);
//   * a thread owns 4 channels x 2 horizontally adjacent outputs (register reuse of the window);
//   * epilogues reduce BatchNorm / SE statistics with warp shuffles -> shared atomics -> one global
//     atomic per channel per CTA, slot = sample.
//
//   forward   y = dw(act(se*(scale*x+shift)))                      + sum y, sum y^2
//   bwd-data  gx = act'(u(x)) * dw^T(alpha*g + beta*y + gamma)     + sum gx, sum gx*x
//   bwd-wgt   dW[c,ky,kx] = sum gy * x_t(shifted)   persistent CTAs, taps accumulate in registers
#include "td3d_kernels.h"

#include <stdlib.h>

namespace td3d {

static const int DT = 8;            // output tile edge
static const int DCG = 32;          // channels per CTA
static const int DTHREADS = 256;    // 8 channel quads x 32 two-pixel strips

template <int K, int S> struct DwGeom {
  static constexpr int P = (K - 1) / 2;
  static constexpr int IT = (DT - 1) * S + K;        // staged input tile edge (with halo)
};

__device__ __forceinline__ void store4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(bf16* p, const float4& v) {
  uint2 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  h[0] = __floats2bfloat162_rn(v.x, v.y);
  h[1] = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = raw;
}
__device__ __forceinline__ float4 load4f(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load4f(const bf16* p) {
  uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ float4 round4(const float4& v) {
  return make_float4(to_f(from_f<T>(v.x)), to_f(from_f<T>(v.y)), to_f(from_f<T>(v.z)), to_f(from_f<T>(v.w)));
}
__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y);
  acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
}

// per-CTA channel constants staged in shared memory
struct DwConsts {
  float scale[DCG], shift[DCG], se[DCG];       // input transform
  float alpha[DCG], beta[DCG], gamma[DCG];     // BN-backward affine of the incoming gradient
};

__device__ __forceinline__ void load_consts(DwConsts& k, const XForm& xf, const float* alpha, const float* beta,
                                            const float* gamma, int b, int c0, int C) {
  for (int i = threadIdx.x; i < DCG; i += blockDim.x) {
    const int c = c0 + i;
    const bool on = c < C;
    k.scale[i] = (on && xf.scale) ? xf.scale[c] : 1.f;
    k.shift[i] = (on && xf.scale) ? xf.shift[c] : 0.f;
    k.se[i] = (on && xf.se) ? xf.se[(size_t)b * C + c] : 1.f;
    k.alpha[i] = (on && alpha) ? alpha[(size_t)b * C + c] : 0.f;
    k.beta[i] = (on && beta) ? beta[c] : 0.f;
    k.gamma[i] = (on && gamma) ? gamma[(size_t)b * C + c] : 0.f;
  }
}

// stage the transformed forward input x_t = act(se*(scale*x+shift)) for rows [r0, r0+NR) x cols [q0, q0+NC)
// of sample b into tile[NR][NC][DCG] (zero outside the image: the conv pads the ACTIVATED tensor)
template <typename T>
__device__ __forceinline__ void stage_input(float* tile, const T* __restrict__ x, const DwConsts& k, int act, int b,
                                            int H, int W, int C, int c0, int r0, int q0, int NR, int NC) {
  const int nvec = NR * NC * (DCG / 8);
  for (int idx = threadIdx.x; idx < nvec; idx += blockDim.x) {
    const int v8 = idx % (DCG / 8), pix = idx / (DCG / 8);
    const int r = pix / NC, q = pix % NC;
    const int gy = r0 + r, gx = q0 + q, c = c0 + v8 * 8;
    float v[8];
    if (gy >= 0 && gy < H && gx >= 0 && gx < W && c < C) {
      load8(x + (((size_t)b * H + gy) * W + gx) * C + c, v);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        v[i] = act_fwd(k.se[v8 * 8 + i] * fmaf(v[i], k.scale[v8 * 8 + i], k.shift[v8 * 8 + i]), act);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    float* d = tile + (size_t)pix * DCG + v8 * 8;
    store4(d, make_float4(v[0], v[1], v[2], v[3]));
    store4(d + 4, make_float4(v[4], v[5], v[6], v[7]));
  }
}

// stage gy = alpha*g + beta*y + gamma (zero outside the output image)
template <typename T>
__device__ __forceinline__ void stage_grad(float* tile, const T* __restrict__ g, const T* __restrict__ yo,
                                           const DwConsts& k, int b, int Ho, int Wo, int C, int c0, int r0, int q0,
                                           int NR, int NC) {
  const int nvec = NR * NC * (DCG / 8);
  for (int idx = threadIdx.x; idx < nvec; idx += blockDim.x) {
    const int v8 = idx % (DCG / 8), pix = idx / (DCG / 8);
    const int r = pix / NC, q = pix % NC;
    const int gy = r0 + r, gx = q0 + q, c = c0 + v8 * 8;
    float v[8];
    if (gy >= 0 && gy < Ho && gx >= 0 && gx < Wo && c < C) {
      const size_t off = (((size_t)b * Ho + gy) * Wo + gx) * C + c;
      float yv[8];
      load8(g + off, v);
      load8(yo + off, yv);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fmaf(k.alpha[v8 * 8 + i], v[i], fmaf(k.beta[v8 * 8 + i], yv[i], k.gamma[v8 * 8 + i]));
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
    float* d = tile + (size_t)pix * DCG + v8 * 8;
    store4(d, make_float4(v[0], v[1], v[2], v[3]));
    store4(d + 4, make_float4(v[4], v[5], v[6], v[7]));
  }
}

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// block-level reduction of per-thread (s1, s2) float4 partials over the 32 strip-threads of each
// channel quad, then one global atomic per channel: stats[b][0|1][c0 + ...]
__device__ __forceinline__ void reduce_stats(float4 s1, float4 s2, float* s_red /*[2][DCG]*/, float* stats, int b,
                                             int c0, int C, int quad) {
  float v[8] = {s1.x, s1.y, s1.z, s1.w, s2.x, s2.y, s2.z, s2.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 8);
    v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
  }
  if ((threadIdx.x & 31) < 8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(&s_red[quad * 4 + i], v[i]);
      atomicAdd(&s_red[DCG + quad * 4 + i], v[4 + i]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * DCG; i += blockDim.x) {
    const int which = i / DCG, cc = c0 + i % DCG;
    if (cc < C) atomicAdd(&stats[((size_t)b * 2 + which) * C + cc], s_red[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int K, int S>
__global__ void __launch_bounds__(DTHREADS)
dw_fwd_tiled_kernel(const T* __restrict__ x, XForm xf, const float* __restrict__ w, T* __restrict__ y,
                    float* __restrict__ stats, int H, int W, int Ho, int Wo, int C, int tiles_x) {
  using G = DwGeom<K, S>;
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                                   // [IT][IT][DCG]
  float* s_w = tile + G::IT * G::IT * DCG;              // [K*K][DCG]
  float* s_red = s_w + K * K * DCG;                     // [2][DCG]
  __shared__ DwConsts kc;
  const int b = blockIdx.z, c0 = blockIdx.y * DCG;
  const int oy0 = (blockIdx.x / tiles_x) * DT, ox0 = (blockIdx.x % tiles_x) * DT;
  load_consts(kc, xf, nullptr, nullptr, nullptr, b, c0, C);
  for (int i = threadIdx.x; i < K * K * DCG; i += blockDim.x) {
    const int c = c0 + i % DCG;
    s_w[i] = c < C ? w[(size_t)(i / DCG) * C + c] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * DCG; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  stage_input<T>(tile, x, kc, xf.act, b, H, W, C, c0, oy0 * S - G::P, ox0 * S - G::P, G::IT, G::IT);
  __syncthreads();
  const int quad = threadIdx.x & 7, strip = threadIdx.x >> 3;
  const int orow = strip >> 2, ocol = (strip & 3) * 2;
  float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
#pragma unroll
  for (int ky = 0; ky < K; ++ky) {
    const float* row = tile + ((size_t)(orow * S + ky) * G::IT + ocol * S) * DCG + quad * 4;
    if (S == 1) {
#pragma unroll
      for (int j = 0; j <= K; ++j) {
        const float4 v = lds4(row + j * DCG);
        if (j < K) fma4(acc0, v, lds4(s_w + (ky * K + j) * DCG + quad * 4));
        if (j >= 1) fma4(acc1, v, lds4(s_w + (ky * K + j - 1) * DCG + quad * 4));
      }
    } else {
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float4 wv = lds4(s_w + (ky * K + kx) * DCG + quad * 4);
        fma4(acc0, lds4(row + kx * DCG), wv);
        fma4(acc1, lds4(row + (S + kx) * DCG), wv);
      }
    }
  }
  const int oy = oy0 + orow, ox = ox0 + ocol, c = c0 + quad * 4;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (oy < Ho && c < C) {
    T* dst = y + (((size_t)b * Ho + oy) * Wo + ox) * C + c;
    if (ox < Wo) {
      store4(dst, acc0);
      const float4 r = round4<T>(acc0);
      s1 = r; s2 = make_float4(r.x * r.x, r.y * r.y, r.z * r.z, r.w * r.w);
    }
    if (ox + 1 < Wo) {
      store4(dst + C, acc1);
      const float4 r = round4<T>(acc1);
      s1.x += r.x; s1.y += r.y; s1.z += r.z; s1.w += r.w;
      s2.x = fmaf(r.x, r.x, s2.x); s2.y = fmaf(r.y, r.y, s2.y); s2.z = fmaf(r.z, r.z, s2.z); s2.w = fmaf(r.w, r.w, s2.w);
    }
  }
  if (stats) reduce_stats(s1, s2, s_red, stats, b, c0, C, quad);
}

// ------------------------------------------------------------------------------------------------
// backward data. Stride 1: a correlation with the flipped filter over the staged gy tile.
// Stride 2: a thread produces the 2x2 input pixels below one coarse (output-grid) position; every
// tap (ky,kx) feeds exactly one of the four parities with a compile-time gy offset, so there is no
// divergence and no wasted multiply.
// ------------------------------------------------------------------------------------------------
template <typename T, int K, int S>
__global__ void __launch_bounds__(DTHREADS)
dw_bwd_data_tiled_kernel(const T* __restrict__ g, const T* __restrict__ yo, const float* __restrict__ alpha,
                         const float* __restrict__ beta, const float* __restrict__ gamma, const T* __restrict__ x,
                         XForm xf, const float* __restrict__ w, T* __restrict__ gx, float* __restrict__ stats, int H,
                         int W, int Ho, int Wo, int C, int tiles_x) {
  constexpr int P = (K - 1) / 2;
  constexpr int HALO = S == 1 ? P : 1;                  // coarse halo on each side
  constexpr int GT = DT + 2 * HALO;                     // staged gy tile edge
  extern __shared__ __align__(16) float smem[];
  float* tile = smem;                                   // [GT][GT][DCG]
  float* s_w = tile + GT * GT * DCG;
  float* s_red = s_w + K * K * DCG;
  __shared__ DwConsts kc;
  const int b = blockIdx.z, c0 = blockIdx.y * DCG;
  const int ty0 = (blockIdx.x / tiles_x) * DT, tx0 = (blockIdx.x % tiles_x) * DT;   // coarse tile origin
  load_consts(kc, xf, alpha, beta, gamma, b, c0, C);
  for (int i = threadIdx.x; i < K * K * DCG; i += blockDim.x) {
    const int c = c0 + i % DCG;
    s_w[i] = c < C ? w[(size_t)(i / DCG) * C + c] : 0.f;
  }
  for (int i = threadIdx.x; i < 2 * DCG; i += blockDim.x) s_red[i] = 0.f;
  __syncthreads();
  stage_grad<T>(tile, g, yo, kc, b, Ho, Wo, C, c0, ty0 - HALO, tx0 - HALO, GT, GT);
  __syncthreads();
  const int quad = threadIdx.x & 7, strip = threadIdx.x >> 3;
  const int trow = strip >> 2, tcol = (strip & 3) * 2;
  const int c = c0 + quad * 4;
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  auto finish = [&](float4 acc, int h, int wv) {
    if (h >= H || wv >= W || c >= C) return;
    const size_t off = (((size_t)b * H + h) * W + wv) * C + c;
    const float4 xv = load4f(x + off);
    const float* sc = kc.scale + quad * 4;
    const float* sh = kc.shift + quad * 4;
    const float* se = kc.se + quad * 4;
    acc.x *= act_bwd(se[0] * fmaf(xv.x, sc[0], sh[0]), xf.act);
    acc.y *= act_bwd(se[1] * fmaf(xv.y, sc[1], sh[1]), xf.act);
    acc.z *= act_bwd(se[2] * fmaf(xv.z, sc[2], sh[2]), xf.act);
    acc.w *= act_bwd(se[3] * fmaf(xv.w, sc[3], sh[3]), xf.act);
    store4(gx + off, acc);
    const float4 r = round4<T>(acc);
    s1.x += r.x; s1.y += r.y; s1.z += r.z; s1.w += r.w;
    s2.x = fmaf(r.x, xv.x, s2.x); s2.y = fmaf(r.y, xv.y, s2.y); s2.z = fmaf(r.z, xv.z, s2.z); s2.w = fmaf(r.w, xv.w, s2.w);
  };
  if (S == 1) {
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
#pragma unroll
    for (int a = 0; a < K; ++a) {          // gx[h,w] = sum_{a,bb} gy[h-P+a, w-P+bb] * W[K-1-a][K-1-bb]
      const float* row = tile + ((size_t)(trow + a) * GT + tcol) * DCG + quad * 4;
#pragma unroll
      for (int j = 0; j <= K; ++j) {
        const float4 v = lds4(row + j * DCG);
        if (j < K) fma4(acc0, v, lds4(s_w + ((K - 1 - a) * K + (K - 1 - j)) * DCG + quad * 4));
        if (j >= 1) fma4(acc1, v, lds4(s_w + ((K - 1 - a) * K + (K - j)) * DCG + quad * 4));
      }
    }
    finish(acc0, ty0 + trow, tx0 + tcol);
    finish(acc1, ty0 + trow, tx0 + tcol + 1);
  } else {
#pragma unroll
    for (int t = 0; t < 2; ++t) {          // two coarse positions per thread
      float4 acc[2][2];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int a = (ky + P) & 1;                  // input-row parity fed by this tap
        const int dy = (a + P - ky) / 2;             // exact: a + P - ky is even; in [-1, 1]
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int bb = (kx + P) & 1;
          const int dx = (bb + P - kx) / 2;
          const float4 gv = lds4(tile + ((size_t)(trow + HALO + dy) * GT + (tcol + t + HALO + dx)) * DCG + quad * 4);
          fma4(acc[a][bb], gv, lds4(s_w + (ky * K + kx) * DCG + quad * 4));
        }
      }
      const int h0 = 2 * (ty0 + trow), w0 = 2 * (tx0 + tcol + t);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) finish(acc[i][j], h0 + i, w0 + j);
    }
  }
  if (stats) reduce_stats(s1, s2, s_red, stats, b, c0, C, quad);
}

// ------------------------------------------------------------------------------------------------
// backward weights: persistent CTAs (grid.x) over (sample, tile) work items of one channel group
// (grid.y); per-thread tap accumulators live in registers across all work items.
// ------------------------------------------------------------------------------------------------
template <typename T, int K, int S>
__global__ void __launch_bounds__(DTHREADS)
dw_bwd_weight_tiled_kernel(const T* __restrict__ g, const T* __restrict__ yo, const float* __restrict__ alpha,
                           const float* __restrict__ beta, const float* __restrict__ gamma, const T* __restrict__ x,
                           XForm xf, float* __restrict__ dw, int B, int H, int W, int Ho, int Wo, int C, int tiles_x,
                           int tiles_y) {
  using G = DwGeom<K, S>;
  extern __shared__ __align__(16) float smem[];
  float* xt = smem;                                     // [IT][IT][DCG] transformed forward input
  float* gt = xt + G::IT * G::IT * DCG;                 // [DT][DT][DCG]  gy
  float* s_dw = gt + DT * DT * DCG;                     // [K*K][DCG]
  __shared__ DwConsts kc;
  const int c0 = blockIdx.y * DCG;
  const int quad = threadIdx.x & 7, strip = threadIdx.x >> 3;
  const int orow = strip >> 2, ocol = (strip & 3) * 2;
  float4 acc[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = threadIdx.x; i < K * K * DCG; i += blockDim.x) s_dw[i] = 0.f;
  const int n_items = B * tiles_x * tiles_y;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int b = item / (tiles_x * tiles_y);
    const int oy0 = ((item / tiles_x) % tiles_y) * DT, ox0 = (item % tiles_x) * DT;
    __syncthreads();                                    // previous item's tiles fully consumed
    load_consts(kc, xf, alpha, beta, gamma, b, c0, C);
    __syncthreads();
    stage_input<T>(xt, x, kc, xf.act, b, H, W, C, c0, oy0 * S - G::P, ox0 * S - G::P, G::IT, G::IT);
    stage_grad<T>(gt, g, yo, kc, b, Ho, Wo, C, c0, oy0, ox0, DT, DT);
    __syncthreads();
    const float4 g0 = lds4(gt + ((size_t)orow * DT + ocol) * DCG + quad * 4);
    const float4 g1 = lds4(gt + ((size_t)orow * DT + ocol + 1) * DCG + quad * 4);
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const float* row = xt + ((size_t)(orow * S + ky) * G::IT + ocol * S) * DCG + quad * 4;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        fma4(acc[ky * K + kx], g0, lds4(row + kx * DCG));
        fma4(acc[ky * K + kx], g1, lds4(row + (S + kx) * DCG));
      }
    }
  }
  // reduce the 32 strip-threads of each quad, then flush (reference layout [C,1,K,K])
#pragma unroll
  for (int i = 0; i < K * K; ++i) {
    float v[4] = {acc[i].x, acc[i].y, acc[i].z, acc[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] += __shfl_xor_sync(0xffffffffu, v[j], 8);
      v[j] += __shfl_xor_sync(0xffffffffu, v[j], 16);
    }
    if ((threadIdx.x & 31) < 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(&s_dw[i * DCG + quad * 4 + j], v[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * DCG; i += blockDim.x) {
    const int tap = i / DCG, c = c0 + i % DCG;
    if (c < C) atomicAdd(&dw[(size_t)c * K * K + tap], s_dw[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// TD3D_DW_IMPL selects the depthwise generation for A/B debugging: 2 = k_dw2.cu (default),
// 1 = the 8x8-tile kernels of this file, 0 = the direct kernels of k_dwconv_simple.cu.
static int dw_impl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TD3D_DW_IMPL");
    v = e ? atoi(e) : 2;
    const char* s = getenv("TD3D_DW_SIMPLE");
    if (s && atoi(s)) v = 0;
  }
  return v;
}

template <typename KernelT>
static int ensure_smem(KernelT kernel, size_t bytes) {
  if (bytes > 48 * 1024) TD3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return TD3D_OK;
}

template <typename T, int K, int S>
static int dw_fwd_t(const DwArgs& a, cudaStream_t st) {
  using G = DwGeom<K, S>;
  const int Ho = (a.H - 1) / S + 1, Wo = (a.W - 1) / S + 1;
  const int tx = ceil_div(Wo, DT), ty = ceil_div(Ho, DT);
  const size_t smem = sizeof(float) * (G::IT * G::IT * DCG + K * K * DCG + 2 * DCG);
  TD3D_TRY(ensure_smem(dw_fwd_tiled_kernel<T, K, S>, smem));
  dim3 grid(tx * ty, ceil_div(a.C, DCG), a.B);
  dw_fwd_tiled_kernel<T, K, S><<<grid, DTHREADS, smem, st>>>((const T*)a.x, a.xf, a.w_taps, (T*)a.y, a.stats, a.H, a.W, Ho,
                                                             Wo, a.C, tx);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

template <typename T, int K, int S>
static int dw_bwd_t(const DwBwdArgs& a, cudaStream_t st) {
  using G = DwGeom<K, S>;
  const int Ho = (a.H - 1) / S + 1, Wo = (a.W - 1) / S + 1;
  if (a.gx) {
    constexpr int HALO = S == 1 ? (K - 1) / 2 : 1;
    constexpr int GT = DT + 2 * HALO;
    // tiles over the input grid (stride 1) or the coarse/output grid (stride 2: 2x2 input pixels each)
    const int ext_y = S == 1 ? a.H : (a.H + 1) / 2, ext_x = S == 1 ? a.W : (a.W + 1) / 2;
    const int tx = ceil_div(ext_x, DT), ty = ceil_div(ext_y, DT);
    const size_t smem = sizeof(float) * (GT * GT * DCG + K * K * DCG + 2 * DCG);
    TD3D_TRY(ensure_smem(dw_bwd_data_tiled_kernel<T, K, S>, smem));
    dim3 grid(tx * ty, ceil_div(a.C, DCG), a.B);
    dw_bwd_data_tiled_kernel<T, K, S><<<grid, DTHREADS, smem, st>>>(
        (const T*)a.g, (const T*)a.y_out, a.alpha, a.beta, a.gamma, (const T*)a.x, a.xf, a.w_taps, (T*)a.gx, a.stats, a.H,
        a.W, Ho, Wo, a.C, tx);
    TD3D_LAUNCH_CHECK();
  }
  if (a.dw) {
    const int tx = ceil_div(Wo, DT), ty = ceil_div(Ho, DT);
    const size_t smem = sizeof(float) * (G::IT * G::IT * DCG + DT * DT * DCG + K * K * DCG);
    TD3D_TRY(ensure_smem(dw_bwd_weight_tiled_kernel<T, K, S>, smem));
    const int groups = ceil_div(a.C, DCG);
    int items = a.B * tx * ty;
    int per = ceil_div(148 * 6, groups);
    if (per > items) per = items;
    dim3 grid(per, groups);
    dw_bwd_weight_tiled_kernel<T, K, S><<<grid, DTHREADS, smem, st>>>(
        (const T*)a.g, (const T*)a.y_out, a.alpha, a.beta, a.gamma, (const T*)a.x, a.xf, a.dw, a.B, a.H, a.W, Ho, Wo, a.C,
        tx, ty);
    TD3D_LAUNCH_CHECK();
  }
  return TD3D_OK;
}

#define DW_DISPATCH(FN, ARGS)                                                                   \
  do {                                                                                          \
    if (dtype == TD3D_BF16) {                                                                   \
      if (a.k == 3 && a.stride == 1) return FN<bf16, 3, 1>(ARGS, st);                           \
      if (a.k == 3 && a.stride == 2) return FN<bf16, 3, 2>(ARGS, st);                           \
      if (a.k == 5 && a.stride == 1) return FN<bf16, 5, 1>(ARGS, st);                           \
      if (a.k == 5 && a.stride == 2) return FN<bf16, 5, 2>(ARGS, st);                           \
    } else {                                                                                    \
      if (a.k == 3 && a.stride == 1) return FN<float, 3, 1>(ARGS, st);                          \
      if (a.k == 3 && a.stride == 2) return FN<float, 3, 2>(ARGS, st);                          \
      if (a.k == 5 && a.stride == 1) return FN<float, 5, 1>(ARGS, st);                          \
      if (a.k == 5 && a.stride == 2) return FN<float, 5, 2>(ARGS, st);                          \
    }                                                                                           \
  } while (0)

int launch_dw_fwd(const DwArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B <= 65535, "dw fwd: C=%d must be a multiple of 8", a.C);
  if (dw_impl() == 0) return launch_dw_fwd_simple(a, dtype, st);
  if (dw_impl() == 2 && dw_walker_supported(a.H, a.W, a.C, a.k, a.stride)) return launch_dw_fwd_walker(a, dtype, st);
  if (dw_impl() == 2) return launch_dw_fwd_v2(a, dtype, st);
  DW_DISPATCH(dw_fwd_t, a);
  set_last_error("dw fwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  return TD3D_EINVAL;
}

int launch_dw_bwd(const DwBwdArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B <= 65535, "dw bwd: C=%d must be a multiple of 8", a.C);
  if (dw_impl() == 0) return launch_dw_bwd_simple(a, dtype, st);
  if (dw_impl() == 2 && dw_walker_supported(a.H, a.W, a.C, a.k, a.stride)) return launch_dw_bwd_walker(a, dtype, st);
  if (dw_impl() == 2) return launch_dw_bwd_v2(a, dtype, st);
  DW_DISPATCH(dw_bwd_t, a);
  set_last_error("dw bwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  return TD3D_EINVAL;
}

}  // namespace td3d
