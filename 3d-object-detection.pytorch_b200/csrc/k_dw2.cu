// Depthwise k x k convolution, generation 2 (k in {3,5}, stride in {1,2}, pad (k-1)/2, NHWC).
// Reference op: nn.Conv2d(hidden, hidden, k, s, (k-1)//2, groups=hidden) inside InvertedResidual,
// torchdet3d/models/mobilenetv3.py:136,152, and its autograd backward.
//
// Why a second generation: the depthwise tensors are the widest of the network, and at B200's HBM
// rate (~23 B/clk/SM) a 3x3 depthwise layer has a budget of only ~22 issued instructions per element,
// a 5x5 layer is FMA-bound outright.  So these kernels are built to be instruction-lean:
//   * host-chosen CTA tiles (pick_tile) as large as shared memory allows, several samples per CTA
//     for 7x7 / 14x14 maps, so halo re-staging and per-CTA fixed costs are amortised;
//   * the producer's lazily applied transform (BatchNorm fold + SE gate + activation, or the
//     BatchNorm-backward affine of the gradient) is evaluated ONCE per staged element, fp32 in smem;
//   * a thread owns 4 channels x (2x4 | 1x4) outputs and walks the window row by row with 128-bit
//     conflict-free shared loads; multiply-adds are packed fp32x2 (FFMA2, sm_100);
//   * statistics (BatchNorm sums / SE squeeze) leave the CTA as one global atomic per channel.
//
//   forward   y  = dw(act(se*(scale*x+shift)))                      + sum y,  sum y^2   per (b,c)
//   bwd-data  gx = act'(u(x)) * dw^T(alpha*g + beta*y + gamma)      + sum gx, sum gx*x  per (b,c)
//   bwd-wgt   dW[c,ky,kx] += sum gy * x_t(shifted)   persistent CTAs, taps accumulate in registers
#include "td3d_kernels.h"

#include <stdlib.h>

#include <type_traits>

namespace td3d {

namespace {

constexpr int D2_THREADS = 256;
constexpr int D2_MAXCG = 32;        // channels per CTA (16 or 32)
constexpr int D2_MAXNB = 8;         // samples per CTA
constexpr int D2_U = 4;             // staging loads in flight per thread

struct D2Tile {
  int cg, ps, nv8;            // channels per CTA, smem pixel stride (floats), 8-channel vectors per pixel
  int tyt, txt, nb;           // thread-tiles per CTA (y, x), samples per CTA
  int th, tw;                 // owned tile extent (pixels of the owned grid)
  int ih, iw;                 // staged (halo) tile extent
  int tiles_y, tiles_x, b_blocks, n_groups;
};

struct D2Consts {
  float sc[D2_MAXCG], sh[D2_MAXCG], be[D2_MAXCG];
  float se[D2_MAXNB][D2_MAXCG], al[D2_MAXNB][D2_MAXCG], ga[D2_MAXNB][D2_MAXCG];
};

// ---- small vector helpers ---------------------------------------------------------------------
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 lds2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void sts4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
  const float2 lo = __ffma2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y), make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void fmav(float4& acc, const float4& a, const float4& b) { fma4(acc, a, b); }
__device__ __forceinline__ void fmav(float2& acc, const float2& a, const float2& b) { acc = __ffma2_rn(a, b, acc); }
template <int V> struct VecOf { typedef float4 type; };
template <> struct VecOf<2> { typedef float2 type; };
__device__ __forceinline__ void ldsv(float4& v, const float* p) { v = lds4(p); }
__device__ __forceinline__ void ldsv(float2& v, const float* p) { v = lds2(p); }
__device__ __forceinline__ void zerov(float4& v) { v = make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void zerov(float2& v) { v = make_float2(0.f, 0.f); }
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
__device__ __forceinline__ float comp(const float2& v, int i) { return i == 0 ? v.x : v.y; }

__device__ __forceinline__ void st4(float* p, const float4& v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, const float4& v) {
  uint2 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  h[0] = __floats2bfloat162_rn(v.x, v.y);
  h[1] = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = raw;
}
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                     __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
}
template <typename T> __device__ __forceinline__ float4 rnd4(const float4& v) {
  return make_float4(to_f(from_f<T>(v.x)), to_f(from_f<T>(v.y)), to_f(from_f<T>(v.z)), to_f(from_f<T>(v.w)));
}

// raw 8-channel vector as loaded from global memory
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  uint4 r;
  __device__ __forceinline__ void load(const bf16* p) { r = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void unpack(float v[8]) const {
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
    v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
  }
};
template <> struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldg(reinterpret_cast<const float4*>(p));
    b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void unpack(float v[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

__device__ __forceinline__ float act_f(float u, int act) {
  if (act == TD3D_ACT_RELU) return fmaxf(u, 0.f);
  if (act == TD3D_ACT_HSWISH) return u * (fminf(fmaxf(u + 3.f, 0.f), 6.f) * (1.f / 6.f));
  return u;
}

// ---- per-CTA constants ------------------------------------------------------------------------
__device__ __forceinline__ void d2_load_consts(D2Consts& k, const XForm& xf, const float* __restrict__ alpha,
                                               const float* __restrict__ beta, const float* __restrict__ gamma,
                                               int b0, int nb, int B, int c0, int C) {
  for (int i = threadIdx.x; i < D2_MAXCG; i += D2_THREADS) {
    const int c = c0 + i;
    const bool on = c < C;
    k.sc[i] = (on && xf.scale) ? xf.scale[c] : 1.f;
    k.sh[i] = (on && xf.scale) ? xf.shift[c] : 0.f;
    k.be[i] = (on && beta) ? beta[c] : 0.f;
  }
  for (int i = threadIdx.x; i < nb * D2_MAXCG; i += D2_THREADS) {
    const int n = i / D2_MAXCG, cc = i % D2_MAXCG;
    const int c = c0 + cc, b = b0 + n;
    const bool on = c < C && b < B;
    k.se[n][cc] = (on && xf.se) ? xf.se[(size_t)b * C + c] : 1.f;
    k.al[n][cc] = (on && alpha) ? alpha[(size_t)b * C + c] : 0.f;
    k.ga[n][cc] = (on && gamma) ? gamma[(size_t)b * C + c] : 0.f;
  }
}

// ---- staging ----------------------------------------------------------------------------------
// Walks the (sample, row, col) items of a [nb][rows][cols] tile for one fixed 8-channel vector per
// thread, D2_U items at a time (loads first, then transform + store), without divisions in the loop.
struct D2Walk {
  int nbi, r, c, dr, dc, rows, cols, nb;
  __device__ __forceinline__ void init(int rows_, int cols_, int nb_, int nv8) {
    rows = rows_; cols = cols_; nb = nb_;
    const int dpix = D2_THREADS / nv8;
    const int pix = threadIdx.x / nv8;
    c = pix % cols;
    const int rr = pix / cols;
    r = rr % rows; nbi = rr / rows;
    dc = dpix % cols; dr = dpix / cols;
  }
  __device__ __forceinline__ bool live() const { return nbi < nb; }
  __device__ __forceinline__ void next() {
    c += dc; r += dr;
    if (c >= cols) { c -= cols; ++r; }
    while (r >= rows) { r -= rows; ++nbi; }
  }
};

// tile[nbi][r][c][ch] = act(se*(scale*x+shift)) for sample b0+nbi, pixel (y0+r, x0+c); zero outside
template <typename T>
__device__ __forceinline__ void d2_stage_x(float* __restrict__ tile, const T* __restrict__ x, const D2Consts& k, int act,
                                           bool has_se, int ps, int nv8, int b0, int B, int H, int W, int C, int c0, int y0,
                                           int x0, int rows, int cols, int nb) {
  const int v8 = threadIdx.x % nv8;
  const int ch = c0 + v8 * 8;
  const bool ch_ok = ch < C;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = k.sc[v8 * 8 + i]; sh[i] = k.sh[v8 * 8 + i]; }
  D2Walk w;
  w.init(rows, cols, nb, nv8);
  while (w.live()) {
    Raw8<T> raw[D2_U];
    float* dst[D2_U];
    int sn[D2_U];
    bool ok[D2_U], lv[D2_U];
#pragma unroll
    for (int u = 0; u < D2_U; ++u) {
      lv[u] = w.live();
      const int gy = y0 + w.r, gx = x0 + w.c, b = b0 + w.nbi;
      ok[u] = lv[u] && ch_ok && gy >= 0 && gy < H && gx >= 0 && gx < W && b < B;
      dst[u] = tile + ((size_t)(w.nbi * rows + w.r) * cols + w.c) * ps + v8 * 8;
      sn[u] = w.nbi;
      if (ok[u]) raw[u].load(x + (((size_t)b * H + gy) * W + gx) * C + ch);
      if (lv[u]) w.next();
    }
#pragma unroll
    for (int u = 0; u < D2_U; ++u) {
      if (!lv[u]) continue;
      float v[8];
      if (ok[u]) {
        raw[u].unpack(v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
        if (has_se) {
          const float4 e0 = lds4(&k.se[sn[u]][v8 * 8]), e1 = lds4(&k.se[sn[u]][v8 * 8 + 4]);
          v[0] *= e0.x; v[1] *= e0.y; v[2] *= e0.z; v[3] *= e0.w;
          v[4] *= e1.x; v[5] *= e1.y; v[6] *= e1.z; v[7] *= e1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act_f(v[i], act);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      sts4(dst[u], v[0], v[1], v[2], v[3]);
      sts4(dst[u] + 4, v[4], v[5], v[6], v[7]);
    }
  }
}

// tile[nbi][r][c][ch] = alpha*g + beta*y + gamma for output pixel (y0+r, x0+c); zero outside
template <typename T>
__device__ __forceinline__ void d2_stage_gy(float* __restrict__ tile, const T* __restrict__ g, const T* __restrict__ yo,
                                            const D2Consts& k, int ps, int nv8, int b0, int B, int Ho, int Wo, int C, int c0,
                                            int y0, int x0, int rows, int cols, int nb) {
  const int v8 = threadIdx.x % nv8;
  const int ch = c0 + v8 * 8;
  const bool ch_ok = ch < C;
  float be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) be[i] = k.be[v8 * 8 + i];
  D2Walk w;
  w.init(rows, cols, nb, nv8);
  constexpr int U = D2_U / 2;
  while (w.live()) {
    Raw8<T> rg[U], ry[U];
    float* dst[U];
    int sn[U];
    bool ok[U], lv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      lv[u] = w.live();
      const int gy = y0 + w.r, gx = x0 + w.c, b = b0 + w.nbi;
      ok[u] = lv[u] && ch_ok && gy >= 0 && gy < Ho && gx >= 0 && gx < Wo && b < B;
      dst[u] = tile + ((size_t)(w.nbi * rows + w.r) * cols + w.c) * ps + v8 * 8;
      sn[u] = w.nbi;
      if (ok[u]) {
        const size_t off = (((size_t)b * Ho + gy) * Wo + gx) * C + ch;
        rg[u].load(g + off);
        ry[u].load(yo + off);
      }
      if (lv[u]) w.next();
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!lv[u]) continue;
      float v[8];
      if (ok[u]) {
        float yv[8];
        rg[u].unpack(v);
        ry[u].unpack(yv);
        const float4 a0 = lds4(&k.al[sn[u]][v8 * 8]), a1 = lds4(&k.al[sn[u]][v8 * 8 + 4]);
        const float4 g0 = lds4(&k.ga[sn[u]][v8 * 8]), g1 = lds4(&k.ga[sn[u]][v8 * 8 + 4]);
        v[0] = fmaf(a0.x, v[0], fmaf(be[0], yv[0], g0.x)); v[1] = fmaf(a0.y, v[1], fmaf(be[1], yv[1], g0.y));
        v[2] = fmaf(a0.z, v[2], fmaf(be[2], yv[2], g0.z)); v[3] = fmaf(a0.w, v[3], fmaf(be[3], yv[3], g0.w));
        v[4] = fmaf(a1.x, v[4], fmaf(be[4], yv[4], g1.x)); v[5] = fmaf(a1.y, v[5], fmaf(be[5], yv[5], g1.y));
        v[6] = fmaf(a1.z, v[6], fmaf(be[6], yv[6], g1.z)); v[7] = fmaf(a1.w, v[7], fmaf(be[7], yv[7], g1.w));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      sts4(dst[u], v[0], v[1], v[2], v[3]);
      sts4(dst[u] + 4, v[4], v[5], v[6], v[7]);
    }
  }
}

// weights [K*K][C] fp32 -> s_w[K*K][D2_MAXCG] (optionally flipped: tap (ky,kx) <- (K-1-ky, K-1-kx))
template <int K>
__device__ __forceinline__ void d2_load_w(float* s_w, const float* __restrict__ w, int c0, int C, bool flip) {
  for (int i = threadIdx.x; i < K * K * D2_MAXCG; i += D2_THREADS) {
    const int tap = i / D2_MAXCG, c = c0 + i % D2_MAXCG;
    const int src = flip ? (K * K - 1 - tap) : tap;
    s_w[i] = c < C ? w[(size_t)src * C + c] : 0.f;
  }
}

// out[oy][ox] = sum_{ky,kx} tile[(oy*S+ky)][(ox*S+kx)] * w[ky][kx] for a thread's OY x OX outputs
template <int K, int S, int OY, int OX>
__device__ __forceinline__ void d2_conv(float4 (&acc)[OY][OX], const float* base, int row_stride, int ps,
                                        const float* s_wq) {
  constexpr int NR = (OY - 1) * S + K, NC = (OX - 1) * S + K;
  float4 wreg[K == 3 ? 9 : 1];
  if (K == 3) {
#pragma unroll
    for (int i = 0; i < 9; ++i) wreg[i] = lds4(s_wq + i * D2_MAXCG);
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    float4 row[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) row[j] = lds4(base + r * row_stride + j * ps);
#pragma unroll
    for (int oy = 0; oy < OY; ++oy) {
      const int ky = r - oy * S;
      if (ky < 0 || ky >= K) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float4 wv = K == 3 ? wreg[ky * K + kx] : lds4(s_wq + (ky * K + kx) * D2_MAXCG);
#pragma unroll
        for (int ox = 0; ox < OX; ++ox) fma4(acc[oy][ox], row[ox * S + kx], wv);
      }
    }
  }
}

// per-(sample, channel) sums: thread partials -> smem [sp][2][cg] -> column sums -> one global atomic
__device__ __forceinline__ void d2_reduce_stats(float* part, const float4& s1, const float4& s2, int sp, int q, int cg,
                                                int per_sample, int nb, float* __restrict__ stats, int b0, int B, int c0,
                                                int C) {
  __syncthreads();                                   // everybody is done reading the tile
  float* mine = part + (size_t)sp * 2 * cg + q * 4;
  st4(mine, s1);
  st4(mine + cg, s2);
  __syncthreads();
  for (int t = threadIdx.x; t < nb * 2 * cg; t += D2_THREADS) {
    const int n = t / (2 * cg), j = t % (2 * cg);
    float s = 0.f;
    for (int i = 0; i < per_sample; ++i) s += part[(size_t)(n * per_sample + i) * 2 * cg + j];
    const int which = j / cg, c = c0 + j % cg, b = b0 + n;
    if (c < C && b < B) atomicAdd(&stats[((size_t)b * 2 + which) * C + c], s);
  }
}

struct D2Map {       // thread -> (channel quad, sample, thread-tile)
  int q, sp, nbi, ty, tx;
  bool active;
  __device__ __forceinline__ void init(const D2Tile& t, int nqv) {
    q = threadIdx.x % nqv;
    sp = threadIdx.x / nqv;
    tx = sp % t.txt;
    ty = (sp / t.txt) % t.tyt;
    nbi = sp / (t.txt * t.tyt);
    active = nbi < t.nb;
  }
};

__device__ __forceinline__ void d2_block(const D2Tile& t, int& b0, int& ty0, int& tx0) {
  const int bx = blockIdx.x;
  tx0 = (bx % t.tiles_x) * t.tw;
  ty0 = ((bx / t.tiles_x) % t.tiles_y) * t.th;
  b0 = (bx / (t.tiles_x * t.tiles_y)) * t.nb;
}

template <int S> struct D2Geo { static constexpr int OY = S == 1 ? 2 : 1, OX = 4; };

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int K, int S>
__global__ void __launch_bounds__(D2_THREADS, K == 3 ? 3 : 2)
d2_fwd_kernel(const T* __restrict__ x, XForm xf, const float* __restrict__ w, T* __restrict__ y,
              float* __restrict__ stats, int B, int H, int W, int Ho, int Wo, int C, D2Tile t) {
  constexpr int P = (K - 1) / 2, OY = D2Geo<S>::OY, OX = D2Geo<S>::OX;
  extern __shared__ __align__(16) float d2_smem[];
  __shared__ D2Consts kc;
  __shared__ __align__(16) float s_w[K * K * D2_MAXCG];
  float* tile = d2_smem;
  int b0, ty0, tx0;
  d2_block(t, b0, ty0, tx0);
  const int c0 = blockIdx.y * t.cg;
  d2_load_consts(kc, xf, nullptr, nullptr, nullptr, b0, t.nb, B, c0, C);
  d2_load_w<K>(s_w, w, c0, C, false);
  __syncthreads();
  d2_stage_x<T>(tile, x, kc, xf.act, xf.se != nullptr, t.ps, t.nv8, b0, B, H, W, C, c0, ty0 * S - P, tx0 * S - P, t.ih,
                t.iw, t.nb);
  __syncthreads();
  D2Map m;
  m.init(t, t.cg / 4);
  float4 acc[OY][OX];
#pragma unroll
  for (int i = 0; i < OY; ++i)
#pragma unroll
    for (int j = 0; j < OX; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  if (m.active) {
    const float* base = tile + ((size_t)(m.nbi * t.ih + m.ty * OY * S) * t.iw + m.tx * OX * S) * t.ps + m.q * 4;
    d2_conv<K, S, OY, OX>(acc, base, t.iw * t.ps, t.ps, s_w + m.q * 4);
    const int b = b0 + m.nbi, c = c0 + m.q * 4;
    if (b < B && c < C) {
#pragma unroll
      for (int oy = 0; oy < OY; ++oy) {
        const int yy = ty0 + m.ty * OY + oy;
        if (yy >= Ho) continue;
#pragma unroll
        for (int ox = 0; ox < OX; ++ox) {
          const int xx = tx0 + m.tx * OX + ox;
          if (xx >= Wo) continue;
          st4(y + (((size_t)b * Ho + yy) * Wo + xx) * C + c, acc[oy][ox]);
          const float4 r = rnd4<T>(acc[oy][ox]);
          s1.x += r.x; s1.y += r.y; s1.z += r.z; s1.w += r.w;
          fma4(s2, r, r);
        }
      }
    }
  }
  if (stats) d2_reduce_stats(tile, s1, s2, m.sp, m.q, t.cg, t.tyt * t.txt, t.nb, stats, b0, B, c0, C);
}

// ------------------------------------------------------------------------------------------------
// backward data, stride 1: correlation of the staged gy tile with the flipped filter
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_d(float u, int act) {
  if (act == TD3D_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (act == TD3D_ACT_HSWISH) return u <= -3.f ? 0.f : (u >= 3.f ? 1.f : fmaf(u, 1.f / 3.f, 0.5f));
  return 1.f;
}

struct D2Fin {       // epilogue of the data gradient: gx = acc * act'(u(x)), statistics of the stored value
  float4 sc, sh, se, s1, s2;
  int act;
  template <typename T>
  __device__ __forceinline__ void apply(float4 acc, const T* __restrict__ x, T* __restrict__ gx, size_t off) {
    const float4 xv = ld4(x + off);
    acc.x *= act_d(se.x * fmaf(xv.x, sc.x, sh.x), act);
    acc.y *= act_d(se.y * fmaf(xv.y, sc.y, sh.y), act);
    acc.z *= act_d(se.z * fmaf(xv.z, sc.z, sh.z), act);
    acc.w *= act_d(se.w * fmaf(xv.w, sc.w, sh.w), act);
    st4(gx + off, acc);
    const float4 r = rnd4<T>(acc);
    s1.x += r.x; s1.y += r.y; s1.z += r.z; s1.w += r.w;
    fma4(s2, r, xv);
  }
};

template <typename T, int K>
__global__ void __launch_bounds__(D2_THREADS, K == 3 ? 3 : 2)
d2_bwd_data_s1_kernel(const T* __restrict__ g, const T* __restrict__ yo, const float* __restrict__ alpha,
                      const float* __restrict__ beta, const float* __restrict__ gamma, const T* __restrict__ x, XForm xf,
                      const float* __restrict__ w, T* __restrict__ gx, float* __restrict__ stats, int B, int H, int W, int C,
                      D2Tile t) {
  constexpr int P = (K - 1) / 2, OY = 2, OX = 4;
  extern __shared__ __align__(16) float d2_smem[];
  __shared__ D2Consts kc;
  __shared__ __align__(16) float s_w[K * K * D2_MAXCG];
  float* tile = d2_smem;
  int b0, ty0, tx0;
  d2_block(t, b0, ty0, tx0);
  const int c0 = blockIdx.y * t.cg;
  d2_load_consts(kc, xf, alpha, beta, gamma, b0, t.nb, B, c0, C);
  d2_load_w<K>(s_w, w, c0, C, true);
  __syncthreads();
  d2_stage_gy<T>(tile, g, yo, kc, t.ps, t.nv8, b0, B, H, W, C, c0, ty0 - P, tx0 - P, t.ih, t.iw, t.nb);
  __syncthreads();
  D2Map m;
  m.init(t, t.cg / 4);
  D2Fin fin;
  fin.s1 = make_float4(0.f, 0.f, 0.f, 0.f); fin.s2 = fin.s1; fin.act = xf.act;
  if (m.active) {
    float4 acc[OY][OX];
#pragma unroll
    for (int i = 0; i < OY; ++i)
#pragma unroll
      for (int j = 0; j < OX; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* base = tile + ((size_t)(m.nbi * t.ih + m.ty * OY) * t.iw + m.tx * OX) * t.ps + m.q * 4;
    d2_conv<K, 1, OY, OX>(acc, base, t.iw * t.ps, t.ps, s_w + m.q * 4);
    const int b = b0 + m.nbi, c = c0 + m.q * 4;
    if (b < B && c < C) {
      fin.sc = lds4(&kc.sc[m.q * 4]); fin.sh = lds4(&kc.sh[m.q * 4]); fin.se = lds4(&kc.se[m.nbi][m.q * 4]);
#pragma unroll
      for (int oy = 0; oy < OY; ++oy) {
        const int yy = ty0 + m.ty * OY + oy;
        if (yy >= H) continue;
#pragma unroll
        for (int ox = 0; ox < OX; ++ox) {
          const int xx = tx0 + m.tx * OX + ox;
          if (xx >= W) continue;
          fin.apply<T>(acc[oy][ox], x, gx, (((size_t)b * H + yy) * W + xx) * C + c);
        }
      }
    }
  }
  if (stats) d2_reduce_stats(tile, fin.s1, fin.s2, m.sp, m.q, t.cg, t.tyt * t.txt, t.nb, stats, b0, B, c0, C);
}

// ------------------------------------------------------------------------------------------------
// backward data, stride 2.  The owned grid is the coarse (= output) grid; a thread produces the
// 2x2 input pixels under each of its two coarse positions.  Tap (ky,kx) feeds exactly one of the
// four input parities with a compile-time gy offset: no divergence, no wasted multiply.
// ------------------------------------------------------------------------------------------------
template <typename T, int K>
__global__ void __launch_bounds__(D2_THREADS, 2)
d2_bwd_data_s2_kernel(const T* __restrict__ g, const T* __restrict__ yo, const float* __restrict__ alpha,
                      const float* __restrict__ beta, const float* __restrict__ gamma, const T* __restrict__ x, XForm xf,
                      const float* __restrict__ w, T* __restrict__ gx, float* __restrict__ stats, int B, int H, int W, int Ho,
                      int Wo, int C, D2Tile t) {
  constexpr int P = (K - 1) / 2, OXC = 2, OYC = 2;
  extern __shared__ __align__(16) float d2_smem[];
  __shared__ D2Consts kc;
  __shared__ __align__(16) float s_w[K * K * D2_MAXCG];
  float* tile = d2_smem;
  int b0, ty0, tx0;
  d2_block(t, b0, ty0, tx0);
  const int c0 = blockIdx.y * t.cg;
  d2_load_consts(kc, xf, alpha, beta, gamma, b0, t.nb, B, c0, C);
  d2_load_w<K>(s_w, w, c0, C, false);
  __syncthreads();
  d2_stage_gy<T>(tile, g, yo, kc, t.ps, t.nv8, b0, B, Ho, Wo, C, c0, ty0 - 1, tx0 - 1, t.ih, t.iw, t.nb);
  __syncthreads();
  D2Map m;
  m.init(t, t.cg / 4);
  D2Fin fin;
  fin.s1 = make_float4(0.f, 0.f, 0.f, 0.f); fin.s2 = fin.s1; fin.act = xf.act;
  if (m.active) {
    const int b = b0 + m.nbi, c = c0 + m.q * 4;
    fin.sc = lds4(&kc.sc[m.q * 4]); fin.sh = lds4(&kc.sh[m.q * 4]); fin.se = lds4(&kc.se[m.nbi][m.q * 4]);
#pragma unroll 1
    for (int rr = 0; rr < OYC; ++rr) {
    const int cy = m.ty * OYC + rr;                // coarse row inside the tile
    // window: coarse rows cy-1..cy+1 (tile rows cy..cy+2), coarse cols tx*2-1 .. tx*2+2 (tile cols tx*2 .. tx*2+3)
    float4 win[3][OXC + 2];
    const float* base = tile + ((size_t)(m.nbi * t.ih + cy) * t.iw + m.tx * OXC) * t.ps + m.q * 4;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int j = 0; j < OXC + 2; ++j) win[r][j] = lds4(base + (r * t.iw + j) * t.ps);
    float4 acc[OXC][2][2];
#pragma unroll
    for (int i = 0; i < OXC; ++i)
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) acc[i][a][bb] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
      const int a = (ky + P) & 1;                  // input-row parity fed by this tap
      const int dy = (a + P - ky) / 2;             // exact (a + P - ky is even); in [-1, 1]
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const int bb = (kx + P) & 1;
        const int dx = (bb + P - kx) / 2;
        const float4 wv = lds4(s_w + (ky * K + kx) * D2_MAXCG + m.q * 4);
#pragma unroll
        for (int i = 0; i < OXC; ++i) fma4(acc[i][a][bb], win[1 + dy][i + 1 + dx], wv);
      }
    }
    if (b < B && c < C) {
#pragma unroll
      for (int i = 0; i < OXC; ++i)
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int hh = 2 * (ty0 + cy) + a;
          if (hh >= H) continue;
#pragma unroll
          for (int bb = 0; bb < 2; ++bb) {
            const int ww = 2 * (tx0 + m.tx * OXC + i) + bb;
            if (ww >= W) continue;
            fin.apply<T>(acc[i][a][bb], x, gx, (((size_t)b * H + hh) * W + ww) * C + c);
          }
        }
    }
    }
  }
  if (stats) d2_reduce_stats(tile, fin.s1, fin.s2, m.sp, m.q, t.cg, t.tyt * t.txt, t.nb, stats, b0, B, c0, C);
}

// ------------------------------------------------------------------------------------------------
// backward weights: persistent CTAs (grid.x) over (sample block, tile) items of one channel group
// (grid.y).  A thread owns V channels x (OY x OX) outputs; its K*K tap accumulators live in registers
// across all items and leave the CTA once (shuffle -> smem -> one global atomic per tap and channel).
// ------------------------------------------------------------------------------------------------
template <int K, int S> struct D2WGeo {
  static constexpr int V = K == 3 ? 4 : 2;
  static constexpr int OY = S == 1 ? 2 : 1, OX = 4;
};

template <typename T, int K, int S>
__global__ void __launch_bounds__(D2_THREADS, 2)
d2_bwd_weight_kernel(const T* __restrict__ g, const T* __restrict__ yo, const float* __restrict__ alpha,
                     const float* __restrict__ beta, const float* __restrict__ gamma, const T* __restrict__ x, XForm xf,
                     float* __restrict__ dw, int B, int H, int W, int Ho, int Wo, int C, D2Tile t, int items_per_cta) {
  constexpr int P = (K - 1) / 2, V = D2WGeo<K, S>::V, OY = D2WGeo<K, S>::OY, OX = D2WGeo<K, S>::OX;
  constexpr int NR = (OY - 1) * S + K, NC = (OX - 1) * S + K;
  typedef typename VecOf<V>::type VT;
  extern __shared__ __align__(16) float d2_smem[];
  __shared__ D2Consts kc;
  float* xt = d2_smem;                                        // [nb][ih][iw][ps]
  float* gt = xt + (size_t)t.nb * t.ih * t.iw * t.ps;         // [nb][th][tw][ps]
  const int c0 = blockIdx.y * t.cg;
  const int nqv = t.cg / V;
  D2Map m;
  m.init(t, nqv);
  VT acc[K * K];
#pragma unroll
  for (int i = 0; i < K * K; ++i) zerov(acc[i]);
  const int n_items = t.b_blocks * t.tiles_y * t.tiles_x;
  const int it0 = blockIdx.x * items_per_cta;
  const int it1 = min(n_items, it0 + items_per_cta);
  int cur_bb = -1;
  for (int item = it0; item < it1; ++item) {
    const int tx0 = (item % t.tiles_x) * t.tw;
    const int ty0 = ((item / t.tiles_x) % t.tiles_y) * t.th;
    const int bb = item / (t.tiles_x * t.tiles_y);
    const int b0 = bb * t.nb;
    __syncthreads();                                          // previous item's tiles fully consumed
    if (bb != cur_bb) {
      d2_load_consts(kc, xf, alpha, beta, gamma, b0, t.nb, B, c0, C);
      cur_bb = bb;
      __syncthreads();
    }
    d2_stage_x<T>(xt, x, kc, xf.act, xf.se != nullptr, t.ps, t.nv8, b0, B, H, W, C, c0, ty0 * S - P, tx0 * S - P, t.ih,
                  t.iw, t.nb);
    d2_stage_gy<T>(gt, g, yo, kc, t.ps, t.nv8, b0, B, Ho, Wo, C, c0, ty0, tx0, t.th, t.tw, t.nb);
    __syncthreads();
    if (m.active) {
      VT gv[OY][OX];
      const float* gb = gt + ((size_t)(m.nbi * t.th + m.ty * OY) * t.tw + m.tx * OX) * t.ps + m.q * V;
#pragma unroll
      for (int oy = 0; oy < OY; ++oy)
#pragma unroll
        for (int ox = 0; ox < OX; ++ox) ldsv(gv[oy][ox], gb + (oy * t.tw + ox) * t.ps);
      const float* xb = xt + ((size_t)(m.nbi * t.ih + m.ty * OY * S) * t.iw + m.tx * OX * S) * t.ps + m.q * V;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        VT row[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) ldsv(row[j], xb + (r * t.iw + j) * t.ps);
#pragma unroll
        for (int oy = 0; oy < OY; ++oy) {
          const int ky = r - oy * S;
          if (ky < 0 || ky >= K) continue;
#pragma unroll
          for (int kx = 0; kx < K; ++kx)
#pragma unroll
            for (int ox = 0; ox < OX; ++ox) fmav(acc[ky * K + kx], gv[oy][ox], row[ox * S + kx]);
        }
      }
    }
  }
  // reduce over the lanes that share a channel vector, then over the warps, then flush
  __syncthreads();
  float* part = d2_smem;                                      // [8 warps][K*K][cg]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < K * K; ++i) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float v = comp(acc[i], j);
      for (int o = nqv; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < nqv) part[((size_t)warp * K * K + i) * t.cg + lane * V + j] = v;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * t.cg; i += D2_THREADS) {
    float s = 0.f;
#pragma unroll
    for (int wp = 0; wp < D2_THREADS / 32; ++wp) s += part[(size_t)wp * K * K * t.cg + i];
    const int tap = i / t.cg, c = c0 + i % t.cg;
    if (c < C) atomicAdd(&dw[(size_t)c * K * K + tap], s);    // reference layout [C,1,K,K]
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
const int D2_SMEM_CAP_FLOATS = 18 * 1024;   // 72 KB per CTA -> 3 CTAs per SM by shared memory

int pick_cg(int C) {
  const int g32 = ceil_div(C, 32) * 32;
  if (C % 32 == 0 || (double)g32 / C <= 1.07) return 32;
  return 16;
}

// owned grid own_h x own_w; a thread owns oy x ox of it; staged extent of t owned pixels = (t-1)*ss + kk;
// extra_own: also stages a plain owned-size tile (weight gradient: gy next to x_t)
D2Tile pick_tile(int B, int C, int own_h, int own_w, int oy, int ox, int ss, int kk, int extra_own, int nqv_div, int K) {
  D2Tile best = {};
  double best_cost = 1e300;
  const int cg = pick_cg(C);
  const int nsp = D2_THREADS / (cg / nqv_div);
  // pixel stride padded by one float4: staging stores of adjacent pixels and the compute loads of
  // adjacent thread-tiles then fall into disjoint bank groups
  const int ps = cg + 4;
  for (int tyt = 1; tyt <= nsp; ++tyt) {
    if ((tyt - 1) * oy >= own_h) break;
    for (int txt = 1; tyt * txt <= nsp; ++txt) {
      if ((txt - 1) * ox >= own_w) break;
      int nb = nsp / (tyt * txt);
      if (nb > D2_MAXNB) nb = D2_MAXNB;
      if (nb > B) nb = B;
      const int th = tyt * oy, tw = txt * ox;
      const int ih = (th - 1) * ss + kk, iw = (tw - 1) * ss + kk;
      for (; nb >= 1; --nb) {
        const long floats = (long)nb * ((long)ih * iw + (long)extra_own * th * tw) * ps;
        if (floats > D2_SMEM_CAP_FLOATS) continue;
        const int tiles_y = ceil_div(own_h, th), tiles_x = ceil_div(own_w, tw), bbl = ceil_div(B, nb);
        // staged pixels that lie inside the image cost a load + transform; the rest only a store
        const double stage = (double)nb * ((double)ih * iw + (double)extra_own * th * tw);
        const double comp_c = (double)nsp * oy * ox * (2.0 * K * K + 16.0) / 30.0;
        const double cost = (double)tiles_y * tiles_x * bbl * (stage + comp_c + 96.0);
        if (cost < best_cost) {
          best_cost = cost;
          best.cg = cg; best.ps = ps; best.nv8 = cg / 8;
          best.tyt = tyt; best.txt = txt; best.nb = nb; best.th = th; best.tw = tw; best.ih = ih; best.iw = iw;
          best.tiles_y = tiles_y; best.tiles_x = tiles_x; best.b_blocks = bbl; best.n_groups = ceil_div(C, cg);
        }
        break;      // smaller nb only ever costs more for this (tyt, txt)
      }
    }
  }
  return best;
}

size_t tile_bytes(const D2Tile& t, int extra_own, size_t min_floats) {
  size_t f = (size_t)t.nb * ((size_t)t.ih * t.iw + (size_t)extra_own * t.th * t.tw) * t.ps;
  if (f < min_floats) f = min_floats;
  return f * sizeof(float);
}

template <typename KernelT>
int d2_ensure_smem(KernelT kernel) {
  TD3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  return TD3D_OK;
}

template <typename T, int K, int S>
int d2_fwd_t(const DwArgs& a, cudaStream_t st) {
  const int Ho = (a.H - 1) / S + 1, Wo = (a.W - 1) / S + 1;
  const D2Tile t = pick_tile(a.B, a.C, Ho, Wo, D2Geo<S>::OY, D2Geo<S>::OX, S, K, 0, 4, K);
  TD3D_REQUIRE(t.cg > 0, "dw fwd: no tile fits (H=%d W=%d C=%d k=%d s=%d)", a.H, a.W, a.C, K, S);
  static bool once = false;
  if (!once) { TD3D_TRY(d2_ensure_smem(d2_fwd_kernel<T, K, S>)); once = true; }
  const size_t smem = tile_bytes(t, 0, (size_t)(D2_THREADS / (t.cg / 4)) * 2 * t.cg);
  dim3 grid(t.tiles_x * t.tiles_y * t.b_blocks, t.n_groups);
  d2_fwd_kernel<T, K, S><<<grid, D2_THREADS, smem, st>>>((const T*)a.x, a.xf, a.w_taps, (T*)a.y, a.stats, a.B, a.H, a.W, Ho,
                                                         Wo, a.C, t);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int d2_num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T, int K, int S>
int d2_bwd_t(const DwBwdArgs& a, cudaStream_t st) {
  const int Ho = (a.H - 1) / S + 1, Wo = (a.W - 1) / S + 1;
  if (a.gx) {
    if (S == 1) {
      const D2Tile t = pick_tile(a.B, a.C, a.H, a.W, 2, 4, 1, K, 0, 4, K);
      TD3D_REQUIRE(t.cg > 0, "dw bwd-data: no tile fits");
      static bool once = false;
      if (!once) { TD3D_TRY(d2_ensure_smem(d2_bwd_data_s1_kernel<T, K>)); once = true; }
      const size_t smem = tile_bytes(t, 0, (size_t)(D2_THREADS / (t.cg / 4)) * 2 * t.cg);
      dim3 grid(t.tiles_x * t.tiles_y * t.b_blocks, t.n_groups);
      d2_bwd_data_s1_kernel<T, K><<<grid, D2_THREADS, smem, st>>>((const T*)a.g, (const T*)a.y_out, a.alpha, a.beta, a.gamma,
                                                                  (const T*)a.x, a.xf, a.w_taps, (T*)a.gx, a.stats, a.B, a.H,
                                                                  a.W, a.C, t);
    } else {
      // owned grid = coarse grid; a thread owns 2 x 2 coarse positions; staged gy = owned + 1-pixel halo
      const D2Tile t = pick_tile(a.B, a.C, (a.H + 1) / 2, (a.W + 1) / 2, 2, 2, 1, 3, 0, 4, K);
      TD3D_REQUIRE(t.cg > 0, "dw bwd-data: no tile fits");
      static bool once = false;
      if (!once) { TD3D_TRY(d2_ensure_smem(d2_bwd_data_s2_kernel<T, K>)); once = true; }
      const size_t smem = tile_bytes(t, 0, (size_t)(D2_THREADS / (t.cg / 4)) * 2 * t.cg);
      dim3 grid(t.tiles_x * t.tiles_y * t.b_blocks, t.n_groups);
      d2_bwd_data_s2_kernel<T, K><<<grid, D2_THREADS, smem, st>>>((const T*)a.g, (const T*)a.y_out, a.alpha, a.beta, a.gamma,
                                                                  (const T*)a.x, a.xf, a.w_taps, (T*)a.gx, a.stats, a.B, a.H,
                                                                  a.W, Ho, Wo, a.C, t);
    }
    TD3D_LAUNCH_CHECK();
  }
  if (a.dw) {
    constexpr int V = D2WGeo<K, S>::V;
    const D2Tile t = pick_tile(a.B, a.C, Ho, Wo, D2WGeo<K, S>::OY, D2WGeo<K, S>::OX, S, K, 1, V, K);
    TD3D_REQUIRE(t.cg > 0, "dw bwd-weight: no tile fits");
    static bool once = false;
    if (!once) { TD3D_TRY(d2_ensure_smem(d2_bwd_weight_kernel<T, K, S>)); once = true; }
    const size_t smem = tile_bytes(t, 1, (size_t)(D2_THREADS / 32) * K * K * t.cg);
    const int n_items = t.b_blocks * t.tiles_y * t.tiles_x;
    int per = ceil_div(d2_num_sms() * 2, t.n_groups);
    if (per > n_items) per = n_items;
    const int items_per_cta = ceil_div(n_items, per);
    per = ceil_div(n_items, items_per_cta);
    dim3 grid(per, t.n_groups);
    d2_bwd_weight_kernel<T, K, S><<<grid, D2_THREADS, smem, st>>>((const T*)a.g, (const T*)a.y_out, a.alpha, a.beta, a.gamma,
                                                                  (const T*)a.x, a.xf, a.dw, a.B, a.H, a.W, Ho, Wo, a.C, t,
                                                                  items_per_cta);
    TD3D_LAUNCH_CHECK();
  }
  return TD3D_OK;
}

}  // namespace

#define D2_DISPATCH(FN, ARGS)                                                                   \
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

int launch_dw_fwd_v2(const DwArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B > 0, "dw fwd: C=%d must be a multiple of 8", a.C);
  D2_DISPATCH(d2_fwd_t, a);
  set_last_error("dw fwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  return TD3D_EINVAL;
}

int launch_dw_bwd_v2(const DwBwdArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B > 0, "dw bwd: C=%d must be a multiple of 8", a.C);
  D2_DISPATCH(d2_bwd_t, a);
  set_last_error("dw bwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  return TD3D_EINVAL;
}

}  // namespace td3d
