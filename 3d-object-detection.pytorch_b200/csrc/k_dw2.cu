// Depthwise k x k convolution FORWARD, tiled generation (k in {3,5}, stride in {1,2}, pad (k-1)/2, NHWC).
// Reference op: nn.Conv2d(hidden, hidden, k, s, (k-1)//2, groups=hidden) inside InvertedResidual,
// torchdet3d/models/mobilenetv3.py:136,152.  (The tiled backward twins of round 1 were superseded by the one-pass
// column walker of k_dwc.cu and removed.)
//
// The depthwise tensors are the widest of the network, and at B200's HBM rate (~23 B/clk/SM) a 3x3
// depthwise layer has a budget of only ~22 issued instructions per element; a 5x5 layer is FMA-bound
// outright.  ncu on the first tiled generations showed half of all issued instructions going to
// per-CTA set-up (index divisions, constants, weights) and the load latency exposed once per small
// CTA, so these kernels are
//   * PERSISTENT: ~2 CTAs per SM, each walking a contiguous range of (sample block, tile) items of
//     one channel group; thread mapping, item geometry, weights and per-channel constants are set up
//     once per CTA;
//   * PREFETCHED: the raw (bf16/fp32) tile of item i+1 is fetched with cp.async into a private
//     staging area while item i is computed; the producer's lazily applied transform (BatchNorm fold
//     + SE gate + activation) is then evaluated ONCE per staged element into an fp32 shared-memory tile;
//   * a thread owns 4 channels x (2x4 | 1x4) outputs and walks the window row by row with 128-bit
//     conflict-free shared loads; multiply-adds are packed fp32x2 (FFMA2, sm_100);
//   * statistics (BatchNorm sums / SE squeeze) stay in registers across the tiles of a sample block
//     and leave the CTA as one global atomic per channel.
//
//   forward   y  = dw(act(se*(scale*x+shift)))                      + sum y,  sum y^2   per (b,c)
#include "td3d_kernels.h"

#include <stdlib.h>

#include <type_traits>

namespace td3d {

namespace {

constexpr int D2_THREADS = 256;
constexpr int D2_MAXNB = 8;         // samples per CTA tile
constexpr int D2_NIT = 8;           // staged 8-channel vectors per thread and tile (upper bound)
constexpr uint32_t D2_DEAD = 0xffffffffu;

struct D2Tile {
  int cg;                     // channels per CTA (16 or 32; template parameter of the kernels)
  int tyt, txt, nb;           // thread-tiles per CTA (y, x), samples per CTA tile
  int th, tw;                 // owned tile extent (pixels of the owned grid)
  int ih, iw;                 // staged (halo) tile extent
  int tiles_y, tiles_x, b_blocks, n_groups;
  int nit;                    // staged vectors per thread: ceil(nb*ih*iw*(cg/8) / 256)
  int nit2;                   // same for the second (owned-extent) tile of the weight gradient
  int items_per_cta;
};

template <int CG> struct D2C {
  // Staging thread map: a thread stages one 8-channel vector v8 of pixels pix0, pix0 + DPIX, ...
  // CG = 32: v8 fastest (4 lanes = one pixel's 64 contiguous bytes).  CG = 16: v8 slowest, lanes are
  // consecutive pixels, otherwise the two 8-float stores of 4 pixels x 2 vectors collide in the banks.
  static constexpr int DPIX = D2_THREADS / (CG / 8);
  static __device__ __forceinline__ int v8() { return CG == 16 ? (int)threadIdx.x / DPIX : (int)threadIdx.x % (CG / 8); }
  static __device__ __forceinline__ int pix0() { return CG == 16 ? (int)threadIdx.x % DPIX : (int)threadIdx.x / (CG / 8); }
  static constexpr int PS = CG + 4;          // smem pixel stride (floats): +1 float4 spreads the banks of adjacent
                                             // pixels (staging stores) and adjacent thread-tiles (compute loads)
  static constexpr int NV8 = CG / 8;
  static constexpr int NQ = CG / 4;
  static constexpr int NSP = D2_THREADS / NQ;
};

template <int CG> struct D2Consts {
  float sc[CG], sh[CG];
  float se[D2_MAXNB][CG];
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
__device__ __forceinline__ void add4(float4& acc, const float4& a) {
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(acc.x, acc.y));
  const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(acc.z, acc.w));
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

// store 4 channels in the activation dtype; returns the values as stored (rounded) for the statistics
__device__ __forceinline__ float4 st4r(float* p, const float4& v) {
  *reinterpret_cast<float4*>(p) = v;
  return v;
}
__device__ __forceinline__ float4 st4r(bf16* p, const float4& v) {
  uint2 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
  h[0] = __floats2bfloat162_rn(v.x, v.y);
  h[1] = __floats2bfloat162_rn(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = raw;
  return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                     __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
}
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  return make_float4(__uint_as_float(raw.x << 16), __uint_as_float(raw.x & 0xffff0000u),
                     __uint_as_float(raw.y << 16), __uint_as_float(raw.y & 0xffff0000u));
}

// 4 channels as loaded from global memory, kept packed until the epilogue needs them
template <typename T> struct Raw4;
template <> struct Raw4<bf16> {
  uint2 r;
  __device__ __forceinline__ void load(const bf16* p) { r = __ldg(reinterpret_cast<const uint2*>(p)); }
  __device__ __forceinline__ float4 get() const {
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
  }
};
template <> struct Raw4<float> {
  float4 r;
  __device__ __forceinline__ void load(const float* p) { r = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ float4 get() const { return r; }
};

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// one staged 8-channel vector in the thread's private raw area: [u][tid] x sizeof(T)*8 bytes
template <typename T> struct Raw8;
template <> struct Raw8<bf16> {
  static constexpr int BYTES = 16;
  static __device__ __forceinline__ void fetch(uint32_t dst, const bf16* src) { cp_async16(dst, src); }
  static __device__ __forceinline__ void read(const uint8_t* p, float v[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
    v[4] = __uint_as_float(r.z << 16); v[5] = __uint_as_float(r.z & 0xffff0000u);
    v[6] = __uint_as_float(r.w << 16); v[7] = __uint_as_float(r.w & 0xffff0000u);
  }
};
template <> struct Raw8<float> {
  static constexpr int BYTES = 32;
  static __device__ __forceinline__ void fetch(uint32_t dst, const float* src) {
    cp_async16(dst, src);
    cp_async16(dst + 16, src + 4);
  }
  static __device__ __forceinline__ void read(const uint8_t* p, float v[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 16);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// h_swish(u) = u * relu6(u+3)/6 = u * sat(u/6 + 1/2): one FFMA.SAT + one FMUL
__device__ __forceinline__ float act_f(float u, int act) {
  if (act == TD3D_ACT_RELU) return fmaxf(u, 0.f);
  if (act == TD3D_ACT_HSWISH) return u * __saturatef(fmaf(u, 1.f / 6.f, 0.5f));
  return u;
}
// ---- per-CTA constants ------------------------------------------------------------------------
template <int CG>
__device__ __forceinline__ void d2_load_chan_consts(D2Consts<CG>& k, const XForm& xf, int c0, int C) {
  for (int i = threadIdx.x; i < CG; i += D2_THREADS) {
    const int c = c0 + i;
    const bool on = c < C;
    k.sc[i] = (on && xf.scale) ? xf.scale[c] : 1.f;
    k.sh[i] = (on && xf.scale) ? xf.shift[c] : 0.f;
  }
}
template <int CG>
__device__ __forceinline__ void d2_load_sample_consts(D2Consts<CG>& k, const XForm& xf, int b0, int nb, int B, int c0, int C) {
  for (int i = threadIdx.x; i < nb * CG; i += D2_THREADS) {
    const int n = i / CG, cc = i % CG;
    const int c = c0 + cc, b = b0 + n;
    const bool on = c < C && b < B;
    k.se[n][cc] = (on && xf.se) ? xf.se[(size_t)b * C + c] : 1.f;
  }
}

// weights [K*K][C] fp32 -> s_w[K*K][CG] (optionally flipped: tap (ky,kx) <- (K-1-ky, K-1-kx))
template <int K, int CG>
__device__ __forceinline__ void d2_load_w(float* s_w, const float* __restrict__ w, int c0, int C, bool flip) {
  for (int i = threadIdx.x; i < K * K * CG; i += D2_THREADS) {
    const int tap = i / CG, c = c0 + i % CG;
    const int src = flip ? (K * K - 1 - tap) : tap;
    s_w[i] = c < C ? w[(size_t)src * C + c] : 0.f;
  }
}

// ---- staging ----------------------------------------------------------------------------------
// A thread stages one fixed 8-channel vector (v8) of pixels pix0, pix0 + 256/NV8, ... of a
// [nb][rows][cols] tile.  The tile-relative geometry of its items never changes, so it is decoded
// once per CTA into rc[u] = r | c << 8 | nbi << 16.
template <int CG>
__device__ __forceinline__ void d2_items(uint32_t (&rc)[D2_NIT], int rows, int cols, int nb) {
  const int dpix = D2C<CG>::DPIX;
  const int pix0 = D2C<CG>::pix0();
#pragma unroll
  for (int u = 0; u < D2_NIT; ++u) {
    const int pix = pix0 + u * dpix;
    const int c = pix % cols, rr = pix / cols;
    const int r = rr % rows, n = rr / rows;
    rc[u] = n < nb ? (uint32_t)(r | (c << 8) | (n << 16)) : D2_DEAD;
  }
}

// issue the cp.async fetches of one tile (origin (y0,x0) of sample block b0 in an [B,Hs,Ws,C] tensor);
// returns the bit mask of the items that lie inside the tensor
template <typename T, int CG>
__device__ __forceinline__ uint32_t d2_fetch(uint32_t raw_s, const T* __restrict__ src, const T* __restrict__ src2,
                                             uint32_t raw2_s, const uint32_t (&rc)[D2_NIT], int nit, int b0, int B, int Hs,
                                             int Ws, int C, int ch, int y0, int x0) {
  uint32_t mask = 0;
  if (ch < C) {
#pragma unroll
    for (int u = 0; u < D2_NIT; ++u) {
      if (u < nit && rc[u] != D2_DEAD) {
        const int gy = y0 + (int)(rc[u] & 0xffu), gx = x0 + (int)((rc[u] >> 8) & 0xffu), b = b0 + (int)(rc[u] >> 16);
        if ((unsigned)gy < (unsigned)Hs && (unsigned)gx < (unsigned)Ws && b < B) {
          const uint32_t off = ((uint32_t)(b * Hs + gy) * (uint32_t)Ws + (uint32_t)gx) * (uint32_t)C + (uint32_t)ch;
          const uint32_t slot = (uint32_t)(u * D2_THREADS + threadIdx.x) * Raw8<T>::BYTES;
          Raw8<T>::fetch(raw_s + slot, src + off);
          if (src2) Raw8<T>::fetch(raw2_s + slot, src2 + off);
          mask |= 1u << u;
        }
      }
    }
  }
  cp_async_commit();
  return mask;
}

// raw -> tile[nbi][r][c][ch] = act(se*(scale*x+shift)); zero outside the tensor
template <typename T, int CG>
__device__ __forceinline__ void d2_xform_x(float* __restrict__ tile, const uint8_t* __restrict__ raw, uint32_t mask,
                                           const uint32_t (&rc)[D2_NIT], int nit, int rows, int cols,
                                           const D2Consts<CG>& k, int act, bool has_se) {
  const int v8 = D2C<CG>::v8();
  float sc[8], sh[8];                // re-read per tile: not live across the compute phase
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = k.sc[v8 * 8 + i]; sh[i] = k.sh[v8 * 8 + i]; }
#pragma unroll
  for (int u = 0; u < D2_NIT; ++u) {
    if (u < nit && rc[u] != D2_DEAD) {
      const int r = rc[u] & 0xffu, c = (rc[u] >> 8) & 0xffu, n = rc[u] >> 16;
      float* dst = tile + ((n * rows + r) * cols + c) * D2C<CG>::PS + v8 * 8;
      float v[8];
      if ((mask >> u) & 1u) {
        Raw8<T>::read(raw + (size_t)(u * D2_THREADS + threadIdx.x) * Raw8<T>::BYTES, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], sc[i], sh[i]);
        if (has_se) {
          const float4 e0 = lds4(&k.se[n][v8 * 8]), e1 = lds4(&k.se[n][v8 * 8 + 4]);
          v[0] *= e0.x; v[1] *= e0.y; v[2] *= e0.z; v[3] *= e0.w;
          v[4] *= e1.x; v[5] *= e1.y; v[6] *= e1.z; v[7] *= e1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = act_f(v[i], act);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      sts4(dst, v[0], v[1], v[2], v[3]);
      sts4(dst + 4, v[4], v[5], v[6], v[7]);
    }
  }
}

// out[oy][ox] = sum_{ky,kx} tile[(oy*S+ky)][(ox*S+kx)] * w[ky][kx] for a thread's OY x OX outputs
template <int K, int S, int OY, int OX, int CG>
__device__ __forceinline__ void d2_conv(float4 (&acc)[OY][OX], const float* base, int row_stride, const float* s_wq) {
  constexpr int NR = (OY - 1) * S + K, NC = (OX - 1) * S + K, PS = D2C<CG>::PS;
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    float4 row[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) row[j] = lds4(base + r * row_stride + j * PS);
#pragma unroll
    for (int oy = 0; oy < OY; ++oy) {
      const int ky = r - oy * S;
      if (ky < 0 || ky >= K) continue;
#pragma unroll
      for (int kx = 0; kx < K; ++kx) {
        const float4 wv = lds4(s_wq + (ky * K + kx) * CG);
#pragma unroll
        for (int ox = 0; ox < OX; ++ox) fma4(acc[oy][ox], row[ox * S + kx], wv);
      }
    }
  }
}

// per-(sample, channel) sums: thread partials -> smem [sp][2][CG] -> column sums -> one global atomic
template <int CG>
__device__ __forceinline__ void d2_flush_stats(float* part, float4& s1, float4& s2, int sp, int q, int per_sample, int nb,
                                               float* __restrict__ stats, int b0, int B, int c0, int C) {
  float* mine = part + (size_t)sp * 2 * CG + q * 4;
  *reinterpret_cast<float4*>(mine) = s1;
  *reinterpret_cast<float4*>(mine + CG) = s2;
  __syncthreads();
  for (int t = threadIdx.x; t < nb * 2 * CG; t += D2_THREADS) {
    const int n = t / (2 * CG), j = t % (2 * CG);
    float s = 0.f;
    for (int i = 0; i < per_sample; ++i) s += part[(size_t)(n * per_sample + i) * 2 * CG + j];
    const int which = j / CG, c = c0 + j % CG, b = b0 + n;
    if (c < C && b < B) atomicAdd(&stats[((size_t)b * 2 + which) * C + c], s);
  }
  s1 = make_float4(0.f, 0.f, 0.f, 0.f);
  s2 = s1;
}

struct D2Map {       // thread -> (channel vector, sample, thread-tile)
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

struct D2Item {      // walks the CTA's contiguous item range: item = (bb * tiles_y + tyi) * tiles_x + txi
  int bb, tyi, txi;
  __device__ __forceinline__ void init(const D2Tile& t, int item) {
    txi = item % t.tiles_x;
    tyi = (item / t.tiles_x) % t.tiles_y;
    bb = item / (t.tiles_x * t.tiles_y);
  }
  __device__ __forceinline__ void next(const D2Tile& t) {
    if (++txi == t.tiles_x) { txi = 0; if (++tyi == t.tiles_y) { tyi = 0; ++bb; } }
  }
};

template <int S> struct D2Geo { static constexpr int OY = S == 1 ? 2 : 1, OX = 4; };

// dynamic smem: [fp32 tile(s)] [stats partials / weight-gradient partials] [raw area(s)]
struct D2Smem { uint32_t tile_floats, tile2_floats, part_floats, raw_bytes, raw2_bytes; };

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <typename T, int K, int S, int CG>
__global__ void __launch_bounds__(D2_THREADS, 2)
d2_fwd_kernel(const T* __restrict__ x, XForm xf, const float* __restrict__ w, T* __restrict__ y,
              float* __restrict__ stats, int B, int H, int W, int Ho, int Wo, int C, D2Tile t, D2Smem sm) {
  pdl_entry();
  constexpr int P = (K - 1) / 2, OY = D2Geo<S>::OY, OX = D2Geo<S>::OX, PS = D2C<CG>::PS;
  extern __shared__ __align__(16) uint8_t d2_smem[];
  __shared__ D2Consts<CG> kc;
  __shared__ __align__(16) float s_w[K * K * CG];
  float* tile = reinterpret_cast<float*>(d2_smem);
  float* part = tile + sm.tile_floats;
  uint8_t* raw = reinterpret_cast<uint8_t*>(part + sm.part_floats);
  const uint32_t raw_s = s_u32(raw);
  const int c0 = blockIdx.y * CG;
  const int n_items = t.b_blocks * t.tiles_y * t.tiles_x;
  const int it0 = blockIdx.x * t.items_per_cta, it1 = min(n_items, it0 + t.items_per_cta);
  d2_load_chan_consts<CG>(kc, xf, c0, C);
  d2_load_w<K, CG>(s_w, w, c0, C, false);
  uint32_t rc[D2_NIT];
  d2_items<CG>(rc, t.ih, t.iw, t.nb);
  D2Map m;
  m.init(t, D2C<CG>::NQ);
  const int v8 = D2C<CG>::v8(), ch = c0 + v8 * 8;
  __syncthreads();
  const bool has_se = xf.se != nullptr;
  D2Item cur, nxt;
  cur.init(t, it0);
  uint32_t mask = d2_fetch<T, CG>(raw_s, x, nullptr, 0u, rc, t.nit, cur.bb * t.nb, B, H, W, C, ch, cur.tyi * t.th * S - P,
                                  cur.txi * t.tw * S - P);
  float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
  int cur_bb = -1;
  const int c = c0 + m.q * 4;
  const float* s_wq = s_w + m.q * 4;
  const int row_stride = t.iw * PS;
  const float* base = tile + ((m.nbi * t.ih + m.ty * OY * S) * t.iw + m.tx * OX * S) * PS + m.q * 4;
  for (int item = it0; item < it1; ++item) {
    const int b0 = cur.bb * t.nb, ty0 = cur.tyi * t.th, tx0 = cur.txi * t.tw;
    if (cur.bb != cur_bb) {
      if (has_se || (stats && cur_bb >= 0)) __syncthreads();     // previous tile fully consumed (kc.se, part)
      if (stats && cur_bb >= 0)
        d2_flush_stats<CG>(part, s1, s2, m.sp, m.q, t.tyt * t.txt, t.nb, stats, cur_bb * t.nb, B, c0, C);
      if (has_se) d2_load_sample_consts<CG>(kc, xf, b0, t.nb, B, c0, C);
      cur_bb = cur.bb;
    }
    cp_async_wait_all();
    __syncthreads();                                  // previous compute done with the tile; constants visible
    d2_xform_x<T, CG>(tile, raw, mask, rc, t.nit, t.ih, t.iw, kc, xf.act, has_se);
    __syncthreads();
    nxt = cur;
    nxt.next(t);
    if (item + 1 < it1)
      mask = d2_fetch<T, CG>(raw_s, x, nullptr, 0u, rc, t.nit, nxt.bb * t.nb, B, H, W, C, ch, nxt.tyi * t.th * S - P,
                             nxt.txi * t.tw * S - P);
    if (m.active) {
      float4 acc[OY][OX];
#pragma unroll
      for (int i = 0; i < OY; ++i)
#pragma unroll
        for (int j = 0; j < OX; ++j) acc[i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
      d2_conv<K, S, OY, OX, CG>(acc, base, row_stride, s_wq);
      const int b = b0 + m.nbi, yy0 = ty0 + m.ty * OY, xx0 = tx0 + m.tx * OX;
      if (b < B && c < C) {
        T* yb = y + (((size_t)b * Ho + yy0) * Wo + xx0) * C + c;
#pragma unroll
        for (int oy = 0; oy < OY; ++oy) {
          if (yy0 + oy >= Ho) continue;
#pragma unroll
          for (int ox = 0; ox < OX; ++ox) {
            if (xx0 + ox >= Wo) continue;
            const float4 r = st4r(yb + (oy * Wo + ox) * C, acc[oy][ox]);
            add4(s1, r);
            fma4(s2, r, r);
          }
        }
      }
    }
    cur = nxt;
  }
  if (stats && cur_bb >= 0) {
    __syncthreads();
    d2_flush_stats<CG>(part, s1, s2, m.sp, m.q, t.tyt * t.txt, t.nb, stats, cur_bb * t.nb, B, c0, C);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
const long D2_SMEM_CAP_BYTES = 104 * 1024;    // per CTA -> 2 CTAs per SM

int pick_cg(int C) {
  const int g32 = ceil_div(C, 32) * 32;
  if (C % 32 == 0 || (double)g32 / C <= 1.07) return 32;
  return 16;
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

// owned grid own_h x own_w; a thread owns oy x ox of it; staged extent of t owned pixels = (t-1)*ss + kk.
// n_src: tensors fetched for the staged tile (1: x, 2: g and y); extra_own: the weight gradient also
// stages an owned-extent gy tile (2 source tensors) next to x_t.  esz = sizeof(T).
D2Tile pick_tile(int B, int C, int own_h, int own_w, int oy, int ox, int ss, int kk, int n_src, int extra_own, int vec,
                 int K, int esz, int part_floats) {
  D2Tile best = {};
  double best_cost = 1e300;
  const int cg = pick_cg(C);
  const int nv8 = cg / 8, ps = cg + 4;
  const int nsp = D2_THREADS / (cg / vec);
  for (int tyt = 1; tyt <= nsp; ++tyt) {
    if ((tyt - 1) * oy >= own_h) break;
    for (int txt = 1; tyt * txt <= nsp; ++txt) {
      if ((txt - 1) * ox >= own_w) break;
      int nb = nsp / (tyt * txt);
      if (nb > D2_MAXNB) nb = D2_MAXNB;
      if (nb > B) nb = B;
      const int th = tyt * oy, tw = txt * ox;
      const int ih = (th - 1) * ss + kk, iw = (tw - 1) * ss + kk;
      if (ih > 255 || iw > 255) continue;
      for (; nb >= 1; --nb) {
        const int nit = ceil_div((long)nb * ih * iw * nv8, D2_THREADS);
        const int nit2 = extra_own ? ceil_div((long)nb * th * tw * nv8, D2_THREADS) : 0;
        if (nit > D2_NIT || nit2 > D2_NIT) continue;
        const long bytes = 4L * nb * ((long)ih * iw + (long)extra_own * th * tw) * ps + 4L * part_floats +
                           (long)D2_THREADS * esz * 8 * ((long)nit * n_src + (long)nit2 * 2);
        if (bytes > D2_SMEM_CAP_BYTES) continue;
        const int tiles_y = ceil_div(own_h, th), tiles_x = ceil_div(own_w, tw), bbl = ceil_div(B, nb);
        const double stage = (double)nb * ((double)ih * iw + (double)extra_own * th * tw);
        const double comp_c = (double)nsp * oy * ox * (2.0 * K * K + 16.0) / 30.0;
        const double cost = (double)tiles_y * tiles_x * bbl * (stage + comp_c + 48.0);
        if (cost < best_cost) {
          best_cost = cost;
          best.cg = cg;
          best.tyt = tyt; best.txt = txt; best.nb = nb; best.th = th; best.tw = tw; best.ih = ih; best.iw = iw;
          best.tiles_y = tiles_y; best.tiles_x = tiles_x; best.b_blocks = bbl; best.n_groups = ceil_div(C, cg);
          best.nit = nit; best.nit2 = nit2;
        }
        break;      // smaller nb only ever costs more for this (tyt, txt)
      }
    }
  }
  if (best.cg) {
    const int n_items = best.b_blocks * best.tiles_y * best.tiles_x;
    int per = ceil_div(d2_num_sms() * 2, best.n_groups);
    if (per > n_items) per = n_items;
    best.items_per_cta = ceil_div(n_items, per);
  }
  return best;
}

D2Smem smem_layout(const D2Tile& t, int extra_own, int esz, int part_floats) {
  D2Smem s;
  const int ps = t.cg + 4;
  s.tile_floats = (uint32_t)(t.nb * t.ih * t.iw * ps);
  s.tile2_floats = extra_own ? (uint32_t)(t.nb * t.th * t.tw * ps) : 0u;
  s.part_floats = (uint32_t)part_floats;
  s.raw_bytes = (uint32_t)(t.nit * D2_THREADS * esz * 8);
  s.raw2_bytes = (uint32_t)(t.nit2 * D2_THREADS * esz * 8);
  return s;
}
size_t smem_bytes(const D2Smem& s, int n_src) {
  return 4 * ((size_t)s.tile_floats + s.tile2_floats + s.part_floats) + (size_t)s.raw_bytes * n_src + (size_t)s.raw2_bytes * 2;
}

template <typename KernelT>
int d2_ensure_smem(KernelT kernel) {
  TD3D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
  return TD3D_OK;
}

int d2_grid_x(const D2Tile& t) { return ceil_div(t.b_blocks * t.tiles_y * t.tiles_x, t.items_per_cta); }

template <typename T, int K, int S, int CG>
int d2_fwd_launch(const DwArgs& a, const D2Tile& t, int Ho, int Wo, cudaStream_t st) {
  static bool once = false;
  if (!once) { TD3D_TRY(d2_ensure_smem(d2_fwd_kernel<T, K, S, CG>)); once = true; }
  const int part = D2C<CG>::NSP * 2 * CG;
  const D2Smem sm = smem_layout(t, 0, sizeof(T), part);
  dim3 grid(d2_grid_x(t), t.n_groups);
  TD3D_CUDA(launch_kernel(d2_fwd_kernel<T, K, S, CG>, grid, D2_THREADS, smem_bytes(sm, 1), st, (const T*)a.x, a.xf, a.w_taps, (T*)a.y, a.stats, a.B,
                                                                          a.H, a.W, Ho, Wo, a.C, t, sm));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

template <typename T, int K, int S>
int d2_fwd_t(const DwArgs& a, cudaStream_t st) {
  const int Ho = (a.H - 1) / S + 1, Wo = (a.W - 1) / S + 1;
  const int cg = pick_cg(a.C);
  const D2Tile t = pick_tile(a.B, a.C, Ho, Wo, D2Geo<S>::OY, D2Geo<S>::OX, S, K, 1, 0, 4, K, sizeof(T),
                             (D2_THREADS / (cg / 4)) * 2 * cg);
  TD3D_REQUIRE(t.cg > 0, "dw fwd: no tile fits (H=%d W=%d C=%d k=%d s=%d)", a.H, a.W, a.C, K, S);
  if (t.cg == 32) return d2_fwd_launch<T, K, S, 32>(a, t, Ho, Wo, st);
  return d2_fwd_launch<T, K, S, 16>(a, t, Ho, Wo, st);
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

static bool d2_fits_32bit(const char* what, int B, int H, int W, int C) {
  if ((double)B * H * W * C < 2147483648.0) return true;
  set_last_error("%s: tensor of %d x %d x %d x %d elements exceeds the 32-bit offset range", what, B, H, W, C);
  return false;
}

int launch_dw_fwd_v2(const DwArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.B > 0, "dw fwd: C=%d must be a multiple of 8", a.C);
  if (!d2_fits_32bit("dw fwd", a.B, a.H, a.W, a.C)) return TD3D_EINVAL;
  D2_DISPATCH(d2_fwd_t, a);
  set_last_error("dw fwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  return TD3D_EINVAL;
}

}  // namespace td3d
