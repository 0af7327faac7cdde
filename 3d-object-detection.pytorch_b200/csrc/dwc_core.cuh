// Depthwise k x k convolution BACKWARD in ONE pass ("column walker"), NHWC.
// Reference op: autograd backward of nn.Conv2d(hidden, hidden, k, s, (k-1)//2, groups=hidden) inside
// InvertedResidual (torchdet3d/models/mobilenetv3.py:136,152), fused with the BatchNorm-backward affine of the
// incoming gradient and the activation backward of the layer's input:
//
//   gy  = alpha[b,c]*g + beta[c]*y_out + gamma[b,c]                  (zero outside the output plane)
//   t   = se[b,c]*(scale[c]*x + shift[c]);  xa = act(t)
//   gx  = act'(t) * sum_{i,j} w[i][j] * gy[(q + pad - (i,j)) / s]    (data gradient, written once)
//   dW[c][i][j] += sum_q xa[q] * gy[(q + pad - (i,j)) / s]           (weight gradient)
//   stats += (sum gx, sum gx*x)                                       (BatchNorm backward of the producer of x)
//
// Round 1 ran this as two kernels (data / weight gradient) that each re-read g, y_out and x: 1.66x the algorithmic
// DRAM bytes and ~40 thread-instructions per element.  Here every thread owns CPT adjacent channels (lanes of a warp
// are channel-contiguous: a warp request is one 128-byte line) and a band of R output rows (R*S input rows) of one
// sample, and WALKS THE COLUMNS of the plane: the gradient window (R+lo+hi rows x 1+lo+hi columns) slides in
// registers, the k*k taps and the k*k weight-gradient accumulators stay in registers for the thread's whole life
// (it loops over many (sample, band) items), g / y_out / x are each loaded once per band (halo rows come from
// L1/L2), and no shared memory or barrier is used in the main loop.
//
// Inner-loop hygiene (the first version spent ~250 SASS instructions per output element, most of them 64-bit address
// arithmetic, row/column predicates with divergence barriers and an inlined SiLU slow path): per item the row offsets are
// computed ONCE as clamped 32-bit element offsets with a 0/1 float mask per row, a step adds one clamped column offset,
// out-of-range taps are loaded from the clamped (valid) address and multiplied by the mask -- no branches in the walk.
//
// The per-thread body is __host__ __device__ so tests/host/dwc_emul.cu can run the exact index logic on the CPU.
#pragma once
#include <math.h>
#include <string.h>

#include "td3d_common.cuh"

namespace td3d {

struct DwcArgs {
  const void* g; const void* y_out;          // [B,Ho,Wo,C]
  const float* alpha; const float* beta; const float* gamma;   // [B,C], [C], [B,C]
  const void* x;                              // [B,H,W,C] raw forward input
  const float* scale; const float* shift;     // [C] or null (identity)
  const float* se;                            // [B,C] or null
  int act;
  const float* w_taps;                        // [K*K][C]
  void* gx;                                   // [B,H,W,C]
  float* stats;                               // [slots][2][C] or null (only the sum over slots is meaningful)
  float* dw;                                  // [C][K*K] (+=)
  int B, H, W, C, Ho, Wo;
  int slots;
  int n_bands, n_items, item_lanes;           // items = B * n_bands; lane il handles items il, il + item_lanes, ...
  int cw;                                     // channel groups (of CPT channels) per block column chunk
  int n_cchunks;                              // ceil((C / CPT) / cw)
  int ilb;                                    // item lanes per block
};

// ---- CPT-wide fp32 vectors -------------------------------------------------------------------------------------
template <int CPT> struct DwcVec;
template <> struct DwcVec<2> { typedef float2 V; };
template <> struct DwcVec<1> { typedef float V; };

__host__ __device__ __forceinline__ float2 dv_fma(float2 a, float2 b, float2 c) {
#ifdef __CUDA_ARCH__
  return __ffma2_rn(a, b, c);
#else
  return make_float2(a.x * b.x + c.x, a.y * b.y + c.y);
#endif
}
__host__ __device__ __forceinline__ float dv_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return fmaf(a, b, c);
#else
  return a * b + c;
#endif
}
__host__ __device__ __forceinline__ float2 dv_mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__host__ __device__ __forceinline__ float dv_mul(float a, float b) { return a * b; }
__host__ __device__ __forceinline__ float2 dv_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float dv_add(float a, float b) { return a + b; }
__host__ __device__ __forceinline__ void dv_set(float2& v, float s) { v = make_float2(s, s); }
__host__ __device__ __forceinline__ void dv_set(float& v, float s) { v = s; }
__host__ __device__ __forceinline__ float2 dv_scale(float2 a, float m) { return make_float2(a.x * m, a.y * m); }
__host__ __device__ __forceinline__ float dv_scale(float a, float m) { return a * m; }
__host__ __device__ __forceinline__ float dv_get(const float2& v, int i) { return i ? v.y : v.x; }
__host__ __device__ __forceinline__ float dv_get(const float& v, int) { return v; }

// Activations are COMPILE-TIME (template int ACT = TD3D_ACT_*): the first versions carried the kind as run-time constants,
// which cost seven registers, a uniform branch per element for SiLU and the dead code of every other activation in each
// kernel -- at 128 registers per thread the compiler then re-derived row offsets inside the walk instead of keeping them.
__host__ __device__ __forceinline__ float dwc_sat(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }
__host__ __device__ __forceinline__ float dwc_sigmoid(float u) {
#ifdef __CUDA_ARCH__
  return __fdividef(1.f, 1.f + __expf(-u));
#else
  return 1.f / (1.f + expf(-u));
#endif
}
template <int ACT> __host__ __device__ __forceinline__ float dwc_act1(float u) {
  if (ACT == TD3D_ACT_RELU) return u > 0.f ? u : 0.f;
  if (ACT == TD3D_ACT_HSWISH) return u * dwc_sat(u * (1.f / 6.f) + 0.5f);
  if (ACT == TD3D_ACT_SILU) return u * dwc_sigmoid(u);
  return u;
}
template <int ACT> __host__ __device__ __forceinline__ float dwc_actd1(float u) {
  if (ACT == TD3D_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (ACT == TD3D_ACT_HSWISH) return u <= -3.f ? 0.f : (u >= 3.f ? 1.f : u * (1.f / 3.f) + 0.5f);
  if (ACT == TD3D_ACT_SILU) { const float s = dwc_sigmoid(u); return s * (1.f + u * (1.f - s)); }
  return 1.f;
}
template <int ACT> __host__ __device__ __forceinline__ float2 dwc_act(float2 u) { return make_float2(dwc_act1<ACT>(u.x), dwc_act1<ACT>(u.y)); }
template <int ACT> __host__ __device__ __forceinline__ float dwc_act(float u) { return dwc_act1<ACT>(u); }
template <int ACT> __host__ __device__ __forceinline__ float2 dwc_actd(float2 u) { return make_float2(dwc_actd1<ACT>(u.x), dwc_actd1<ACT>(u.y)); }
template <int ACT> __host__ __device__ __forceinline__ float dwc_actd(float u) { return dwc_actd1<ACT>(u); }

// ---- raw loads / stores of CPT channels in the activation dtype ----------------------------------------------
template <typename T, int CPT> struct DwcIo;
template <> struct DwcIo<bf16, 2> {
  typedef uint32_t Raw;
  static __host__ __device__ __forceinline__ Raw zero() { return 0u; }
  static __host__ __device__ __forceinline__ Raw ld(const bf16* p) {
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const unsigned int*>(p));
#else
    return *reinterpret_cast<const uint32_t*>(p);
#endif
  }
  static __host__ __device__ __forceinline__ float2 cvt(Raw r) {
#ifdef __CUDA_ARCH__
    return make_float2(__uint_as_float(r << 16), __uint_as_float(r & 0xffff0000u));
#else
    uint32_t lo = r << 16, hi = r & 0xffff0000u;
    float2 v;
    memcpy(&v.x, &lo, 4); memcpy(&v.y, &hi, 4);
    return v;
#endif
  }
  // stores v rounded to bf16 (RNE) and returns the rounded values
  static __host__ __device__ __forceinline__ float2 st(bf16* p, float2 v) {
#ifdef __CUDA_ARCH__
    __nv_bfloat162 h = __floats2bfloat162_rn(v.x, v.y);
    *reinterpret_cast<__nv_bfloat162*>(p) = h;
    return __bfloat1622float2(h);
#else
    p[0] = __float2bfloat16_rn(v.x); p[1] = __float2bfloat16_rn(v.y);
    return make_float2(__bfloat162float(p[0]), __bfloat162float(p[1]));
#endif
  }
};
template <> struct DwcIo<bf16, 1> {
  typedef uint16_t Raw;
  static __host__ __device__ __forceinline__ Raw zero() { return 0; }
  static __host__ __device__ __forceinline__ Raw ld(const bf16* p) {
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const unsigned short*>(p));
#else
    return *reinterpret_cast<const uint16_t*>(p);
#endif
  }
  static __host__ __device__ __forceinline__ float cvt(Raw r) {
    uint32_t u = (uint32_t)r << 16;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
  }
  static __host__ __device__ __forceinline__ float st(bf16* p, float v) {
    bf16 h = __float2bfloat16_rn(v);
    *p = h;
    return __bfloat162float(h);
  }
};
template <> struct DwcIo<float, 2> {
  typedef float2 Raw;
  static __host__ __device__ __forceinline__ Raw zero() { return make_float2(0.f, 0.f); }
  static __host__ __device__ __forceinline__ Raw ld(const float* p) {
#ifdef __CUDA_ARCH__
    return __ldg(reinterpret_cast<const float2*>(p));
#else
    return *reinterpret_cast<const float2*>(p);
#endif
  }
  static __host__ __device__ __forceinline__ float2 cvt(Raw r) { return r; }
  static __host__ __device__ __forceinline__ float2 st(float* p, float2 v) { *reinterpret_cast<float2*>(p) = v; return v; }
};
template <> struct DwcIo<float, 1> {
  typedef float Raw;
  static __host__ __device__ __forceinline__ Raw zero() { return 0.f; }
  static __host__ __device__ __forceinline__ Raw ld(const float* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
  }
  static __host__ __device__ __forceinline__ float cvt(Raw r) { return r; }
  static __host__ __device__ __forceinline__ float st(float* p, float v) { *p = v; return v; }
};

template <int CPT> __host__ __device__ __forceinline__ typename DwcVec<CPT>::V dwc_ldc(const float* p);
template <> __host__ __device__ __forceinline__ float2 dwc_ldc<2>(const float* p) { return make_float2(p[0], p[1]); }
template <> __host__ __device__ __forceinline__ float dwc_ldc<1>(const float* p) { return p[0]; }

// ---------------------------------------------------------------------------------------------------------------
// One thread's whole life.  `c` = first of its CPT channels, `il` = item lane.  `Sink` receives the final sums:
//   sink.dw(c + i, tap, value), sink.stat(which, c + i, value)
// ---------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int dwc_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// Pins a per-item value in a register: without it the compiler re-derives 64-bit base pointers and row offsets from the
// kernel arguments inside the walk (~130 integer instructions per step in the SASS of the first versions).
template <typename P> __host__ __device__ __forceinline__ void dwc_pin(P*& p) {
#ifdef __CUDA_ARCH__
  asm volatile("" : "+l"(p));
#endif
}
__host__ __device__ __forceinline__ void dwc_pin(uint32_t& v) {
#ifdef __CUDA_ARCH__
  asm volatile("" : "+r"(v));
#endif
}

template <typename T, int K, int S, int R, int CPT, int ACT>
struct DwcBwd {
  typedef typename DwcVec<CPT>::V V;
  typedef DwcIo<T, CPT> Io;
  typedef typename Io::Raw Raw;
  static constexpr int PAD = (K - 1) / 2;
  static constexpr int LO = PAD / S;                 // gradient halo (rows / columns) before the band
  static constexpr int HI = (S - 1 + PAD) / S;       // ... after it
  static constexpr int NA = R + LO + HI;             // gradient window rows
  static constexpr int NB = 1 + LO + HI;             // gradient window columns
  static constexpr int RI = R * S;                   // input rows of a band

  struct State {
    V G[NA][NB];        // gradient window; column slot (wb + phase) % NB holds gy column px - LO + wb
    V wt[K * K];        // taps
    V dwa[K * K];       // weight-gradient accumulators
    V s1, s2;           // sum gx, sum gx * x
    Raw rg[NA], ry[NA]; // prefetched raw gradient column
    Raw rx[RI][S];      // prefetched raw input pixels of the next step
    uint32_t og[NA];    // per item: clamped element offset of gradient row ar (row * Wo * C)
    uint32_t ox[RI];    // ... of input row ey
    float mg[NA];       // 1 inside the output plane, 0 outside
    float mx[RI];
  };

  // raw loads step `px` needs: gradient column px + HI, input columns S*px .. S*px + S-1 (clamped addresses: always valid)
  static __host__ __device__ __forceinline__ void prefetch(State& st, const DwcArgs& a, const T* gb, const T* yb,
                                                           const T* xb, int px) {
    const uint32_t cgo = (uint32_t)(dwc_clampi(px + HI, 0, a.Wo - 1) * a.C);
#pragma unroll
    for (int ar = 0; ar < NA; ++ar) {
      st.rg[ar] = Io::ld(gb + (st.og[ar] + cgo));
      st.ry[ar] = Io::ld(yb + (st.og[ar] + cgo));
    }
#pragma unroll
    for (int ex = 0; ex < S; ++ex) {
      const uint32_t cxo = (uint32_t)(dwc_clampi(S * px + ex, 0, a.W - 1) * a.C);
#pragma unroll
      for (int ey = 0; ey < RI; ++ey) st.rx[ey][ex] = Io::ld(xb + (st.ox[ey] + cxo));
    }
  }

  template <int PH>
  static __host__ __device__ __forceinline__ void step(State& st, const DwcArgs& a, const T* gb, const T* yb, const T* xb,
                                                       T* ob, int px, V al, V be, V ga, V sc, V sh) {
    // 1. the prefetched gradient column enters the window (slot of column px + HI), zero outside the plane
    constexpr int SLOT_NEW = (NB - 1 + PH) % NB;
    {
      const int cg = px + HI;
      const float cm = (cg >= 0 && cg < a.Wo) ? 1.f : 0.f;
#pragma unroll
      for (int ar = 0; ar < NA; ++ar)
        st.G[ar][SLOT_NEW] = dv_scale(dv_fma(al, Io::cvt(st.rg[ar]), dv_fma(be, Io::cvt(st.ry[ar]), ga)), st.mg[ar] * cm);
    }
    V xv[RI][S];
#pragma unroll
    for (int ey = 0; ey < RI; ++ey)
#pragma unroll
      for (int ex = 0; ex < S; ++ex) xv[ey][ex] = Io::cvt(st.rx[ey][ex]);
    // 2. next step's loads go out before this step's math
    prefetch(st, a, gb, yb, xb, px + 1);
    if (px < 0) return;
    // 3. every input pixel of this step: row S*r0 + ey, column S*px + ex
#pragma unroll
    for (int ex = 0; ex < S; ++ex) {
      const int qx = S * px + ex;
      const bool col_ok = qx < a.W;
      const uint32_t cxo = (uint32_t)(dwc_clampi(qx, 0, a.W - 1) * a.C);
#pragma unroll
      for (int ey = 0; ey < RI; ++ey) {
        const float vm = col_ok ? st.mx[ey] : 0.f;
        const V xr = xv[ey][ex];
        const V t = dv_fma(sc, xr, sh);                 // the SE gate (if any) is folded into sc / sh per item
        const V xa = dv_scale(dwc_act<ACT>(t), vm);
        const V da = dv_scale(dwc_actd<ACT>(t), vm);
        V acc;
        dv_set(acc, 0.f);
        const int r = ey / S, e = ey % S;
#pragma unroll
        for (int ar = 0; ar < NA; ++ar) {
          const int i = S * r + e + PAD - S * (ar - LO);          // tap row that links input row qy to window row ar
          if (i < 0 || i >= K) continue;
#pragma unroll
          for (int wb = 0; wb < NB; ++wb) {
            const int j = ex + PAD - S * (wb - LO);
            if (j < 0 || j >= K) continue;
            const V gv = st.G[ar][(wb + PH) % NB];
            acc = dv_fma(st.wt[i * K + j], gv, acc);
            st.dwa[i * K + j] = dv_fma(xa, gv, st.dwa[i * K + j]);
          }
        }
        const V gxv = dv_mul(acc, da);
        V gxr = gxv;
        if (vm != 0.f) gxr = Io::st(ob + (st.ox[ey] + cxo), gxv);
        st.s1 = dv_add(st.s1, gxr);
        st.s2 = dv_fma(gxr, xr, st.s2);
      }
    }
  }

  static __host__ __device__ __forceinline__ void run_phases(State& st, const DwcArgs& a, const T* gb, const T* yb,
                                                             const T* xb, T* ob, int px0, int px_end, V al, V be,
                                                             V ga, V sc, V sh) {
    if (px0 + 0 < px_end) step<0>(st, a, gb, yb, xb, ob, px0 + 0, al, be, ga, sc, sh);
    if (NB > 1 && px0 + 1 < px_end) step<(NB > 1 ? 1 : 0)>(st, a, gb, yb, xb, ob, px0 + 1, al, be, ga, sc, sh);
    if (NB > 2 && px0 + 2 < px_end) step<(NB > 2 ? 2 : 0)>(st, a, gb, yb, xb, ob, px0 + 2, al, be, ga, sc, sh);
    if (NB > 3 && px0 + 3 < px_end) step<(NB > 3 ? 3 : 0)>(st, a, gb, yb, xb, ob, px0 + 3, al, be, ga, sc, sh);
    if (NB > 4 && px0 + 4 < px_end) step<(NB > 4 ? 4 : 0)>(st, a, gb, yb, xb, ob, px0 + 4, al, be, ga, sc, sh);
  }

  template <class Sink>
  static __host__ __device__ void thread_main(const DwcArgs& a, int c, int il, Sink& sink) {
    const T* g = reinterpret_cast<const T*>(a.g);
    const T* y = reinterpret_cast<const T*>(a.y_out);
    const T* x = reinterpret_cast<const T*>(a.x);
    T* gx = reinterpret_cast<T*>(a.gx);
    State st;
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      st.wt[t] = dwc_ldc<CPT>(a.w_taps + (size_t)t * a.C + c);
      dv_set(st.dwa[t], 0.f);
    }
    dv_set(st.s1, 0.f); dv_set(st.s2, 0.f);
    V be = dwc_ldc<CPT>(a.beta + c), sc, sh;
    dv_set(sc, 1.f); dv_set(sh, 0.f);
    if (a.scale) { sc = dwc_ldc<CPT>(a.scale + c); sh = dwc_ldc<CPT>(a.shift + c); }
    V sc0 = sc, sh0 = sh;
    for (int item = il; item < a.n_items; item += a.item_lanes) {
      const int b = item / a.n_bands, band = item - b * a.n_bands;
      const int r0 = band * R;
      const V al = dwc_ldc<CPT>(a.alpha + (size_t)b * a.C + c), ga = dwc_ldc<CPT>(a.gamma + (size_t)b * a.C + c);
      if (a.se) {                                       // u = se*(scale*x + shift) = (se*scale)*x + se*shift
        const V se = dwc_ldc<CPT>(a.se + (size_t)b * a.C + c);
        sc = dv_mul(se, sc0); sh = dv_mul(se, sh0);
      }
      const T* gb = g + (size_t)b * a.Ho * a.Wo * a.C + c;
      const T* yb = y + (size_t)b * a.Ho * a.Wo * a.C + c;
      const T* xb = x + (size_t)b * a.H * a.W * a.C + c;
      T* ob = gx + (size_t)b * a.H * a.W * a.C + c;
      dwc_pin(gb); dwc_pin(yb); dwc_pin(xb); dwc_pin(ob);
#pragma unroll
      for (int ar = 0; ar < NA; ++ar) {
        const int py = r0 + ar - LO;
        st.og[ar] = (uint32_t)(dwc_clampi(py, 0, a.Ho - 1) * a.Wo * a.C);
        dwc_pin(st.og[ar]);
        st.mg[ar] = (py >= 0 && py < a.Ho) ? 1.f : 0.f;
      }
#pragma unroll
      for (int ey = 0; ey < RI; ++ey) {
        const int qy = S * r0 + ey;
        st.ox[ey] = (uint32_t)(dwc_clampi(qy, 0, a.H - 1) * a.W * a.C);
        dwc_pin(st.ox[ey]);
        st.mx[ey] = qy < a.H ? 1.f : 0.f;
      }
      // window fill: steps px = -(LO+HI) .. -1 only shift columns in; compute starts at px = 0
      const int px_begin = -(LO + HI);
      prefetch(st, a, gb, yb, xb, px_begin);
#pragma unroll 1
      for (int px0 = px_begin; px0 < a.Wo; px0 += NB)      // NB phases with compile-time window slots (no register moves)
        run_phases(st, a, gb, yb, xb, ob, px0, a.Wo, al, be, ga, sc, sh);
    }
#pragma unroll
    for (int t = 0; t < K * K; ++t)
#pragma unroll
      for (int i = 0; i < CPT; ++i) sink.dw(c + i, t, dv_get(st.dwa[t], i));
    if (a.stats != nullptr) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        sink.stat(0, c + i, dv_get(st.s1, i));
        sink.stat(1, c + i, dv_get(st.s2, i));
      }
    }
  }
};

// =================================================================================================================
// FORWARD column walker:  y = out_act( dw(act(se*(scale*x+shift))) + out_bias )   (+ per-sample sums of y, y^2)
//   training  : input transform = BatchNorm fold (+SE gate) + activation of the producer, raw y + sums for this BatchNorm
//   inference : BatchNorm folded into the taps / out_bias, activation applied before the single store
// Thread = CPT channels x a band of R output rows; the K-column input window slides in registers (slot of an input
// column is compile time after unrolling the walk K times), every input element is loaded and transformed once per band.
// =================================================================================================================
struct DwcFwdArgs {
  const void* x;                              // [B,H,W,C]
  const float* scale; const float* shift;     // [C] or null
  const float* se;                            // [B,C] or null
  int act;
  const float* w_taps;                        // [K*K][C]
  const float* out_bias;                      // [C] or null
  int out_act;
  void* y;                                    // [B,Ho,Wo,C]
  float* stats;                               // [B][2][C] per-sample sums (SE squeeze needs them per sample) or null
  int B, H, W, C, Ho, Wo;
  int n_bands, n_items, item_lanes;
  int cw, n_cchunks, ilb;
};

template <typename T, int K, int S, int R, int CPT, int ACT, int OACT>
struct DwcFwd {
  typedef typename DwcVec<CPT>::V V;
  typedef DwcIo<T, CPT> Io;
  typedef typename Io::Raw Raw;
  static constexpr int PAD = (K - 1) / 2;
  static constexpr int NAI = S * (R - 1) + K;        // input rows of a band
  static constexpr int PXB = -((K - 1) / S);         // first (fill) step: every window column of step 0 has entered by then

  struct State {
    V Wn[NAI][K];       // transformed input window
    V wt[K * K];
    Raw rx[NAI][S];     // prefetched raw input columns of the next step
    uint32_t ox[NAI];   // per item: clamped element offset of input row ai
    uint32_t oy[R];     // ... of output row r
    float mx[NAI];      // 1 inside the image, 0 in the padding
  };

  static __host__ __device__ __forceinline__ void prefetch(State& st, const DwcFwdArgs& a, const T* xb, int px) {
#pragma unroll
    for (int e = 0; e < S; ++e) {
      const uint32_t cxo = (uint32_t)(dwc_clampi(S * px + PAD - S + 1 + e, 0, a.W - 1) * a.C);
#pragma unroll
      for (int ai = 0; ai < NAI; ++ai) st.rx[ai][e] = Io::ld(xb + (st.ox[ai] + cxo));
    }
  }

  template <int PH>
  static __host__ __device__ __forceinline__ void step(State& st, const DwcFwdArgs& a, const T* xb, T* ob, int r0, int px,
                                                       V sc, V sh, V ob_bias, V& s1, V& s2) {
    // 1. the prefetched columns enter the window, transformed once (zero outside the image: conv padding)
#pragma unroll
    for (int e = 0; e < S; ++e) {
      const int qx = S * px + PAD - S + 1 + e;
      const float cm = (qx >= 0 && qx < a.W) ? 1.f : 0.f;
#pragma unroll
      for (int ai = 0; ai < NAI; ++ai)
        st.Wn[ai][(S * PH + e) % K] = dv_scale(dwc_act<ACT>(dv_fma(sc, Io::cvt(st.rx[ai][e]), sh)), st.mx[ai] * cm);
    }
    prefetch(st, a, xb, px + 1);
    if (px < 0) return;
    // 2. R outputs of this column
    const uint32_t cyo = (uint32_t)(px * a.C);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      V acc = ob_bias;
#pragma unroll
      for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) acc = dv_fma(st.wt[i * K + j], st.Wn[S * r + i][(S * PH + j + S) % K], acc);
      if (r0 + r < a.Ho) {
        const V yr = Io::st(ob + (st.oy[r] + cyo), dwc_act<OACT>(acc));
        s1 = dv_add(s1, yr);
        s2 = dv_fma(yr, yr, s2);
      }
    }
  }

  static __host__ __device__ __forceinline__ void run_phases(State& st, const DwcFwdArgs& a, const T* xb, T* ob, int r0,
                                                             int px0, V sc, V sh, V obb, V& s1, V& s2) {
    if (px0 + 0 < a.Wo) step<0>(st, a, xb, ob, r0, px0 + 0, sc, sh, obb, s1, s2);
    if (K > 1 && px0 + 1 < a.Wo) step<(K > 1 ? 1 : 0)>(st, a, xb, ob, r0, px0 + 1, sc, sh, obb, s1, s2);
    if (K > 2 && px0 + 2 < a.Wo) step<(K > 2 ? 2 : 0)>(st, a, xb, ob, r0, px0 + 2, sc, sh, obb, s1, s2);
    if (K > 3 && px0 + 3 < a.Wo) step<(K > 3 ? 3 : 0)>(st, a, xb, ob, r0, px0 + 3, sc, sh, obb, s1, s2);
    if (K > 4 && px0 + 4 < a.Wo) step<(K > 4 ? 4 : 0)>(st, a, xb, ob, r0, px0 + 4, sc, sh, obb, s1, s2);
  }

  // `Sink::stat(b, which, c, value)` receives the per-sample sums of one item
  template <class Sink>
  static __host__ __device__ void thread_main(const DwcFwdArgs& a, int c, int il, Sink& sink) {
    const T* x = reinterpret_cast<const T*>(a.x);
    T* y = reinterpret_cast<T*>(a.y);
    State st;
#pragma unroll
    for (int t = 0; t < K * K; ++t) st.wt[t] = dwc_ldc<CPT>(a.w_taps + (size_t)t * a.C + c);
    V sc, sh, obb;
    dv_set(sc, 1.f); dv_set(sh, 0.f); dv_set(obb, 0.f);
    if (a.scale) { sc = dwc_ldc<CPT>(a.scale + c); sh = dwc_ldc<CPT>(a.shift + c); }
    if (a.out_bias) obb = dwc_ldc<CPT>(a.out_bias + c);
    V sc0 = sc, sh0 = sh;
    for (int item = il; item < a.n_items; item += a.item_lanes) {
      const int b = item / a.n_bands, band = item - b * a.n_bands;
      const int r0 = band * R;
      if (a.se) {                                       // u = se*(scale*x + shift) = (se*scale)*x + se*shift
        const V se = dwc_ldc<CPT>(a.se + (size_t)b * a.C + c);
        sc = dv_mul(se, sc0); sh = dv_mul(se, sh0);
      }
      const T* xb = x + (size_t)b * a.H * a.W * a.C + c;
      T* ob = y + (size_t)b * a.Ho * a.Wo * a.C + c;
      dwc_pin(xb); dwc_pin(ob);
#pragma unroll
      for (int ai = 0; ai < NAI; ++ai) {
        const int qy = S * r0 - PAD + ai;
        st.ox[ai] = (uint32_t)(dwc_clampi(qy, 0, a.H - 1) * a.W * a.C);
        dwc_pin(st.ox[ai]);
        st.mx[ai] = (qy >= 0 && qy < a.H) ? 1.f : 0.f;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) { st.oy[r] = (uint32_t)(dwc_clampi(r0 + r, 0, a.Ho - 1) * a.Wo * a.C); dwc_pin(st.oy[r]); }
      V s1, s2;
      dv_set(s1, 0.f); dv_set(s2, 0.f);
      prefetch(st, a, xb, PXB);
#pragma unroll 1
      for (int px0 = PXB; px0 < a.Wo; px0 += K) run_phases(st, a, xb, ob, r0, px0, sc, sh, obb, s1, s2);
      if (a.stats) {
#pragma unroll
        for (int i = 0; i < CPT; ++i) {
          sink.stat(b, 0, c + i, dv_get(s1, i));
          sink.stat(b, 1, c + i, dv_get(s2, i));
        }
      }
    }
  }
};

}  // namespace td3d
