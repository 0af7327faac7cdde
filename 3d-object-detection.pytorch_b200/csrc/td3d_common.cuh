// Common device/host helpers for the td3d (torchdet3d second-stage regressor) sm_100a kernels.
// Activation tensors are NHWC ("rows" = B*H*W pixels, "channels" contiguous, C % 8 == 0) stored
// in the compute dtype T (float or __nv_bfloat16); all arithmetic is fp32.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "../../include/td3d.h"

namespace td3d {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing (no exceptions cross the C ABI)
// ---------------------------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define TD3D_CUDA(call)                                                         \
  do {                                                                          \
    cudaError_t _e = (call);                                                    \
    if (_e != cudaSuccess) return ::td3d::cuda_fail(_e, #call, __FILE__, __LINE__); \
  } while (0)

#define TD3D_LAUNCH_CHECK() TD3D_CUDA(cudaPeekAtLastError())

#define TD3D_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != TD3D_OK) return _rc; \
  } while (0)

#define TD3D_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      ::td3d::set_last_error(__VA_ARGS__);      \
      return TD3D_EINVAL;                       \
    }                                           \
  } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL) for EAGER launches.  A step is a chain of ~360 dependent kernels, a third of them
// tiny (BatchNorm finalize, SE FCs), and launched one by one it is launch-latency bound: 12.18 ms against 11.69 ms for the
// CUDA-graph replay (MobileNetV3-large, B=256).  Every kernel of this library therefore starts with pdl_entry() -- wait for
// the previous grid to complete and flush, then let the next grid be scheduled -- and every launch goes through
// launch_kernel(), which sets the programmatic-stream-serialization attribute when the previous operation this library
// enqueued on the stream was one of its own kernels.  The next grid's CTAs then become resident while the last wave of the
// current one drains and block in griddepcontrol.wait until it has completed: launch latency and CTA scheduling leave the
// critical path, ordering and memory visibility stay those of a plain stream.  Measured: eager 12.18 -> 11.70 ms, i.e. the
// graph's speed without a graph (variable batch shapes, the first steps before capture).  Under stream capture the
// attribute is NOT set: graph edges are already that cheap (11.65 vs 11.69 ms with programmatic edges, inference 1 % slower),
// so captured graphs keep plain edges.  TD3D_PDL=0 disables it.
// ---------------------------------------------------------------------------------------------
bool pdl_take(cudaStream_t st);    // true: the launch may carry the attribute; either way `st` now ends in one of our kernels
void pdl_break(cudaStream_t st);   // a non-kernel operation (memset, copy, event record / wait) was enqueued on `st`
void pdl_break_all();              // C-ABI entry: other libraries' work may sit between two calls

// ---------------------------------------------------------------------------------------------
// lazily-applied producer transform:  u = se[b,c] * (scale[c]*y + shift[c]);  x = act(u)
//   (BatchNorm fold + SE scale + activation; reference mobilenetv3.py:133-160)
// ---------------------------------------------------------------------------------------------
struct XForm {
  const float* scale;  // [C] or nullptr (=> 1, 0)
  const float* shift;  // [C]
  const float* se;     // [B,C] or nullptr
  int act;             // TD3D_ACT_*
  int se_post;         // 0: x = act(se*(scale*y+shift)) (MobileNetV3 expanded block); 1: x = act(scale*y+shift)*se (SE after the
                       // activation: torchvision MBConv, MobileNetV3 dw-first block) -- honoured by the elementwise kernels only
};

#ifdef __CUDACC__

__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.wait;" ::: "memory");               // no-op for a launch without the attribute
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... P, typename... A>
inline cudaError_t launch_kernel(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess) {      // e.g. the legacy stream while another stream captures: the launch
    cudaGetLastError();                                      // itself will report it; do not leave a stale error behind
    cap = cudaStreamCaptureStatusActive;
  }
  cfg.numAttrs = (cap == cudaStreamCaptureStatusNone && pdl_take(st)) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

__device__ __forceinline__ float act_fwd(float u, int act) {
  if (act == TD3D_ACT_RELU) return fmaxf(u, 0.f);
  if (act == TD3D_ACT_HSWISH) return u * fminf(fmaxf(u + 3.f, 0.f), 6.f) * (1.f / 6.f);
  if (act == TD3D_ACT_SILU) return u * __fdividef(1.f, 1.f + __expf(-u));
  return u;
}
// d act(u) / du, matching autograd of x*relu6(x+3)/6 (hardtanh grad is 0 at both clamp points)
__device__ __forceinline__ float act_bwd(float u, int act) {
  if (act == TD3D_ACT_RELU) return u > 0.f ? 1.f : 0.f;
  if (act == TD3D_ACT_HSWISH) return u <= -3.f ? 0.f : (u >= 3.f ? 1.f : (2.f * u + 3.f) * (1.f / 6.f));
  if (act == TD3D_ACT_SILU) { const float s = __fdividef(1.f, 1.f + __expf(-u)); return s * fmaf(u, 1.f - s, 1.f); }
  return 1.f;
}
// Branch-free activations from per-kernel uniform constants (set up once with make_actk): the `act` switch of
// act_fwd/act_bwd costs two uniform compare+branch pairs PER ELEMENT when it sits in an inner loop (ncu: 30 % of
// the issued instructions of apply_xform).  act(u) = u * sat(a*u + b):  none a=0,b=1 | relu a=2^100,b=0 |
// h_swish a=1/6,b=1/2.   act'(u) = u <= lo ? 0 : (u >= hi ? 1 : da*u + db).
struct ActK { float a, b, da, db, lo, hi; int silu; };   // silu: x*sigmoid(x) has no clamp form -> one uniform branch
__device__ __forceinline__ ActK make_actk(int act) {
  ActK k;
  const bool hs = act == TD3D_ACT_HSWISH, re = act == TD3D_ACT_RELU;
  k.a = hs ? (1.f / 6.f) : (re ? 1.2676506e30f : 0.f);
  k.b = hs ? 0.5f : (re ? 0.f : 1.f);
  k.da = hs ? (1.f / 3.f) : 0.f;
  k.db = hs ? 0.5f : 1.f;
  k.lo = hs ? -3.f : (re ? 0.f : -__int_as_float(0x7f800000));
  k.hi = hs ? 3.f : __int_as_float(0x7f800000);
  k.silu = act == TD3D_ACT_SILU;
  return k;
}
__device__ __forceinline__ float actk_fwd(float u, const ActK& k) {
  if (k.silu) return u * __fdividef(1.f, 1.f + __expf(-u));     // MUFU.EX2 + MUFU.RCP, no full-precision division slow path
  return u * __saturatef(fmaf(k.a, u, k.b));
}
__device__ __forceinline__ float actk_bwd(float u, const ActK& k) {
  if (k.silu) { const float s = __fdividef(1.f, 1.f + __expf(-u)); return s * fmaf(u, 1.f - s, 1.f); }
  float d = fmaf(u, k.da, k.db);
  d = u >= k.hi ? 1.f : d;
  return u <= k.lo ? 0.f : d;
}
__device__ __forceinline__ float hsigmoid(float u) { return fminf(fmaxf(u + 3.f, 0.f), 6.f) * (1.f / 6.f); }
__device__ __forceinline__ float hsigmoid_bwd(float u) { return (u > -3.f && u < 3.f) ? (1.f / 6.f) : 0.f; }

// ---- 8-channel vector access ------------------------------------------------------------------
__device__ __forceinline__ void load8(const float* p, float v[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float v[8]) {
  uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float v[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float v[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = raw;
}
__device__ __forceinline__ void loadf8(const float* p, float v[8]) { load8(p, v); }

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

// apply XForm to 8 consecutive channels starting at c of sample b (C = channel count)
__device__ __forceinline__ void xform8(float v[8], const XForm& xf, int b, int c, int C) {
  if (xf.scale) {
    float s[8], t[8];
    loadf8(xf.scale + c, s);
    loadf8(xf.shift + c, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], s[i], t[i]);
  }
  if (xf.se) {
    float e[8];
    loadf8(xf.se + (size_t)b * C + c, e);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= e[i];
  }
  if (xf.act != TD3D_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = act_fwd(v[i], xf.act);
  }
}
// u (pre-activation) for 8 channels; used by backward kernels to evaluate act'(u)
__device__ __forceinline__ void preact8(float v[8], const XForm& xf, int b, int c, int C) {
  if (xf.scale) {
    float s[8], t[8];
    loadf8(xf.scale + c, s);
    loadf8(xf.shift + c, t);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], s[i], t[i]);
  }
  if (xf.se) {
    float e[8];
    loadf8(xf.se + (size_t)b * C + c, e);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= e[i];
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

}  // namespace td3d
