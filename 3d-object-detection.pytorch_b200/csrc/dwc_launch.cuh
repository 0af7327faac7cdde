// Device wrappers + launchers of the depthwise column walkers (per-thread bodies: dwc_core.cuh), shared by the four
// translation units k_dwc_{bwd,fwd}_{bf16,f32}.cu (one per direction and dtype so that `make -j` compiles the ~150
// template instances -- kernel size x stride x band height x activation -- in parallel).
//
// Block = CW channel groups (of CPT channels; lanes channel-contiguous) x ILB item lanes; a backward thread keeps its taps,
// weight-gradient accumulators and BatchNorm sums in registers over all the (sample, band) items it walks and
// leaves them through ONE shared-memory reduction per block (ILB-way) followed by one global atomic per
// (channel, tap) per block.
#pragma once
#include "dwc_core.cuh"
#include "td3d_kernels.h"

#include <stdlib.h>

namespace td3d {

struct SmemSink {
  float* s;          // [(KK + 2)][chw]  chw = CW * CPT channels of this block
  int c_base, chw, KK;
  __device__ __forceinline__ void dw(int c, int tap, float v) { atomicAdd(&s[tap * chw + (c - c_base)], v); }
  __device__ __forceinline__ void stat(int which, int c, float v) { atomicAdd(&s[(KK + which) * chw + (c - c_base)], v); }
};

template <typename T, int K, int S, int R, int CPT, int ACT>
__global__ void __launch_bounds__(256, 2) dwc_bwd_kernel(DwcArgs a) {
  pdl_entry();
  extern __shared__ float s_acc[];
  constexpr int KK = K * K;
  const int chunk = blockIdx.x % a.n_cchunks, ib = blockIdx.x / a.n_cchunks;
  const int ncg = a.C / CPT;                               // channel groups in the tensor
  const int cg0 = chunk * a.cw;
  const int cw_here = min(a.cw, ncg - cg0);
  const int chw = a.cw * CPT;
  for (int i = threadIdx.x; i < (KK + 2) * chw; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  const int cgl = threadIdx.x % a.cw, ilb = threadIdx.x / a.cw;
  const int il = ib * a.ilb + ilb;
  if (cgl < cw_here && ilb < a.ilb && il < a.item_lanes) {
    SmemSink sink = {s_acc, cg0 * CPT, chw, KK};
    DwcBwd<T, K, S, R, CPT, ACT>::thread_main(a, (cg0 + cgl) * CPT, il, sink);
  }
  __syncthreads();
  const int nch = cw_here * CPT;
  // One thread per channel walks its KK taps (contiguous in the reference layout [C][1][k][k]) with 16-byte vector
  // reductions where the address allows: ~600 blocks x 480 channels x 25 taps of scalar atomics were 15-25 % of the
  // late 5x5 layers.  a.dw is 16-byte aligned (parameter offsets are multiples of 4 floats).
  for (int cl = threadIdx.x; cl < nch; cl += blockDim.x) {
    const int c = cg0 * CPT + cl;
    float* dst = a.dw + (size_t)c * KK;
    const int head = (4 - (int)(((size_t)c * KK) & 3)) & 3;            // scalars up to the next 16-byte boundary
    int t = 0;
    for (; t < head && t < KK; ++t) atomicAdd(dst + t, s_acc[t * chw + cl]);
    for (; t + 4 <= KK; t += 4)
      atomicAdd(reinterpret_cast<float4*>(dst + t),
                make_float4(s_acc[t * chw + cl], s_acc[(t + 1) * chw + cl], s_acc[(t + 2) * chw + cl], s_acc[(t + 3) * chw + cl]));
    for (; t < KK; ++t) atomicAdd(dst + t, s_acc[t * chw + cl]);
  }
  if (a.stats) {
    const int slot = ib % a.slots;
    // stats rows are C floats with C % 8 == 0 and the chunk starts at a multiple of 8 channels: float4 groups stay inside the chunk
    for (int i = threadIdx.x; i < 2 * (nch >> 2); i += blockDim.x) {
      const int which = i / (nch >> 2), cl = (i - which * (nch >> 2)) << 2;
      const float* sp = &s_acc[(KK + which) * chw + cl];
      atomicAdd(reinterpret_cast<float4*>(&a.stats[((size_t)slot * 2 + which) * a.C + cg0 * CPT + cl]), make_float4(sp[0], sp[1], sp[2], sp[3]));
    }
  }
}

// TD3D_DWC_TALL=0 (tuning): 2-row bands also on tall planes (fewer registers, no spills, more halo re-reads)
inline int tall_bands() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TD3D_DWC_TALL");
    v = e ? atoi(e) : 1;
  }
  return v;
}

inline int sm_count() {
  static int g_sms = 0;
  if (!g_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

template <typename T, int K, int S, int R, int CPT, int ACT>
int dwc_launch_a(const DwBwdArgs& b, cudaStream_t st) {
  DwcArgs a;
  a.g = b.g; a.y_out = b.y_out; a.alpha = b.alpha; a.beta = b.beta; a.gamma = b.gamma;
  a.x = b.x; a.scale = b.xf.scale; a.shift = b.xf.shift; a.se = b.xf.se; a.act = b.xf.act;
  a.w_taps = b.w_taps; a.gx = b.gx; a.stats = b.stats; a.dw = b.dw;
  a.B = b.B; a.H = b.H; a.W = b.W; a.C = b.C;
  a.Ho = (b.H - 1) / S + 1; a.Wo = (b.W - 1) / S + 1;
  a.slots = b.B;
  a.n_bands = ceil_div(a.Ho, R);
  a.n_items = a.B * a.n_bands;
  const int ncg = a.C / CPT;
  // channel groups per block: all of them if they fit 256 threads, else equal chunks (multiples of 8 groups = whole sectors)
  a.n_cchunks = ceil_div(ncg, 256);
  a.cw = ceil_div(ceil_div(ncg, a.n_cchunks), 8) * 8;
  if (a.cw > ncg) a.cw = ncg;
  a.n_cchunks = ceil_div(ncg, a.cw);
  a.ilb = 256 / a.cw;
  if (a.ilb < 1) a.ilb = 1;
  const int threads = ceil_div(a.cw * a.ilb, 32) * 32;
  // item lanes: enough blocks for ~2 resident waves, but every thread should walk several items so the register-held
  // sums are flushed rarely
  int ib_max = ceil_div(a.n_items, a.ilb);
  int ib = (sm_count() * 4) / a.n_cchunks;
  if (ib < 1) ib = 1;
  if (ib > ib_max) ib = ib_max;
  a.item_lanes = ib * a.ilb;
  if (a.item_lanes > a.n_items) a.item_lanes = a.n_items;
  const size_t smem = sizeof(float) * (size_t)(K * K + 2) * a.cw * CPT;
  auto kern = dwc_bwd_kernel<T, K, S, R, CPT, ACT>;
  if (smem > 48 * 1024) TD3D_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TD3D_CUDA(launch_kernel(kern, ib * a.n_cchunks, threads, smem, st, a));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// run-time activation kind -> compile-time instance
template <typename T, int K, int S, int R, int CPT>
int dwc_launch(const DwBwdArgs& b, cudaStream_t st) {
  switch (b.xf.act) {
    case TD3D_ACT_NONE: return dwc_launch_a<T, K, S, R, CPT, TD3D_ACT_NONE>(b, st);
    case TD3D_ACT_RELU: return dwc_launch_a<T, K, S, R, CPT, TD3D_ACT_RELU>(b, st);
    case TD3D_ACT_HSWISH: return dwc_launch_a<T, K, S, R, CPT, TD3D_ACT_HSWISH>(b, st);
    case TD3D_ACT_SILU: return dwc_launch_a<T, K, S, R, CPT, TD3D_ACT_SILU>(b, st);
  }
  set_last_error("dw_bwd: unknown activation %d", b.xf.act);
  return TD3D_EINVAL;
}

template <typename T>
int dwc_dispatch(const DwBwdArgs& b, cudaStream_t st) {
  const int Ho = (b.H - 1) / b.stride + 1;
  // R output rows per band: tall planes take 4 (less halo), short ones 2 (more items to spread)
  const bool tall = Ho >= 28 && tall_bands();
  // 3x3 stride 1: 2-row bands on every plane size (measured r02 call D, layers 1 / 3: 247 vs 324 us and 237 vs 312 us
  // with 4-row bands, which spill at 128 registers); stride 2 keeps the taller band (407 vs 514 us on layer 2)
  if (b.k == 3 && b.stride == 1) return (tall && tall_bands() > 1) ? dwc_launch<T, 3, 1, 4, 2>(b, st) : dwc_launch<T, 3, 1, 2, 2>(b, st);
  if (b.k == 3 && b.stride == 2) return tall ? dwc_launch<T, 3, 2, 2, 2>(b, st) : dwc_launch<T, 3, 2, 1, 2>(b, st);
  if (b.k == 5 && b.stride == 1) return dwc_launch<T, 5, 1, 2, 1>(b, st);
  if (b.k == 5 && b.stride == 2) return dwc_launch<T, 5, 2, 1, 1>(b, st);
  set_last_error("dw_bwd: unsupported kernel/stride %d/%d", b.k, b.stride);
  return TD3D_EINVAL;
}


// ---- forward -----------------------------------------------------------------------------------------------------
struct GlobalStatSink {
  float* stats; int C;
  __device__ __forceinline__ void stat(int b, int which, int c, float v) { atomicAdd(&stats[((size_t)b * 2 + which) * C + c], v); }
};

template <typename T, int K, int S, int R, int CPT, int ACT, int OACT>
__global__ void __launch_bounds__(256, 2) dwc_fwd_kernel(DwcFwdArgs a) {
  pdl_entry();
  const int chunk = blockIdx.x % a.n_cchunks, ib = blockIdx.x / a.n_cchunks;
  const int ncg = a.C / CPT;
  const int cg0 = chunk * a.cw;
  const int cw_here = min(a.cw, ncg - cg0);
  const int cgl = threadIdx.x % a.cw, ilb = threadIdx.x / a.cw;
  const int il = ib * a.ilb + ilb;
  if (cgl < cw_here && ilb < a.ilb && il < a.item_lanes) {
    GlobalStatSink sink = {a.stats, a.C};
    DwcFwd<T, K, S, R, CPT, ACT, OACT>::thread_main(a, (cg0 + cgl) * CPT, il, sink);
  }
}

template <typename T, int K, int S, int R, int CPT, int ACT, int OACT>
int dwc_fwd_launch_a(const DwArgs& b, cudaStream_t st) {
  DwcFwdArgs a;
  a.x = b.x; a.scale = b.xf.scale; a.shift = b.xf.shift; a.se = b.xf.se; a.act = b.xf.act;
  a.w_taps = b.w_taps; a.out_bias = b.out_bias; a.out_act = b.out_act; a.y = b.y; a.stats = b.stats;
  a.B = b.B; a.H = b.H; a.W = b.W; a.C = b.C;
  a.Ho = (b.H - 1) / S + 1; a.Wo = (b.W - 1) / S + 1;
  a.n_bands = ceil_div(a.Ho, R);
  a.n_items = a.B * a.n_bands;
  const int ncg = a.C / CPT;
  a.n_cchunks = ceil_div(ncg, 256);
  a.cw = ceil_div(ceil_div(ncg, a.n_cchunks), 8) * 8;
  if (a.cw > ncg) a.cw = ncg;
  a.n_cchunks = ceil_div(ncg, a.cw);
  a.ilb = 256 / a.cw;
  if (a.ilb < 1) a.ilb = 1;
  const int threads = ceil_div(a.cw * a.ilb, 32) * 32;
  int ib_max = ceil_div(a.n_items, a.ilb);
  int ib = (sm_count() * 8) / a.n_cchunks;      // no per-thread sums to amortise beyond an item: more, shorter lanes balance better
  if (ib < 1) ib = 1;
  if (ib > ib_max) ib = ib_max;
  a.item_lanes = ib * a.ilb;
  if (a.item_lanes > a.n_items) a.item_lanes = a.n_items;
  TD3D_CUDA(launch_kernel(dwc_fwd_kernel<T, K, S, R, CPT, ACT, OACT>, ib * a.n_cchunks, threads, 0, st, a));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// (input activation, output activation): training applies the producer's activation on load and stores raw outputs;
// inference (BatchNorm folded) reads activated inputs and applies bias + activation before the store
template <typename T, int K, int S, int R, int CPT>
int dwc_fwd_launch(const DwArgs& b, cudaStream_t st) {
  if (b.out_act == TD3D_ACT_NONE) {
    switch (b.xf.act) {
      case TD3D_ACT_NONE: return dwc_fwd_launch_a<T, K, S, R, CPT, TD3D_ACT_NONE, TD3D_ACT_NONE>(b, st);
      case TD3D_ACT_RELU: return dwc_fwd_launch_a<T, K, S, R, CPT, TD3D_ACT_RELU, TD3D_ACT_NONE>(b, st);
      case TD3D_ACT_HSWISH: return dwc_fwd_launch_a<T, K, S, R, CPT, TD3D_ACT_HSWISH, TD3D_ACT_NONE>(b, st);
      case TD3D_ACT_SILU: return dwc_fwd_launch_a<T, K, S, R, CPT, TD3D_ACT_SILU, TD3D_ACT_NONE>(b, st);
    }
  } else if (b.xf.act == TD3D_ACT_NONE) {
    switch (b.out_act) {
      case TD3D_ACT_RELU: return dwc_fwd_launch_a<T, K, S, R, CPT, TD3D_ACT_NONE, TD3D_ACT_RELU>(b, st);
      case TD3D_ACT_HSWISH: return dwc_fwd_launch_a<T, K, S, R, CPT, TD3D_ACT_NONE, TD3D_ACT_HSWISH>(b, st);
      case TD3D_ACT_SILU: return dwc_fwd_launch_a<T, K, S, R, CPT, TD3D_ACT_NONE, TD3D_ACT_SILU>(b, st);
    }
  }
  set_last_error("dw_fwd: unsupported activation pair in=%d out=%d (an output activation needs an untransformed input)", b.xf.act, b.out_act);
  return TD3D_EINVAL;
}

template <typename T>
int dwc_fwd_dispatch(const DwArgs& b, cudaStream_t st) {
  const int Ho = (b.H - 1) / b.stride + 1;
  const bool tall = Ho >= 28 && tall_bands();
  if (b.k == 3 && b.stride == 1) return tall ? dwc_fwd_launch<T, 3, 1, 4, 2>(b, st) : dwc_fwd_launch<T, 3, 1, 2, 2>(b, st);
  if (b.k == 3 && b.stride == 2) return tall ? dwc_fwd_launch<T, 3, 2, 2, 2>(b, st) : dwc_fwd_launch<T, 3, 2, 1, 2>(b, st);
  if (b.k == 5 && b.stride == 1) return dwc_fwd_launch<T, 5, 1, 2, 1>(b, st);
  if (b.k == 5 && b.stride == 2) return tall ? dwc_fwd_launch<T, 5, 2, 2, 1>(b, st) : dwc_fwd_launch<T, 5, 2, 1, 1>(b, st);
  set_last_error("dw_fwd: unsupported kernel/stride %d/%d", b.k, b.stride);
  return TD3D_EINVAL;
}

}  // namespace td3d
