// REFERENCE IMPLEMENTATION (direct, untiled) kept for A/B debugging: TD3D_DW_SIMPLE=1 selects it.
// The production kernels are in k_dwconv.cu.
//
// Depthwise k x k convolution (k in {3,5}, stride in {1,2}, pad (k-1)/2), NHWC, 8-channel vectors
// (reference: nn.Conv2d(groups=hidden_dim) in InvertedResidual, torchdet3d/models/mobilenetv3.py:136,152).
//
//   forward   y = dw(x_t),  x_t = act(se*(scale*x+shift)) applied on load (zero padding AFTER the
//             transform, as the reference pads the activated tensor); epilogue accumulates the
//             BatchNorm statistics and the SE squeeze sums per (sample, channel)
//   bwd-data  gx = dw^T(gy) * act'(u(x)),  gy = alpha*g + beta*y + gamma applied on load (lazy BN
//             backward); epilogue accumulates sum(gx), sum(gx*x) for the producer's BN backward
//   bwd-wgt   dW[c,ky,kx] = sum gy * x_t(shifted); persistent blocks, one flush of atomics
//
// Mapping: a block works inside one sample on a chunk of pixels; a thread owns a fixed 8-channel
// vector and walks pixels ("row loop", see k_elementwise.cu), so per-channel constants and the
// filter taps of its channels stay in registers/L1.
#include "td3d_kernels.h"

namespace td3d {

static const int DW_THREADS = 256;
static const int DW_ITERS = 4;

template <typename T, int K, int S>
__global__ void __launch_bounds__(DW_THREADS)
dw_fwd_kernel(const T* __restrict__ x, XForm xf, const float* __restrict__ w, T* __restrict__ y,
              float* __restrict__ stats, int H, int W, int Ho, int Wo, int C, int pix_per_block) {
  extern __shared__ float s_acc[];   // [2][C]
  constexpr int P = (K - 1) / 2;
  const int b = blockIdx.y;
  if (stats) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
  }
  const int CV = C >> 3;
  for (int cv0 = 0; cv0 < CV; cv0 += blockDim.x) {
    int CVb = min((int)blockDim.x, CV - cv0);
    int PL = max(1, (int)blockDim.x / CVb);
    int cv = threadIdx.x % CVb, pl = threadIdx.x / CVb;
    if (pl >= PL) continue;
    const int c = (cv0 + cv) << 3;
    float sc[8], sh[8], se[8], a1[8], a2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; se[i] = 1.f; a1[i] = 0.f; a2[i] = 0.f; }
    if (xf.scale) { loadf8(xf.scale + c, sc); loadf8(xf.shift + c, sh); }
    if (xf.se) loadf8(xf.se + (size_t)b * C + c, se);
    const int p0 = blockIdx.x * pix_per_block, p1 = min(Ho * Wo, p0 + pix_per_block);
    for (int p = p0 + pl; p < p1; p += PL) {
      const int oh = p / Wo, ow = p % Wo;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int ih = oh * S + ky - P;
        if (ih < 0 || ih >= H) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int iw = ow * S + kx - P;
          if (iw < 0 || iw >= W) continue;
          float v[8], wt[8];
          load8(x + (((size_t)b * H + ih) * W + iw) * C + c, v);
          loadf8(w + (size_t)(ky * K + kx) * C + c, wt);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float t = act_fwd(se[i] * fmaf(v[i], sc[i], sh[i]), xf.act);
            acc[i] = fmaf(t, wt[i], acc[i]);
          }
        }
      }
      store8(y + (((size_t)b * Ho + oh) * Wo + ow) * C + c, acc);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float r = to_f(from_f<T>(acc[i]));     // statistics of the stored (rounded) tensor
        a1[i] += r;
        a2[i] = fmaf(r, r, a2[i]);
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
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[(size_t)b * 2 * C + i], s_acc[i]);
  }
}

// gx[b,h,w,c] = act'(u(x)) * sum_{ky,kx : (h+P-ky)%S==0, (w+P-kx)%S==0} gy[b,(h+P-ky)/S,(w+P-kx)/S,c] * w[ky,kx,c]
template <typename T, int K, int S>
__global__ void __launch_bounds__(DW_THREADS)
dw_bwd_data_kernel(const T* __restrict__ g, const T* __restrict__ yo, const float* __restrict__ alpha,
                   const float* __restrict__ beta, const float* __restrict__ gamma, const T* __restrict__ x,
                   XForm xf, const float* __restrict__ w, T* __restrict__ gx, float* __restrict__ stats,
                   int H, int W, int Ho, int Wo, int C, int pix_per_block) {
  extern __shared__ float s_acc[];   // [2][C]
  constexpr int P = (K - 1) / 2;
  const int b = blockIdx.y;
  if (stats) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_acc[i] = 0.f;
    __syncthreads();
  }
  const int CV = C >> 3;
  for (int cv0 = 0; cv0 < CV; cv0 += blockDim.x) {
    int CVb = min((int)blockDim.x, CV - cv0);
    int PL = max(1, (int)blockDim.x / CVb);
    int cv = threadIdx.x % CVb, pl = threadIdx.x / CVb;
    if (pl >= PL) continue;
    const int c = (cv0 + cv) << 3;
    float sc[8], sh[8], se[8], al[8], be[8], ga[8], a1[8], a2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; se[i] = 1.f; a1[i] = 0.f; a2[i] = 0.f; }
    if (xf.scale) { loadf8(xf.scale + c, sc); loadf8(xf.shift + c, sh); }
    if (xf.se) loadf8(xf.se + (size_t)b * C + c, se);
    loadf8(alpha + (size_t)b * C + c, al);
    loadf8(beta + c, be);
    loadf8(gamma + (size_t)b * C + c, ga);
    const int p0 = blockIdx.x * pix_per_block, p1 = min(H * W, p0 + pix_per_block);
    for (int p = p0 + pl; p < p1; p += PL) {
      const int h = p / W, wi = p % W;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int th = h + P - ky;
        if (th < 0 || (th % S) != 0) continue;
        const int oh = th / S;
        if (oh >= Ho) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
          const int tw = wi + P - kx;
          if (tw < 0 || (tw % S) != 0) continue;
          const int ow = tw / S;
          if (ow >= Wo) continue;
          const size_t off = (((size_t)b * Ho + oh) * Wo + ow) * C + c;
          float gv[8], yv[8], wt[8];
          load8(g + off, gv);
          load8(yo + off, yv);
          loadf8(w + (size_t)(ky * K + kx) * C + c, wt);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float gy = fmaf(al[i], gv[i], fmaf(be[i], yv[i], ga[i]));
            acc[i] = fmaf(gy, wt[i], acc[i]);
          }
        }
      }
      const size_t xoff = (((size_t)b * H + h) * W + wi) * C + c;
      float xv[8];
      load8(x + xoff, xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float u = se[i] * fmaf(xv[i], sc[i], sh[i]);
        acc[i] *= act_bwd(u, xf.act);
      }
      store8(gx + xoff, acc);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float r = to_f(from_f<T>(acc[i]));
        a1[i] += r;
        a2[i] = fmaf(r, xv[i], a2[i]);
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
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&stats[(size_t)b * 2 * C + i], s_acc[i]);
  }
}

// Persistent: thread = (channel vector, tap); block = CVB vectors x K*K taps; grid.y partitions
// the B*Ho output rows. dW layout = reference [C,1,K,K] -> dw[c*K*K + tap].
template <typename T, int K, int S>
__global__ void __launch_bounds__(DW_THREADS)
dw_bwd_weight_kernel(const T* __restrict__ g, const T* __restrict__ yo, const float* __restrict__ alpha,
                     const float* __restrict__ beta, const float* __restrict__ gamma, const T* __restrict__ x,
                     XForm xf, float* __restrict__ dw, int B, int H, int W, int Ho, int Wo, int C, int CVB,
                     int rows_per_part) {
  constexpr int P = (K - 1) / 2;
  constexpr int KK = K * K;
  const int cvl = threadIdx.x / KK, tap = threadIdx.x % KK;
  const int cvi = blockIdx.x * CVB + cvl;
  if (cvl >= CVB || cvi >= (C >> 3)) return;
  const int c = cvi << 3;
  const int ky = tap / K, kx = tap % K;
  float sc[8], sh[8], be[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; acc[i] = 0.f; }
  if (xf.scale) { loadf8(xf.scale + c, sc); loadf8(xf.shift + c, sh); }
  loadf8(beta + c, be);
  const int r0 = blockIdx.y * rows_per_part, r1 = min(B * Ho, r0 + rows_per_part);
  int cur_b = -1;
  float se[8], al[8], ga[8];
  for (int r = r0; r < r1; ++r) {
    const int b = r / Ho, oh = r % Ho;
    if (b != cur_b) {
      cur_b = b;
#pragma unroll
      for (int i = 0; i < 8; ++i) se[i] = 1.f;
      if (xf.se) loadf8(xf.se + (size_t)b * C + c, se);
      loadf8(alpha + (size_t)b * C + c, al);
      loadf8(gamma + (size_t)b * C + c, ga);
    }
    const int ih = oh * S + ky - P;
    if (ih < 0 || ih >= H) continue;
    for (int ow = 0; ow < Wo; ++ow) {
      const int iw = ow * S + kx - P;
      if (iw < 0 || iw >= W) continue;
      const size_t off = (((size_t)b * Ho + oh) * Wo + ow) * C + c;
      float gv[8], yv[8], xv[8];
      load8(g + off, gv);
      load8(yo + off, yv);
      load8(x + (((size_t)b * H + ih) * W + iw) * C + c, xv);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float gy = fmaf(al[i], gv[i], fmaf(be[i], yv[i], ga[i]));
        float t = act_fwd(se[i] * fmaf(xv[i], sc[i], sh[i]), xf.act);
        acc[i] = fmaf(gy, t, acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(&dw[(size_t)(c + i) * KK + tap], acc[i]);
}

static void dw_grid(int B, int npix, int C, dim3* grid, int* ppb) {
  int CV = C >> 3;
  if (CV > DW_THREADS) CV = DW_THREADS;
  int PL = DW_THREADS / CV;
  if (PL < 1) PL = 1;
  int p = PL * DW_ITERS;
  if (p > npix) p = npix;
  *ppb = p;
  *grid = dim3((unsigned)ceil_div(npix, p), (unsigned)B);
}

template <typename T, int K, int S>
static int dw_fwd_t(const DwArgs& a, cudaStream_t st) {
  int Ho = (a.H - 1) / S + 1, Wo = (a.W - 1) / S + 1;
  dim3 grid; int ppb;
  dw_grid(a.B, Ho * Wo, a.C, &grid, &ppb);
  size_t smem = a.stats ? sizeof(float) * 2 * a.C : 0;
  dw_fwd_kernel<T, K, S><<<grid, DW_THREADS, smem, st>>>((const T*)a.x, a.xf, a.w_taps, (T*)a.y, a.stats, a.H, a.W,
                                                         Ho, Wo, a.C, ppb);
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

template <typename T, int K, int S>
static int dw_bwd_t(const DwBwdArgs& a, cudaStream_t st) {
  int Ho = (a.H - 1) / S + 1, Wo = (a.W - 1) / S + 1;
  if (a.gx) {
    dim3 grid; int ppb;
    dw_grid(a.B, a.H * a.W, a.C, &grid, &ppb);
    size_t smem = a.stats ? sizeof(float) * 2 * a.C : 0;
    dw_bwd_data_kernel<T, K, S><<<grid, DW_THREADS, smem, st>>>(
        (const T*)a.g, (const T*)a.y_out, a.alpha, a.beta, a.gamma, (const T*)a.x, a.xf, a.w_taps, (T*)a.gx, a.stats,
        a.H, a.W, Ho, Wo, a.C, ppb);
    TD3D_LAUNCH_CHECK();
  }
  if (a.dw) {
    const int KK = K * K;
    int CVB = DW_THREADS / KK;
    int CV = a.C >> 3;
    if (CVB > CV) CVB = CV;
    int gx = ceil_div(CV, CVB);
    int rows = a.B * Ho;
    int parts = ceil_div(148 * 8, gx);
    if (parts > rows) parts = rows;
    int rpp = ceil_div(rows, parts);
    parts = ceil_div(rows, rpp);
    dim3 grid(gx, parts);
    dw_bwd_weight_kernel<T, K, S><<<grid, DW_THREADS, 0, st>>>(
        (const T*)a.g, (const T*)a.y_out, a.alpha, a.beta, a.gamma, (const T*)a.x, a.xf, a.dw, a.B, a.H, a.W, Ho, Wo,
        a.C, CVB, rpp);
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

int launch_dw_fwd_simple(const DwArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.C <= 4096, "dw fwd: C=%d must be a multiple of 8", a.C);
  DW_DISPATCH(dw_fwd_t, a);
  set_last_error("dw fwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  return TD3D_EINVAL;
}

int launch_dw_bwd_simple(const DwBwdArgs& a, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(a.C % 8 == 0 && a.C <= 4096, "dw bwd: C=%d must be a multiple of 8", a.C);
  DW_DISPATCH(dw_bwd_t, a);
  set_last_error("dw bwd: unsupported kernel=%d stride=%d", a.k, a.stride);
  return TD3D_EINVAL;
}

}  // namespace td3d
