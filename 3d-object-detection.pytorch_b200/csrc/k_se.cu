// Squeeze-and-Excite FCs (reference SELayer, torchdet3d/models/mobilenetv3.py:92-107):
//   zbar = mean_HW(z);  hid = relu(W1 zbar + b1);  gate = h_sigmoid(W2 hid + b2);  x = z * gate
// The squeeze (per-(b,c) pixel sums) is produced by the depthwise-conv epilogue and the gate is
// applied by the consumer's load transform, so only the tiny [B,C]x[C,C/4] products remain.  They
// are latency-bound (<= 0.25 GFLOP per layer), so the design goal is few, wide launches:
//   forward  = 2 launches: (squeeze-finalize + FC1 + ReLU), (FC2 + h_sigmoid)
//   backward = 3 launches: (gate' + FC2^T + ReLU mask), (FC1^T), (both weight/bias gradients)
// All arithmetic is fp32 FFMA (bit-comparable to the fp32 oracle up to summation order).
//
// se_fc_kernel: Y[b,n] = epi(sum_k A[b,k] W[n,k] + bias[n]).  A CTA stages SE_SB sample rows of A in
// shared memory (computing them on the fly from the statistic slots where A is not materialised
// yet), a warp owns 4 output columns: lanes stride over k with 128-bit loads of the W rows, and
// the 4 x 8 partial sums of a lane are reduced with one 31-shuffle transpose-sum.
#include "td3d_kernels.h"

namespace td3d {

static const int SE_SB = 8;          // samples per CTA
static const int SE_NB = 32;         // output columns per CTA (8 warps x 4)
static const int SE_THREADS = 256;

enum { SE_A_PLAIN = 0, SE_A_ZBAR = 1, SE_A_GPRE = 2 };
enum { SE_E_RELU = 0, SE_E_GATE = 1, SE_E_MASK = 2, SE_E_NONE = 3 };

struct SeFc {
  const float* a;            // PLAIN: A [B,K]
  const float* stats;        // ZBAR / GPRE: [B][2][K] statistic slots
  const float* scale; const float* shift;   // [K] or null
  const float* pre;          // GPRE: [B,K] pre-activation of the gate
  float inv_hw;
  float* a_out;              // ZBAR / GPRE: materialised A [B,K] (written by the blockIdx.y == 0 CTAs)
  const float* w;            // [N,K]
  const float* bias;         // [N] or null
  const float* mask;         // MASK: [B,N]; y = 0 where mask <= 0
  float* y;                  // [B,N]
  float* y2;                 // GATE: h_sigmoid(y)
  int B, N, K;
};

// lane l ends with the sum over the warp of v[l] (31 shuffles instead of 32*5)
__device__ __forceinline__ float se_transpose_sum32(float v[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16, n = 16; o >= 1; o >>= 1, n >>= 1) {
    const bool hi = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      float keep = hi ? v[i + n] : v[i];
      float send = hi ? v[i] : v[i + n];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

template <int AMODE, int EMODE>
__global__ void __launch_bounds__(SE_THREADS) se_fc_kernel(SeFc p) {
  pdl_entry();
  extern __shared__ __align__(16) float se_sa[];     // [SE_SB][K]
  const int b0 = blockIdx.x * SE_SB, n0 = blockIdx.y * SE_NB;
  const int K = p.K;
  // ---- stage A (zero rows past the batch): a thread owns column k of all SE_SB rows, so the loads of the 8
  // samples are in flight together and the per-column constants are read once ----
  for (int k = threadIdx.x; k < K; k += SE_THREADS) {
    float v[SE_SB];
    float sck = 1.f, shk = 0.f;
    if (AMODE != SE_A_PLAIN && p.scale) { sck = __ldg(p.scale + k); shk = __ldg(p.shift + k); }
    if (AMODE == SE_A_PLAIN) {
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) v[s] = b0 + s < p.B ? __ldg(p.a + (size_t)(b0 + s) * K + k) : 0.f;
    } else if (AMODE == SE_A_ZBAR) {
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) v[s] = b0 + s < p.B ? __ldg(p.stats + ((size_t)(b0 + s) * 2 + 0) * K + k) : 0.f;
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) {
        v[s] *= p.inv_hw;
        if (p.scale) v[s] = fmaf(v[s], sck, shk);
        if (b0 + s >= p.B) v[s] = 0.f;
      }
    } else {
      float p1[SE_SB], p2[SE_SB], pr[SE_SB];
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) {
        const bool on = b0 + s < p.B;
        p1[s] = on ? __ldg(p.stats + ((size_t)(b0 + s) * 2 + 0) * K + k) : 0.f;
        p2[s] = on ? __ldg(p.stats + ((size_t)(b0 + s) * 2 + 1) * K + k) : 0.f;
        pr[s] = on ? __ldg(p.pre + (size_t)(b0 + s) * K + k) : 0.f;
      }
#pragma unroll
      for (int s = 0; s < SE_SB; ++s) {
        const float gs = p.scale ? fmaf(sck, p2[s], shk * p1[s]) : p2[s];
        v[s] = b0 + s < p.B ? gs * hsigmoid_bwd(pr[s]) : 0.f;
      }
    }
#pragma unroll
    for (int s = 0; s < SE_SB; ++s) {
      se_sa[s * K + k] = v[s];
      if (AMODE != SE_A_PLAIN && blockIdx.y == 0 && b0 + s < p.B) p.a_out[(size_t)(b0 + s) * K + k] = v[s];
    }
  }
  __syncthreads();
  // ---- 4 columns per warp ----
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nw = n0 + warp * 4;
  float acc[32];                                     // [j][s] -> j*8 + s
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const float* wr[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wr[j] = p.w + (size_t)min(nw + j, p.N - 1) * K;   // clamped rows are discarded below
#pragma unroll 4
  for (int k = lane * 4; k < K; k += 128) {
    float4 wv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) wv[j] = __ldg(reinterpret_cast<const float4*>(wr[j] + k));
#pragma unroll
    for (int s = 0; s < SE_SB; ++s) {
      const float4 av = *reinterpret_cast<const float4*>(se_sa + s * K + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float t = acc[j * 8 + s];
        t = fmaf(av.x, wv[j].x, t);
        t = fmaf(av.y, wv[j].y, t);
        t = fmaf(av.z, wv[j].z, t);
        t = fmaf(av.w, wv[j].w, t);
        acc[j * 8 + s] = t;
      }
    }
  }
  float v = se_transpose_sum32(acc);
  const int j = lane >> 3, s = lane & 7;
  const int n = nw + j, b = b0 + s;
  if (n < p.N && b < p.B) {
    if (p.bias) v += p.bias[n];
    const size_t o = (size_t)b * p.N + n;
    if (EMODE == SE_E_RELU) v = fmaxf(v, 0.f);
    if (EMODE == SE_E_MASK) v = p.mask[o] > 0.f ? v : 0.f;
    p.y[o] = v;
    if (EMODE == SE_E_GATE) p.y2[o] = hsigmoid(v);
  }
}

template <int AMODE, int EMODE>
static int se_fc_launch(const SeFc& p, cudaStream_t st) {
  TD3D_REQUIRE(p.K % 4 == 0 && p.K > 0 && p.N > 0 && p.B > 0, "se fc: bad shape B=%d N=%d K=%d", p.B, p.N, p.K);
  const size_t smem = sizeof(float) * SE_SB * (size_t)p.K;
  TD3D_REQUIRE(smem <= 160 * 1024, "se fc: K=%d too large for the staged rows", p.K);
  static size_t configured = 48 * 1024;
  if (smem > configured) {
    TD3D_CUDA(cudaFuncSetAttribute(se_fc_kernel<AMODE, EMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = 160 * 1024;
  }
  dim3 grid(ceil_div(p.B, SE_SB), ceil_div(p.N, SE_NB));
  TD3D_CUDA(launch_kernel(se_fc_kernel<AMODE, EMODE>, grid, SE_THREADS, smem, st, p));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_se_fwd(const SeArgs& a, cudaStream_t st) {
  SeFc f1 = {};
  f1.stats = a.pool_stats; f1.scale = a.scale; f1.shift = a.shift; f1.inv_hw = a.inv_hw; f1.a_out = a.zbar;
  f1.w = a.w1; f1.bias = a.b1; f1.y = a.hid; f1.B = a.B; f1.N = a.Ch; f1.K = a.C;
  TD3D_TRY((se_fc_launch<SE_A_ZBAR, SE_E_RELU>(f1, st)));
  SeFc f2 = {};
  f2.a = a.hid; f2.w = a.w2; f2.bias = a.b2; f2.y = a.pre; f2.y2 = a.gate; f2.B = a.B; f2.N = a.C; f2.K = a.Ch;
  TD3D_TRY((se_fc_launch<SE_A_PLAIN, SE_E_GATE>(f2, st)));
  return TD3D_OK;
}

// ---- weight / bias gradients: C[n1,n2] += sum_b A[b,n1] Bm[b,n2];  dbias[n1] += sum_b A[b,n1] ----
// Two independent problems share one launch (dW2 = g_pre^T hid, dW1 = g_hid^T zbar).  A CTA owns a
// 32 x 32 output tile over the whole batch (no atomics), staging 64 samples at a time.
struct SeTn { const float* a; const float* b; float* c; float* dbias; int N1, N2, tiles2, n_ctas; };
struct SeTnPair { SeTn p[2]; int B; };

static const int SE_TM = 64;

__global__ void __launch_bounds__(SE_THREADS) se_wgrad_kernel(SeTnPair q) {
  pdl_entry();
  __shared__ __align__(16) float sa[SE_TM][32];
  __shared__ __align__(16) float sb[SE_TM][32];
  const int which = (int)blockIdx.x < q.p[0].n_ctas ? 0 : 1;
  const SeTn& p = q.p[which];
  const int cta = (int)blockIdx.x - (which ? q.p[0].n_ctas : 0);
  const int n10 = (cta / p.tiles2) * 32, n20 = (cta % p.tiles2) * 32;
  const int ty = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lr = threadIdx.x >> 3, lc = (threadIdx.x & 7) * 4;      // loader: 32 rows x 8 column quads, 2 passes
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float colsum = 0.f;
  for (int m0 = 0; m0 < q.B; m0 += SE_TM) {
    float4 av[2], bv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = m0 + lr + 32 * h;
      av[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      bv[h] = av[h];
      if (m < q.B) {
        if (n10 + lc < p.N1) av[h] = __ldg(reinterpret_cast<const float4*>(p.a + (size_t)m * p.N1 + n10 + lc));
        if (n20 + lc < p.N2) bv[h] = __ldg(reinterpret_cast<const float4*>(p.b + (size_t)m * p.N2 + n20 + lc));
      }
    }
    __syncthreads();                                   // previous chunk consumed
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(&sa[lr + 32 * h][lc]) = av[h];
      *reinterpret_cast<float4*>(&sb[lr + 32 * h][lc]) = bv[h];
    }
    __syncthreads();
#pragma unroll 8
    for (int m = 0; m < SE_TM; ++m) {
      const float4 a4 = *reinterpret_cast<const float4*>(&sa[m][ty * 4]);
      const float bvv = sb[m][lane];
      acc[0] = fmaf(a4.x, bvv, acc[0]);
      acc[1] = fmaf(a4.y, bvv, acc[1]);
      acc[2] = fmaf(a4.z, bvv, acc[2]);
      acc[3] = fmaf(a4.w, bvv, acc[3]);
    }
    if (n20 == 0 && ty == 0) {
#pragma unroll 8
      for (int m = 0; m < SE_TM; ++m) colsum += sa[m][lane];
    }
  }
  const int n2 = n20 + lane;
  if (n2 < p.N2) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n1 = n10 + ty * 4 + i;
      if (n1 < p.N1) p.c[(size_t)n1 * p.N2 + n2] += acc[i];
    }
  }
  if (n20 == 0 && ty == 0 && n10 + lane < p.N1) p.dbias[n10 + lane] += colsum;
}

int launch_se_bwd(const SeBwdArgs& a, cudaStream_t st) {
  // g_pre = gs * h_sigmoid'(pre);  g_hid = (g_pre W2) * [hid > 0]      (W operand = W2^T [Ch, C])
  SeFc f1 = {};
  f1.stats = a.bwd_stats; f1.scale = a.scale; f1.shift = a.shift; f1.pre = a.pre; f1.a_out = a.g_pre;
  f1.w = a.w2t; f1.mask = a.hid; f1.y = a.g_hid; f1.B = a.B; f1.N = a.Ch; f1.K = a.C;
  TD3D_TRY((se_fc_launch<SE_A_GPRE, SE_E_MASK>(f1, st)));
  // g_zbar = g_hid W1                                                   (W operand = W1^T [C, Ch])
  SeFc f2 = {};
  f2.a = a.g_hid; f2.w = a.w1t; f2.y = a.g_pool; f2.B = a.B; f2.N = a.C; f2.K = a.Ch;
  TD3D_TRY((se_fc_launch<SE_A_PLAIN, SE_E_NONE>(f2, st)));
  // weight / bias gradients (the gradient arena is zeroed at the start of backward; += accumulates)
  TD3D_REQUIRE(a.C % 4 == 0 && a.Ch % 4 == 0, "se bwd: C=%d Ch=%d must be multiples of 4", a.C, a.Ch);
  SeTnPair q;
  q.B = a.B;
  q.p[0].a = a.g_pre; q.p[0].b = a.hid; q.p[0].c = a.dw2; q.p[0].dbias = a.db2; q.p[0].N1 = a.C; q.p[0].N2 = a.Ch;
  q.p[1].a = a.g_hid; q.p[1].b = a.zbar; q.p[1].c = a.dw1; q.p[1].dbias = a.db1; q.p[1].N1 = a.Ch; q.p[1].N2 = a.C;
  for (int i = 0; i < 2; ++i) {
    q.p[i].tiles2 = ceil_div(q.p[i].N2, 32);
    q.p[i].n_ctas = ceil_div(q.p[i].N1, 32) * q.p[i].tiles2;
  }
  TD3D_CUDA(launch_kernel(se_wgrad_kernel, q.p[0].n_ctas + q.p[1].n_ctas, SE_THREADS, 0, st, q));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
