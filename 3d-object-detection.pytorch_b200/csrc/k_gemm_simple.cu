// CUDA-core (FFMA) GEMMs: the exact-fp32 path used for fp32 parity runs, and a cross-check for the
// tcgen05 kernels in k_gemm_tc.cu.  Same contract as the tensor-core kernels:
//   NT: Y[M,N] = A[M,K] W[N,K]^T (+bias) (+addend), epilogue statistics per column into
//       stats[slot][0][n] += sum v, stats[slot][1][n] += sum v*(ysaved ? ysaved : v)
//   TN: C[N1,N2] += A[M,N1]^T B[M,N2]   (weight gradients; split over M, fp32 atomics)
// (1x1 convolutions on NHWC are exactly these GEMMs: reference mobilenetv3.py:120,142,148,158 and
//  the classifier Linear :192.)
#include "td3d_kernels.h"

namespace td3d {

static const int GS_BM = 64, GS_BN = 64, GS_BK = 16, GS_THREADS = 256;

template <typename T>
__device__ __forceinline__ void load4(const T* p, float v[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float v[4]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float v[4]) {
  uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename T>
__global__ void __launch_bounds__(GS_THREADS)
gemm_nt_simt_kernel(const T* __restrict__ A, const T* __restrict__ Wt, T* __restrict__ Y, float* __restrict__ Yf,
                    const T* __restrict__ addend, const float* __restrict__ bias, const T* __restrict__ ysaved,
                    float* __restrict__ stats, int slots, int M, int N, int K, int act) {
  pdl_entry();
  __shared__ __align__(16) float As[GS_BK][GS_BM + 4];
  __shared__ __align__(16) float Bs[GS_BK][GS_BN + 4];
  __shared__ float s_stat[2][GS_BN];
  const int m0 = blockIdx.x * GS_BM, n0 = blockIdx.y * GS_BN;
  const int tid = threadIdx.x;
  const int lrow = tid >> 2, lk = (tid & 3) << 2;     // loader: 64 rows x 4 k-quads
  const int ty = tid >> 4, tx = tid & 15;
  if (tid < 2 * GS_BN) s_stat[tid / GS_BN][tid % GS_BN] = 0.f;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GS_BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (m0 + lrow < M && k0 + lk < K) load4<T>(A + (size_t)(m0 + lrow) * K + k0 + lk, av);
    if (n0 + lrow < N && k0 + lk < K) load4<T>(Wt + (size_t)(n0 + lrow) * K + k0 + lk, bv);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      As[lk + i][lrow] = av[i];
      Bs[lk + i][lrow] = bv[i];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GS_BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  // ---- epilogue ----
  float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[n];
      v = act_fwd(v, act);
      if (addend) v += to_f(addend[(size_t)m * N + n]);
      float r;
      if (Yf) { Yf[(size_t)m * N + n] = v; r = v; }
      else { T o = from_f<T>(v); Y[(size_t)m * N + n] = o; r = to_f(o); }
      st1[j] += r;
      st2[j] = fmaf(r, ysaved ? to_f(ysaved[(size_t)m * N + n]) : r, st2[j]);
    }
  }
  if (stats) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      atomicAdd(&s_stat[0][tx * 4 + j], st1[j]);
      atomicAdd(&s_stat[1][tx * 4 + j], st2[j]);
    }
    __syncthreads();
    const int slot = blockIdx.x % slots;
    if (tid < 2 * GS_BN) {
      int which = tid / GS_BN, n = n0 + tid % GS_BN;
      if (n < N) atomicAdd(&stats[((size_t)slot * 2 + which) * N + n], s_stat[which][tid % GS_BN]);
    }
  }
}

int launch_gemm_nt_simt(const GemmNT& g, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(g.K % 4 == 0 && g.M > 0 && g.N > 0, "gemm_nt: bad shape M=%d N=%d K=%d", g.M, g.N, g.K);
  TD3D_REQUIRE(!g.stats || g.slots > 0, "gemm_nt: stats need slots > 0");
  dim3 grid(ceil_div(g.M, GS_BM), ceil_div(g.N, GS_BN));
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(gemm_nt_simt_kernel<bf16>, grid, GS_THREADS, 0, st, 
        (const bf16*)g.a, (const bf16*)g.w, g.out_f32 ? nullptr : (bf16*)g.y, g.out_f32 ? (float*)g.y : nullptr,
        (const bf16*)g.addend, g.bias, (const bf16*)g.ysaved, g.stats, g.slots, g.M, g.N, g.K, g.act));
  else
    TD3D_CUDA(launch_kernel(gemm_nt_simt_kernel<float>, grid, GS_THREADS, 0, st, 
        (const float*)g.a, (const float*)g.w, g.out_f32 ? nullptr : (float*)g.y, g.out_f32 ? (float*)g.y : nullptr,
        (const float*)g.addend, g.bias, (const float*)g.ysaved, g.stats, g.slots, g.M, g.N, g.K, g.act));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

// C[N1,N2] += sum_{m in chunk} A[m,n1] * B[m,n2]
template <typename T>
__global__ void __launch_bounds__(GS_THREADS)
gemm_tn_simt_kernel(const T* __restrict__ A, const T* __restrict__ Bm, float* __restrict__ C, int M, int N1, int N2,
                    int m_per_part) {
  pdl_entry();
  __shared__ __align__(16) float As[GS_BK][GS_BM + 4];
  __shared__ __align__(16) float Bs[GS_BK][GS_BN + 4];
  const int n10 = blockIdx.x * GS_BM, n20 = blockIdx.y * GS_BN;
  const int ms = blockIdx.z * m_per_part, me = min(M, ms + m_per_part);
  const int tid = threadIdx.x;
  const int lm = tid >> 4, lc = (tid & 15) << 2;       // loader: 16 m-rows x 16 column quads
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int mk = ms; mk < me; mk += GS_BK) {
    float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
    const int m = mk + lm;
    if (m < me) {
      if (n10 + lc < N1) load4<T>(A + (size_t)m * N1 + n10 + lc, av);
      if (n20 + lc < N2) load4<T>(Bm + (size_t)m * N2 + n20 + lc, bv);
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&As[lm][lc]) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4*>(&Bs[lm][lc]) = make_float4(bv[0], bv[1], bv[2], bv[3]);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GS_BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n1 = n10 + ty * 4 + i;
    if (n1 >= N1) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n2 = n20 + tx * 4 + j;
      if (n2 < N2) atomicAdd(&C[(size_t)n1 * N2 + n2], acc[i][j]);
    }
  }
}

int launch_gemm_tn_simt(const GemmTN& g, int dtype, cudaStream_t st) {
  TD3D_REQUIRE(g.N1 % 4 == 0 && g.N2 % 4 == 0 && g.M > 0, "gemm_tn: bad shape M=%d N1=%d N2=%d", g.M, g.N1, g.N2);
  int tiles = ceil_div(g.N1, GS_BM) * ceil_div(g.N2, GS_BN);
  int parts = ceil_div(148 * 4, tiles);
  int max_parts = ceil_div(g.M, GS_BK * 8);
  if (parts > max_parts) parts = max_parts;
  if (parts < 1) parts = 1;
  int mpp = ceil_div(ceil_div(g.M, parts), GS_BK) * GS_BK;
  parts = ceil_div(g.M, mpp);
  dim3 grid(ceil_div(g.N1, GS_BM), ceil_div(g.N2, GS_BN), parts);
  if (dtype == TD3D_BF16)
    TD3D_CUDA(launch_kernel(gemm_tn_simt_kernel<bf16>, grid, GS_THREADS, 0, st, (const bf16*)g.a, (const bf16*)g.b, g.c, g.M, g.N1, g.N2, mpp));
  else
    TD3D_CUDA(launch_kernel(gemm_tn_simt_kernel<float>, grid, GS_THREADS, 0, st, (const float*)g.a, (const float*)g.b, g.c, g.M, g.N1, g.N2, mpp));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
