// Fused multi-tensor optimizer over the flat parameter arena + weight packing.
// Reference: torchdet3d/builders/optim_builder.py:5-19 -> torch.optim.SGD(momentum, nesterov) /
// AdamW (name 'adam') / RMSprop / Adadelta over ALL parameters with one global weight decay.
// The reference launches a few kernels per tensor (~150-400 tensors); here one launch updates
// the whole arena. Regressor heads whose class is absent from the batch have grad=None in the
// reference (model_builder.py:137) and are skipped entirely: `present` + per-head step counters.
#include "td3d_kernels.h"

namespace td3d {

__global__ void optim_prepare_kernel(int32_t* steps, const int32_t* present, int n_heads) {
  pdl_entry();
  int k = threadIdx.x;
  if (k > n_heads) return;
  if (k == 0 || !present || present[k - 1]) steps[k] += 1;
}

__global__ void __launch_bounds__(256) optim_kernel(OptimArgs a) {
  pdl_entry();
  __shared__ float s_bc1[33], s_bc2s[33];
  __shared__ int s_step[33];
  if ((int)threadIdx.x <= a.n_heads) {
    int st = a.steps[threadIdx.x];
    s_step[threadIdx.x] = st;
    if (a.d.kind == TD3D_OPT_ADAMW) {
      s_bc1[threadIdx.x] = (float)(1.0 - pow((double)a.d.beta1, (double)st));
      s_bc2s[threadIdx.x] = (float)sqrt(1.0 - pow((double)a.d.beta2, (double)st));
    }
  }
  __syncthreads();
  const float lr = a.d.lr, wd = a.d.weight_decay;
  const int64_t head_end = a.head_off + a.head_stride * a.n_heads;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    int seg = 0;
    if (i >= a.skip_begin && i < a.skip_end) continue;
    if (i >= a.head_off && i < head_end) {
      int k = (int)((i - a.head_off) / a.head_stride);
      if (a.present && !a.present[k]) continue;        // grad=None: tensor untouched
      seg = k + 1;
    }
    float p = a.p[i];
    float g = a.g[i] * a.d.grad_scale;
    if (a.d.kind == TD3D_OPT_ADAMW) {
      p *= 1.f - lr * wd;                               // decoupled decay
      float m = a.s0[i], v = a.s1[i];
      m = m + (g - m) * (1.f - a.d.beta1);              // lerp, as torch
      v = v * a.d.beta2 + (1.f - a.d.beta2) * g * g;
      float denom = sqrtf(v) / s_bc2s[seg] + a.d.eps;
      p -= (lr / s_bc1[seg]) * (m / denom);
      a.s0[i] = m; a.s1[i] = v;
    } else if (a.d.kind == TD3D_OPT_SGD) {
      g = fmaf(wd, p, g);
      if (a.d.momentum != 0.f) {
        float buf = s_step[seg] <= 1 ? g : fmaf(a.s0[i], a.d.momentum, g);
        a.s0[i] = buf;
        g = a.d.nesterov ? fmaf(a.d.momentum, buf, g) : buf;
      }
      p -= lr * g;
    } else if (a.d.kind == TD3D_OPT_RMSPROP) {
      g = fmaf(wd, p, g);
      float sq = a.s0[i] * a.d.alpha + (1.f - a.d.alpha) * g * g;
      a.s0[i] = sq;
      p -= lr * g / (sqrtf(sq) + a.d.eps);
    } else {                                            // Adadelta (eps = 1e-6 default)
      g = fmaf(wd, p, g);
      float sq = a.s0[i] * a.d.rho + (1.f - a.d.rho) * g * g;
      float ad = a.s1[i];
      float delta = sqrtf(ad + a.d.eps) / sqrtf(sq + a.d.eps) * g;
      a.s0[i] = sq;
      a.s1[i] = ad * a.d.rho + (1.f - a.d.rho) * delta * delta;
      p -= lr * delta;
    }
    a.p[i] = p;
  }
}

int launch_optim(const OptimArgs& a, cudaStream_t st) {
  TD3D_REQUIRE(a.n_heads <= 32, "optim: too many heads");
  TD3D_CUDA(launch_kernel(optim_prepare_kernel, 1, 64, 0, st, a.steps, a.present, a.n_heads));
  TD3D_LAUNCH_CHECK();
  int blocks = (int)((a.n + 256 * 4 - 1) / (256 * 4));
  if (blocks > 148 * 16) blocks = 148 * 16;
  TD3D_CUDA(launch_kernel(optim_kernel, blocks, 256, 0, st, a));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t n) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = from_f<T>(src[i]);
}
template <typename T>
__global__ void cast_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t n) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = to_f(src[i]);
}
// dst[c*rows + r] = src[r*cols + c]
template <typename T>
__global__ void transpose_cast_kernel(const float* __restrict__ src, T* __restrict__ dst, int rows, int cols) {
  pdl_entry();
  __shared__ float tile[32][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[(size_t)r * cols + c];
  }
  __syncthreads();
  int r = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int cc = blockIdx.x * 32 + j;
    if (r < rows && cc < cols) dst[(size_t)cc * rows + r] = from_f<T>(tile[threadIdx.x][j]);
  }
}

// ---- table-driven weight packing: ONE launch refreshes every compute-layout copy ------------------
__global__ void __launch_bounds__(256) pack_table_kernel(const __grid_constant__ PackTable t) {
  pdl_entry();
  const PackSeg sg = t.seg[blockIdx.y];
  if (sg.transpose) {
    // dst[c * rows + r] = src[r * cols + c] through 32 x 32 shared-memory tiles: both sides coalesced (the element-wise form
    // scattered 2-byte writes at a stride of `rows`; the two 1.2 M-element classifier matrices took 70 us per step)
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    const int tr = (sg.rows + 31) >> 5, tc = (sg.cols + 31) >> 5;
    for (int tl = blockIdx.x; tl < tr * tc; tl += gridDim.x) {
      const int r0 = (tl / tc) << 5, c0 = (tl % tc) << 5;
      __syncthreads();
      for (int j = ty; j < 32; j += 8) {
        const int r = r0 + j, c = c0 + tx;
        if (r < sg.rows && c < sg.cols) tile[j][tx] = sg.src[(int64_t)r * sg.cols + c] * (sg.row_scale ? sg.row_scale[r] : 1.f);
      }
      __syncthreads();
      for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, r = r0 + tx;
        if (r < sg.rows && c < sg.cols) {
          const int64_t o = (int64_t)c * sg.rows + r;
          if (sg.out_dtype == TD3D_BF16) reinterpret_cast<bf16*>(sg.dst)[o] = __float2bfloat16_rn(tile[tx][j]);
          else reinterpret_cast<float*>(sg.dst)[o] = tile[tx][j];
        }
      }
    }
    return;
  }
  const int64_t n = (int64_t)sg.rows * sg.cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = sg.src[i];
    if (sg.row_scale) v *= sg.row_scale[(int)(i / sg.cols)];
    if (sg.out_dtype == TD3D_BF16) reinterpret_cast<bf16*>(sg.dst)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(sg.dst)[i] = v;
  }
}
__global__ void __launch_bounds__(128) bn_fold_table_kernel(const __grid_constant__ BnFoldTable t, float eps) {
  pdl_entry();
  const BnFoldSeg sg = t.seg[blockIdx.y];
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < sg.C; c += gridDim.x * blockDim.x) {
    float invstd = 1.f / sqrtf(sg.rv[c] + eps);
    float sc = sg.gamma[c] * invstd;
    sg.scale[c] = sc;
    sg.shift[c] = sg.beta[c] - sg.rm[c] * sc + (sg.lin_bias ? sg.lin_bias[c] * sc : 0.f);
  }
}
int launch_pack_table(const PackTable& t, cudaStream_t st) {
  if (t.n <= 0) return TD3D_OK;
  TD3D_CUDA(launch_kernel(pack_table_kernel, dim3(148, t.n), 256, 0, st, t));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}
int launch_bn_fold_table(const BnFoldTable& t, float eps, cudaStream_t st) {
  if (t.n <= 0) return TD3D_OK;
  TD3D_CUDA(launch_kernel(bn_fold_table_kernel, dim3(2, t.n), 128, 0, st, t, eps));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_cast(const float* src, void* dst, int64_t n, int dtype, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  if (dtype == TD3D_BF16) TD3D_CUDA(launch_kernel(cast_kernel<bf16>, blocks, 256, 0, st, src, (bf16*)dst, n));
  else TD3D_CUDA(launch_kernel(cast_kernel<float>, blocks, 256, 0, st, src, (float*)dst, n));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}
int launch_cast_f32(const void* src, float* dst, int64_t n, int dtype, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  if (dtype == TD3D_BF16) TD3D_CUDA(launch_kernel(cast_f32_kernel<bf16>, blocks, 256, 0, st, (const bf16*)src, dst, n));
  else TD3D_CUDA(launch_kernel(cast_f32_kernel<float>, blocks, 256, 0, st, (const float*)src, dst, n));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}
int launch_transpose_cast(const float* src, void* dst, int rows, int cols, int dtype, cudaStream_t st) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32)), block(32, 8);
  if (dtype == TD3D_BF16) TD3D_CUDA(launch_kernel(transpose_cast_kernel<bf16>, grid, block, 0, st, src, (bf16*)dst, rows, cols));
  else TD3D_CUDA(launch_kernel(transpose_cast_kernel<float>, grid, block, 0, st, src, (float*)dst, rows, cols));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
