// tcgen05 / TMEM / TMA GEMMs for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   NT  Y[M,N]   = A[M,K] W[N,K]^T (+bias)(+addend)   1x1 convs fwd + dgrad, classifier Linear
//                  (reference mobilenetv3.py:120,142,148,158,192); both operands K-major
//   TN  C[N1,N2] += A[M,N1]^T B[M,N2]                  weight gradients; both operands MN-major
//
// These GEMMs stream a huge M (= B*H*W pixels) against small K/N (16..960 channels): they are
// HBM-bound (arithmetic intensity 9-140 flop/B vs a ridge of ~217), so the design goal is bytes in
// flight, not tensor-pipe occupancy:
//   * persistent CTAs, one per SM, warp-specialised: warp0 = TMA producer, warp1 = MMA issuer
//     (one elected thread, tcgen05.mma), warps 2-13 = three 4-warp epilogue groups (tcgen05.ld ->
//     registers -> global), each draining its own TMEM accumulator stage
//   * operands are staged by TMA (cp.async.bulk.tensor, hardware swizzle chosen from K so that a
//     16-channel layer does not waste 7/8 of each shared-memory row) in a multi-stage mbarrier ring;
//     the whole W operand stays resident in shared memory when it fits
//   * up to 8 accumulator stages in TMEM so the epilogues of tiles i .. i+2 overlap the MMAs of the
//     next tiles; BatchNorm statistics are reduced in the epilogue (lane-local registers or a per-warp
//     shared-memory column reduction) and leave as one vector of atomics per CTA
//   * the optional epilogue terms (bias, addend, saved-y, activation) are compile-time: one kernel
//     instance per combination the plan uses
//   * M/N/K tails are handled by TMA out-of-bounds zero fill + masked epilogue stores
// Every mbarrier wait is bounded (trap instead of hanging the GPU).
#include "td3d_kernels.h"

#include <cuda.h>
#include <stdlib.h>

namespace td3d {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t i = 0; i < (1u << 26); ++i)
    if (mbar_try_wait(bar, parity)) return;
  __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t r[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 pack8_bf16(const float x[8]) {
  uint4 raw;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
  return raw;
}

// Shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor"):
//  [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) swizzle mode
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D:
//  [4,6) D fmt=1(f32) | [7,10) A fmt=1(bf16) | [10,13) B fmt=1 | 15 A major | 16 B major |
//  [17,23) N>>3 | [24,29) M>>4
__host__ __device__ __forceinline__ uint32_t make_idesc(uint32_t M, uint32_t N, uint32_t a_mn_major, uint32_t b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

__device__ __forceinline__ float warp_transpose_sum32_tc(float v[32]) {
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

// Debug timeline (TD3D_TC_DBG bit 32): CTA 0 records %globaltimer at pipeline events of its first tiles.
__device__ unsigned long long g_tc_timeline[8 * 64];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(ev, idx)                                                                         \
  do {                                                                                            \
    if (EPI < 0 && (p.dbg & 32) && blockIdx.x == 0 && (idx) < 64) g_tc_timeline[(ev) * 64 + (idx)] = gtimer(); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// NT kernel
// ------------------------------------------------------------------------------------------------
static const int TC_THREADS = 192;          // 6 warps (TN kernel)
static const int TC_EPI_GROUPS = 3;         // NT kernel: independent 4-warp epilogue groups (one warp per TMEM lane quarter)
static const int TC_NT_THREADS = 32 * 2 + TC_EPI_GROUPS * 128;   // NT kernel: TMA producer warp, MMA issuer warp, epilogue groups
static const int TC_MAX_ACC = 8;            // TMEM accumulator stages
static const int TC_BLOCK_M = 128;
static const int TC_MAX_STAGES = 12;
static const int TC_TMEM_COLS = 512;

struct TcNtParams {
  int M, N, K;
  int block_n;          // UMMA N (multiple of 16, <= 256)
  int block_k;          // elements per k block (16 / 32 / 64) == swizzle span
  int swizzle_bytes;    // 32 / 64 / 128
  int stages;
  int a_stage_bytes, b_stage_bytes;   // ring slot sizes (b padded to 1024 B)
  int tx_bytes;         // bytes TMA actually delivers per stage (A box + W box)
  int m_tiles, n_tiles;
  bf16* y; float* yf;
  const bf16* addend; const float* bias; const bf16* ysaved;
  float* stats; int slots;
  int act;              // epilogue activation after the bias (TD3D_ACT_*), before the addend
  int n_acc, acc_stride; // TMEM accumulator stages; columns of one 128-row accumulator (power of two >= block_n)
  int m_sub;            // 128-row MMA blocks per tile (1 or 2).  The single-thread producer / issuer loops and the epilogue
                        // hand-off cost ~0.8 us per tile whatever its size (timeline in profiles/r01_v3_gemm_bench_timeline.txt);
                        // 256-row tiles (one TMA box, one barrier round trip, two tcgen05.mma per k step into adjacent
                        // accumulators) halve that per row on the long-M layers
  int stage_cols;       // TMEM columns of one accumulator stage = acc_stride * m_sub
  int w_resident;       // 1: the whole W operand is loaded ONCE per CTA into its own smem region (all 148 CTAs
                        // re-fetching the same few-KB W tile for every 128-row tile hot-spots one L2 slice)
  int epi_groups;       // active epilogue groups = min(TC_EPI_GROUPS, n_acc), see the hand-off note in the epilogue
  int dbg;              // TD3D_TC_DBG: 32 = CTA 0 records the pipeline timeline (TC_EPI_ANY instances only)
};

// Measured on B200 (TD3D_TC_DBG=32 timeline): with a single 4-warp epilogue group every 32-column chunk
// costs ~1 us of one-warp-per-scheduler latency-bound issue, the MMA/TMA side idles, and the kernel
// runs at ~1 TB/s.  Hence TC_EPI_GROUPS groups work on different tiles concurrently, each draining its
// own TMEM accumulator stage (tile sequence number ti -> stage ti % n_acc, group ti % TC_EPI_GROUPS).
// STATS: how the BatchNorm statistic sums of the epilogue are formed.  The epilogue warps bound every wide layer (the
// shuffle transpose-sum of round 1 -- 62 shuffles + 124 selects + 62 adds per 32 x 32 chunk -- doubled the kernel time:
// N=64, K=16: 151 us without statistics, 297 us with), so each flavour is its own instance:
//   TC_ST_NONE  : no statistics (inference, data gradients without a BatchNorm behind them): 94-108 registers, no spill
//   TC_ST_LOCAL : N <= 32 -- every tile shows a thread (lane = row) the same 32 columns, so it keeps its contributions in 64
//                 registers over ALL tiles of the CTA and the transpose-sum runs once per CTA (N=16: 158 -> 100 us together
//                 with the 256-row tiles).  The chunk is read from TMEM in 16-column halves (a 32-register load next to the
//                 64 live sums spilled).
//   TC_ST_SMEM  : per chunk, the warp writes its packed bf16 rows (and the saved-y rows of the data-gradient flavour) into a
//                 private shared-memory tile, then lane j adds up column j: 4-8 STS.128 + 32-64 LDS.U16 + 64 FADD / FFMA
//                 (N=64, K=16: 297 -> 217 us; 72 x 24: 102 -> 75; 240 x 40: 84 -> 61; profiles/r02_gemm_bench3.txt)
// The statistics always describe the stored bf16 values; an fp32 output has no statistics flavour.
enum { TC_ST_NONE = 0, TC_ST_LOCAL = 2, TC_ST_SMEM = 3 };
static const int TC_RED_ROW = 80;                 // bytes per row of the reduction tile: 64 B of bf16 + 16 B pad (conflict-free STS.128 / LDS.U16)
static const int TC_RED_TILE = 32 * TC_RED_ROW;       // per epilogue warp: one tile for y (+ one for the saved y)
// EPI: which optional epilogue terms exist, as a compile-time bit mask -- or TC_EPI_ANY: every term tested at run time (the
// instance behind td3d_k_gemm_nt's odd combinations, fp32 outputs and the debug timeline).  ncu on the all-run-time kernel
// (profiles/r02_ncu_gemm_nt.txt) showed the predicated-off bias / addend / saved-y code of a plain forward GEMM taking a
// quarter of the issued instructions of the epilogue warps, which bound the kernel.
enum { TC_EPI_BIAS = 1, TC_EPI_ADDEND = 2, TC_EPI_YSAVED = 4, TC_EPI_ACT = 8, TC_EPI_ANY = -1 };
template <int EPI, int STATS>
__global__ void __launch_bounds__(TC_NT_THREADS, 1)
gemm_nt_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, TcNtParams p) {
  const bool has_bias = EPI < 0 ? p.bias != nullptr : (EPI & TC_EPI_BIAS) != 0;
  const bool has_addend = EPI < 0 ? p.addend != nullptr : (EPI & TC_EPI_ADDEND) != 0;
  const bool has_ysaved = EPI < 0 ? p.ysaved != nullptr : (EPI & TC_EPI_YSAVED) != 0;
  const bool has_act = EPI < 0 ? p.act != TD3D_ACT_NONE : (EPI & TC_EPI_ACT) != 0;
  const bool out_f32 = EPI < 0 && p.yf != nullptr;
  pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[TC_MAX_STAGES], s_empty[TC_MAX_STAGES], s_tfull[TC_MAX_ACC], s_tempty[TC_MAX_ACC];
  __shared__ __align__(8) uint64_t s_wfull;
  __shared__ uint32_t s_tmem_base;
  __shared__ float s_stat[TC_EPI_GROUPS][4][2][256];     // per epilogue warp: no atomics while a tile is reduced

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1024-byte aligned operand ring (SWIZZLE_128B atoms must be 1024B aligned)
  const uint32_t k_blocks_u = (uint32_t)((p.K + p.block_k - 1) / p.block_k);
  const uint32_t wres = (smem_u32(smem_raw) + 1023u) & ~1023u;      // resident W: [n_tiles][k_blocks][b_stage]
  const uint32_t ring = wres + (p.w_resident ? (uint32_t)p.n_tiles * k_blocks_u * (uint32_t)p.b_stage_bytes : 0u);
  const uint32_t stage_bytes = (uint32_t)(p.a_stage_bytes + (p.w_resident ? 0 : p.b_stage_bytes));
  // epilogue staging: per epilogue warp two [32 rows x 64 B] SWIZZLE_64B boxes (1 KB aligned)
  const uint32_t ystage = (ring + (uint32_t)p.stages * stage_bytes + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&s_full[s]), 1); mbar_init(smem_u32(&s_empty[s]), 1); }
    for (int s = 0; s < p.n_acc; ++s) { mbar_init(smem_u32(&s_tfull[s]), 1); mbar_init(smem_u32(&s_tempty[s]), 4); }
    mbar_init(smem_u32(&s_wfull), 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < TC_EPI_GROUPS * 4 * 2 * 256; i += blockDim.x) (&s_stat[0][0][0][0])[i] = 0.f;
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_w); }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem_base), TC_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  const int num_tiles = p.m_tiles * p.n_tiles;
  const int k_blocks = (p.K + p.block_k - 1) / p.block_k;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      if (p.w_resident && blockIdx.x < num_tiles) {
        const uint32_t wf = smem_u32(&s_wfull);
        mbar_expect_tx(wf, (uint32_t)(p.n_tiles * k_blocks * p.block_n * p.swizzle_bytes));
        for (int nt = 0; nt < p.n_tiles; ++nt)
          for (int kb = 0; kb < k_blocks; ++kb)
            tma_load_2d(wres + (uint32_t)((nt * k_blocks + kb) * p.b_stage_bytes), &map_w, wf, kb * p.block_k, nt * p.block_n);
      }
      int stage = 0; uint32_t phase = 0;
      int ti = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        // single-thread loops: no runtime divisions on the common single-N-tile path (timeline: the MMA
        // issuer needs ~800 ns per 128-row tile of a K=16 layer, which caps such layers at ~1.4 TB/s)
        const int m0 = (p.n_tiles == 1 ? tile : tile / p.n_tiles) * TC_BLOCK_M * p.m_sub;
        const int n0 = p.n_tiles == 1 ? 0 : (tile % p.n_tiles) * p.block_n;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(smem_u32(&s_empty[stage]), phase ^ 1u);
          TC_STAMP(0, ti);
          const uint32_t full = smem_u32(&s_full[stage]);
          const uint32_t a_dst = ring + stage * stage_bytes, b_dst = a_dst + p.a_stage_bytes;
          mbar_expect_tx(full, (uint32_t)(p.w_resident ? p.a_stage_bytes : p.tx_bytes));
          tma_load_2d(a_dst, &map_a, full, kb * p.block_k, m0);
          if (!p.w_resident) tma_load_2d(b_dst, &map_w, full, kb * p.block_k, n0);
          TC_STAMP(1, ti);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one elected thread) =====================
    // Timeline (TD3D_TC_DBG=32, profiles/r02_gemm_timeline.txt): the thread needs ~0.8 us per tile for two barrier waits,
    // descriptors, tcgen05.mma and 2 x tcgen05.commit, which bounds the K=16 layers.  A second issuer warp (round 1: +20 % on
    // those layers) was removed: with two issuers the accumulator stages are committed out of tile order, nothing bounds how
    // far one issuer may fall behind the other, and an epilogue group's parity wait could then alias a stage's
    // previous-but-one phase.  The 256-row tiles (m_sub) are the safe way to halve the per-row cost.
    if (lane == 0) {
      const uint32_t idesc = make_idesc(TC_BLOCK_M, (uint32_t)p.block_n, 0, 0);
      const uint32_t layout_type = p.swizzle_bytes == 128 ? 2u : (p.swizzle_bytes == 64 ? 4u : 6u);
      const uint32_t sbo = 8u * (uint32_t)p.swizzle_bytes;     // 8 rows of one swizzle span
      int stage = 0; uint32_t phase = 0;
      if (p.w_resident && blockIdx.x < num_tiles) mbar_wait(smem_u32(&s_wfull), 0);
      int ti = 0;
      int as = 0;                       // accumulator stage ti % n_acc and its phase (ti / n_acc) & 1, kept incrementally
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++ti) {
        const int n_tile = p.n_tiles == 1 ? 0 : tile % p.n_tiles;
        mbar_wait(smem_u32(&s_tempty[as]), aphase ^ 1u);          // epilogue drained this accumulator
        TC_STAMP(2, ti);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * p.stage_cols);
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(smem_u32(&s_full[stage]), phase);
          TC_STAMP(3, ti);
          tc_fence_after();
          const uint32_t a_src = ring + stage * stage_bytes;
          const uint32_t b_src = p.w_resident ? wres + (uint32_t)((n_tile * k_blocks + kb) * p.b_stage_bytes) : a_src + p.a_stage_bytes;
          const int k_left = p.K - kb * p.block_k;
          const int k_steps = ((k_left < p.block_k ? k_left : p.block_k) + 15) >> 4;   // UMMA_K = 16 (bf16)
          for (int h = 0; h < p.m_sub; ++h) {                      // rows h * 128 ... of the tile -> accumulator h of the stage
            const uint32_t a_h = a_src + (uint32_t)(h * TC_BLOCK_M * p.swizzle_bytes);
            for (int ks = 0; ks < k_steps; ++ks) {
              const uint64_t da = make_smem_desc(a_h + ks * 32, 16, sbo, layout_type);     // LBO is ignored for K-major swizzled operands
              const uint64_t db = make_smem_desc(b_src + ks * 32, 16, sbo, layout_type);
              umma_bf16(d_tmem + (uint32_t)(h * p.acc_stride), da, db, idesc, (kb | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&s_empty[stage]));                 // frees the smem slot when MMAs retire
          if (kb == k_blocks - 1) umma_commit(smem_u32(&s_tfull[as]));
          TC_STAMP(4, ti);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (++as == p.n_acc) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue warps (TMEM -> registers -> global) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int eg = (warp - 2) >> 2;         // epilogue group (any 4 consecutive warps cover the 4 quarters)
    const int et = (threadIdx.x - 64) & 127;   // 0..127 within the epilogue group
    const int n_chunks = (p.block_n + 31) >> 5;
    float (*gstat)[2][256] = s_stat[eg];
    const ActK eak = make_actk(has_act ? p.act : TD3D_ACT_NONE);
    // TC_ST_SMEM: this warp's reduction tiles
    const uint32_t red_y = ystage + (uint32_t)((eg * 4 + q) * (has_ysaved ? 2 : 1) * TC_RED_TILE);
    const uint32_t red_s = red_y + (uint32_t)TC_RED_TILE;
    int as = eg % p.n_acc;
    uint32_t aphase = (uint32_t)(eg / p.n_acc) & 1u;
    // lane = row, v[j] / w2[j] = this row's contribution to the two sums of column j of the current 32-column chunk
    constexpr int NV = STATS == TC_ST_LOCAL ? 32 : 1;
    float v[NV], w2[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) { v[i] = 0.f; w2[i] = 0.f; }
    // Accumulator hand-off safety (root cause of the round-1 "launch failure at batch >= 1024" and of the rare aborts of
    // stat-less GEMMs): tile ti uses barrier s_tfull[ti % n_acc]; a group that finished tile ti waits next for tile
    // ti + G on barrier (ti + G) % n_acc with a PARITY wait, which is only meaningful if the previous phase of that
    // barrier -- tile ti + G - n_acc -- has already completed.  With one in-order issuer that holds iff
    // ti + G - n_acc <= ti, i.e. G <= n_acc.  Wide tiles (block_n > 128 -> n_acc = 2) with 3 groups broke it: the group
    // could pass a wait on the not-yet-committed tile ti + 1, read a stage still being drained and release it early.
    // Hence only min(TC_EPI_GROUPS, n_acc) groups take tiles.
    const int G = p.epi_groups;
    for (int ti = eg, tile = blockIdx.x + eg * gridDim.x; eg < G && tile < num_tiles; tile += G * gridDim.x, ti += G) {
      const int m_tile = p.n_tiles == 1 ? tile : tile / p.n_tiles;
      const int n0 = p.n_tiles == 1 ? 0 : (tile % p.n_tiles) * p.block_n;
      mbar_wait(smem_u32(&s_tfull[as]), aphase);
      if (q == 2 && lane == 0) TC_STAMP(5, ti);
      tc_fence_after();
      for (int hc = 0; hc < p.m_sub * n_chunks; ++hc) {
        const int ch = p.m_sub == 2 ? hc >> 1 : hc, h = p.m_sub == 2 ? hc & 1 : 0;   // 32-column chunk, 128-row block of the tile
        const int m0 = (m_tile * p.m_sub + h) * TC_BLOCK_M;
        const int m = m0 + q * 32 + lane;
        const bool row_ok = m < p.M;
        const int nb = n0 + ch * 32;
        constexpr int NR = STATS == TC_ST_LOCAL ? 16 : 32;
        uint32_t r[NR];
        const uint32_t t_chunk = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * p.stage_cols + h * p.acc_stride + ch * 32);
        if (STATS != TC_ST_LOCAL) tmem_ld32(t_chunk, *reinterpret_cast<uint32_t(*)[32]>(r));
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (STATS == TC_ST_LOCAL && (g & 1) == 0) tmem_ld16(t_chunk + g * 8, *reinterpret_cast<uint32_t(*)[16]>(r));
          const int n = nb + g * 8;
          const bool ok = row_ok && n < p.N && (n - n0) < p.block_n;
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(r[(STATS == TC_ST_LOCAL ? (g & 1) : g) * 8 + i]);
          if (ok) {
            if (has_bias) {
              float bb[8];
              loadf8(p.bias + n, bb);
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] += bb[i];
            }
            if (has_act) {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = actk_fwd(x[i], eak);
            }
            const size_t off = (size_t)m * p.N + n;
            if (has_addend) {
              float ad[8];
              load8(p.addend + off, ad);
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] += ad[i];
            }
            if (out_f32) {
              store8(p.yf + off, x);
            } else {
              const uint4 pk = pack8_bf16(x);
              *reinterpret_cast<uint4*>(p.y + off) = pk;
              if (STATS == TC_ST_SMEM) {
                sts_v4(red_y + (uint32_t)(lane * TC_RED_ROW + g * 16), pk);
                if (has_ysaved) sts_v4(red_s + (uint32_t)(lane * TC_RED_ROW + g * 16), __ldg(reinterpret_cast<const uint4*>(p.ysaved + off)));
              }
              if (STATS == TC_ST_LOCAL) {     // the statistics describe the stored (rounded) values
                x[0] = __uint_as_float(pk.x << 16); x[1] = __uint_as_float(pk.x & 0xffff0000u);
                x[2] = __uint_as_float(pk.y << 16); x[3] = __uint_as_float(pk.y & 0xffff0000u);
                x[4] = __uint_as_float(pk.z << 16); x[5] = __uint_as_float(pk.z & 0xffff0000u);
                x[6] = __uint_as_float(pk.w << 16); x[7] = __uint_as_float(pk.w & 0xffff0000u);
              }
            }
            if (STATS == TC_ST_LOCAL) {
              float ys[8];
              if (has_ysaved) load8(p.ysaved + off, ys);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[(g * 8 + i) % NV] += x[i];
                w2[(g * 8 + i) % NV] = fmaf(x[i], has_ysaved ? ys[i] : x[i], w2[(g * 8 + i) % NV]);
              }
            }
          } else if (STATS == TC_ST_SMEM) {
            sts_v4(red_y + (uint32_t)(lane * TC_RED_ROW + g * 16), make_uint4(0u, 0u, 0u, 0u));      // rows / columns outside the matrix add nothing
            if (has_ysaved) sts_v4(red_s + (uint32_t)(lane * TC_RED_ROW + g * 16), make_uint4(0u, 0u, 0u, 0u));
          }
        }
        // column sums of this warp's rows, accumulated over ALL tiles of the CTA in the warp's own shared-memory
        // row (every tile of a CTA covers the same N tile, see the launcher): no barrier, no atomics per tile
        if (STATS == TC_ST_SMEM) {
          __syncwarp();
          float s1 = 0.f, s2 = 0.f;
          const uint32_t cy = red_y + (uint32_t)(lane * 2), cs = red_s + (uint32_t)(lane * 2);
          if (has_ysaved) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              uint32_t a, b;
              asm volatile("ld.shared.u16 %0, [%1];" : "=r"(a) : "r"(cy + (uint32_t)(i * TC_RED_ROW)));
              asm volatile("ld.shared.u16 %0, [%1];" : "=r"(b) : "r"(cs + (uint32_t)(i * TC_RED_ROW)));
              const float fa = __uint_as_float(a << 16);
              s1 += fa;
              s2 = fmaf(fa, __uint_as_float(b << 16), s2);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              uint32_t a;
              asm volatile("ld.shared.u16 %0, [%1];" : "=r"(a) : "r"(cy + (uint32_t)(i * TC_RED_ROW)));
              const float fa = __uint_as_float(a << 16);
              s1 += fa;
              s2 = fmaf(fa, fa, s2);
            }
          }
          gstat[q][0][ch * 32 + lane] += s1;
          gstat[q][1][ch * 32 + lane] += s2;
          __syncwarp();                                   // the tile is rewritten by the next chunk
        }
      }
      tc_fence_before();
      __syncwarp();
      if (q == 2 && lane == 0) TC_STAMP(6, ti);
      if (lane == 0) mbar_arrive(smem_u32(&s_tempty[as]));
      as += G;                                   // stage / phase of tile ti + G
      while (as >= p.n_acc) { as -= p.n_acc; aphase ^= 1u; }
    }
    if (STATS == TC_ST_LOCAL && eg < G) {
      const float t1 = warp_transpose_sum32_tc(*reinterpret_cast<float(*)[32]>(v));
      const float t2 = warp_transpose_sum32_tc(*reinterpret_cast<float(*)[32]>(w2));
      gstat[q][0][lane] += t1;
      gstat[q][1][lane] += t2;
    }
    // one flush per CTA and epilogue group: the four lane-quarter rows are added and leave as one atomic per column.
    // The slot only spreads the atomics of the CTAs (the finalize kernels add all slots).
    if (STATS != TC_ST_NONE && eg < G && (int)blockIdx.x + eg * (int)gridDim.x < num_tiles) {
      asm volatile("bar.sync %0, 128;" ::"r"(eg + 1) : "memory");
      const int n0 = p.n_tiles == 1 ? 0 : ((int)blockIdx.x % p.n_tiles) * p.block_n;
      const int slot = ((int)blockIdx.x / p.n_tiles) % p.slots;
      for (int j = et; j < 2 * p.block_n; j += 128) {
        const int which = j / p.block_n, nn = j % p.block_n;
        const float tot = (gstat[0][which][nn] + gstat[1][which][nn]) + (gstat[2][which][nn] + gstat[3][which][nn]);
        if (n0 + nn < p.N) atomicAdd(&p.stats[((size_t)slot * 2 + which) * p.N + n0 + nn], tot);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TC_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// TN kernel (weight gradients): C[N1,N2] += A[M,N1]^T B[M,N2], one output tile + one M slice per CTA
// ------------------------------------------------------------------------------------------------
static const int TN_BK = 64;                // M rows per stage (4 UMMA_K steps)
static const int TN_STAGES = 4;
static const int TN_BOX_BYTES = TN_BK * 128;   // one [64 rows x 64 channels] box

struct TcTnParams {
  int M, N1, N2;
  int n2_block;         // UMMA N (multiple of 16, <= 256)
  int n2_boxes;         // ceil(n2_block / 64)
  int a_boxes;          // 64-channel boxes of the A operand (1 when N1 <= 64: the MMA still multiplies 128 channel rows, the
                        // upper 64 read the neighbouring B box and their results are never stored)
  int m_per_part;       // multiple of TN_BK
  float* c;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, TcTnParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_full[TN_STAGES], s_empty[TN_STAGES], s_tfull;
  __shared__ uint32_t s_tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_bytes = (uint32_t)p.a_boxes * TN_BOX_BYTES, b_bytes = (uint32_t)p.n2_boxes * TN_BOX_BYTES;
  const uint32_t stage_bytes = a_bytes + b_bytes;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TN_STAGES; ++s) { mbar_init(smem_u32(&s_full[s]), 1); mbar_init(smem_u32(&s_empty[s]), 1); }
    mbar_init(smem_u32(&s_tfull), 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem_base), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  const int n1_0 = blockIdx.x * 128, n2_0 = blockIdx.y * p.n2_block;
  const int ms = blockIdx.z * p.m_per_part;
  const int me = min(p.M, ms + p.m_per_part);
  const int k_blocks = (me - ms + TN_BK - 1) / TN_BK;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(smem_u32(&s_empty[stage]), phase ^ 1u);
        const uint32_t full = smem_u32(&s_full[stage]);
        const uint32_t a_dst = ring + stage * stage_bytes, b_dst = a_dst + a_bytes;
        mbar_expect_tx(full, stage_bytes);
        const int row = ms + kb * TN_BK;
        tma_load_2d(a_dst, &map_a, full, n1_0, row);
        if (p.a_boxes == 2) tma_load_2d(a_dst + TN_BOX_BYTES, &map_a, full, n1_0 + 64, row);
        for (int j = 0; j < p.n2_boxes; ++j) tma_load_2d(b_dst + j * TN_BOX_BYTES, &map_b, full, n2_0 + j * 64, row);
        if (++stage == TN_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, (uint32_t)p.n2_block, 1, 1);
      // MN-major, 128B swizzle: atom = 64 channels x 8 rows (1024 B). Rows (the contraction index)
      // advance by 128 B; 8-row groups are SBO apart; 64-channel blocks are LBO apart.
      const uint32_t lbo = TN_BOX_BYTES, sbo = 1024;
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(smem_u32(&s_full[stage]), phase);
        tc_fence_after();
        const uint32_t a_src = ring + stage * stage_bytes, b_src = a_src + a_bytes;
        for (int ks = 0; ks < TN_BK / 16; ++ks) {
          const uint64_t da = make_smem_desc(a_src + ks * 2048, lbo, sbo, 2u);
          const uint64_t db = make_smem_desc(b_src + ks * 2048, lbo, sbo, 2u);
          umma_bf16(tmem_base, da, db, idesc, (kb | ks) != 0 ? 1u : 0u);
        }
        umma_commit(smem_u32(&s_empty[stage]));
        if (kb == k_blocks - 1) umma_commit(smem_u32(&s_tfull));
        if (++stage == TN_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (k_blocks > 0) {
    const int q = warp & 3;
    mbar_wait(smem_u32(&s_tfull), 0);
    tc_fence_after();
    const int n1 = n1_0 + q * 32 + lane;
    const int n_chunks = (p.n2_block + 31) >> 5;
    for (int ch = 0; ch < n_chunks; ++ch) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 32), r);
      if (n1 < p.N1) {
        // 16-byte vector reductions (red.global.add.v4.f32, sm_90+): the split over M ends in 128 x n2_block atomics per CTA,
        // which bounded the short-M layers (3 M scalar atomics on a 50176 x 480 x 80 problem); N2 and the chunk starts are
        // multiples of 8, so every group of 4 columns is 16-byte aligned and either entirely inside the matrix or outside
        float* row = p.c + (size_t)n1 * p.N2 + n2_0 + ch * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int nn = ch * 32 + i;
          if (nn < p.n2_block && n2_0 + nn < p.N2)
            atomicAdd(reinterpret_cast<float4*>(row + i), make_float4(__uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                                                       __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2D row-major bf16 tensor [rows][cols], box [box_rows][box_cols]
static int make_map_2d(CUtensorMap* map, const void* base, int rows, int cols, int box_rows, int box_cols,
                       int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_last_error("cuTensorMapEncodeTiled unavailable (driver too old?)"); return TD3D_ECUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d box=%dx%d sw=%d base=%p", (int)r, rows, cols,
                   box_rows, box_cols, swizzle_bytes, base);
    return TD3D_ECUDA;
  }
  return TD3D_OK;
}

static int env_raw(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
// Tuning / debugging knobs of the NT GEMM.  Read ONCE (getenv on every eager launch showed up in host profiles);
// TD3D_TC_LIVE_ENV=1 (micro-benchmarks that flip knobs inside one process) re-reads them on every launch.
struct TcKnobs { int dbg, max_bn, m_sub; };
static TcKnobs read_knobs() {
  TcKnobs k;
  k.dbg = env_raw("TD3D_TC_DBG", 0);
  k.m_sub = env_raw("TD3D_TC_MSUB", 0);            // 0 auto, 1 / 2 force (A/B measurements)
  if (k.m_sub < 0 || k.m_sub > 2) k.m_sub = 0;
  k.max_bn = env_raw("TD3D_TC_MAXBN", 0);        // 0 = rule in launch_gemm_nt_tc
  return k;
}
static const TcKnobs& knobs() {
  static const bool live = env_raw("TD3D_TC_LIVE_ENV", 0) != 0;
  static TcKnobs k = read_knobs();
  if (live) k = read_knobs();
  return k;
}

int tc_timeline_read(unsigned long long* out, int n) {
  if (n > 8 * 64) n = 8 * 64;
  TD3D_CUDA(cudaMemcpyFromSymbol(out, g_tc_timeline, sizeof(unsigned long long) * n));
  return TD3D_OK;
}

bool tc_gemm_supported(int M, int N, int K) { return M > 0 && N >= 8 && K >= 8 && (N % 8) == 0 && (K % 8) == 0; }

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

typedef void (*NtKernel)(const CUtensorMap, const CUtensorMap, TcNtParams);
struct NtInstance { int epi, stats; NtKernel fn; };
#define TD3D_NT(E, S) {E, S, gemm_nt_tc_kernel<E, S>}
static const NtInstance kNtInstances[] = {
    // training forward (1x1 convs), data gradient of the projection conv
    TD3D_NT(0, TC_ST_NONE), TD3D_NT(0, TC_ST_SMEM), TD3D_NT(0, TC_ST_LOCAL),
    // data gradient of the expansion conv: + sums of g and g * saved-y for the previous block's BatchNorm (+ residual gradient)
    TD3D_NT(TC_EPI_YSAVED, TC_ST_SMEM), TD3D_NT(TC_EPI_YSAVED, TC_ST_LOCAL),
    TD3D_NT(TC_EPI_YSAVED | TC_EPI_ADDEND, TC_ST_SMEM), TD3D_NT(TC_EPI_YSAVED | TC_EPI_ADDEND, TC_ST_LOCAL),
    TD3D_NT(TC_EPI_ADDEND, TC_ST_NONE),
    // classifier Linear (bias) with BatchNorm1d statistics
    TD3D_NT(TC_EPI_BIAS, TC_ST_SMEM),
    // inference: folded BatchNorm bias (+ activation | + residual)
    TD3D_NT(TC_EPI_BIAS | TC_EPI_ACT, TC_ST_NONE), TD3D_NT(TC_EPI_BIAS, TC_ST_NONE), TD3D_NT(TC_EPI_BIAS | TC_EPI_ADDEND, TC_ST_NONE),
    // everything else (fp32 output, td3d_k_gemm_nt combinations, debug timeline)
    TD3D_NT(TC_EPI_ANY, TC_ST_NONE), TD3D_NT(TC_EPI_ANY, TC_ST_SMEM), TD3D_NT(TC_EPI_ANY, TC_ST_LOCAL),
};
#undef TD3D_NT

int launch_gemm_nt_tc(const GemmNT& g, cudaStream_t st) {
  TD3D_REQUIRE(tc_gemm_supported(g.M, g.N, g.K), "gemm_nt_tc: unsupported shape M=%d N=%d K=%d", g.M, g.N, g.K);
  TD3D_REQUIRE(((uintptr_t)g.a & 15) == 0 && ((uintptr_t)g.w & 15) == 0, "gemm_nt_tc: operands must be 16B aligned");
  TcNtParams p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  const TcKnobs& kn = knobs();
  int sw = 128;
  if (g.K <= 16) sw = 32;
  else if (g.K <= 32) sw = 64;
  p.swizzle_bytes = sw;
  p.block_k = sw / 2;
  // N tiling: equal tiles of <= 256 columns, each a multiple of 16
  // N tile width.  256-column tiles leave 2 TMEM accumulator stages (hence 2 epilogue groups, see the hand-off note in the
  // kernel); 128-column tiles keep 4 stages and all 3 groups at the price of fetching the A tile once per N tile.  Measured
  // (scripts/gemm_bench2.py, profiles/r02_gemm_bench2.txt): 128 wins 10-20 % on every wide-N layer with K <= 112 and loses
  // 15 % on 160 x 960 (K = 960: the A tile is the traffic).  TD3D_TC_MAXBN overrides.
  int max_bn = kn.max_bn >= 16 && kn.max_bn <= 256 ? kn.max_bn : (g.K <= 256 ? 128 : 256);
  int n_tiles = ceil_div(g.N, max_bn);
  int bn = ceil_div(ceil_div(g.N, n_tiles), 16) * 16;
  p.block_n = bn;
  p.n_tiles = ceil_div(g.N, bn);
  // 256-row tiles where every CTA still gets several of them.  Measured (scripts/gemm_bench3.py, profiles/r02_gemm_bench3.txt,
  // M = 0.2 .. 3.2 M rows): they win for N <= 40 with or without statistics (N=16: 137 -> 86 us, N=24: 47 -> 36 us, N=40:
  // 24 -> 20 us); at N = 64 they tie with statistics (217 vs 220 us) and lose without (151 -> 195 us: two 64-column
  // accumulators per tile leave 4 TMEM stages for 3 epilogue groups); for N > 64 they leave 2 stages, hence 2 groups, and lose.
  const bool long_m = (int64_t)ceil_div(g.M, 2 * TC_BLOCK_M) * p.n_tiles >= 3 * (int64_t)num_sms();
  p.m_sub = kn.m_sub ? kn.m_sub : ((long_m && bn <= 48) ? 2 : 1);
  if (bn > 128) p.m_sub = 1;
  p.m_tiles = ceil_div(g.M, TC_BLOCK_M * p.m_sub);
  p.a_stage_bytes = TC_BLOCK_M * p.m_sub * sw;
  p.b_stage_bytes = ceil_div(bn * sw, 1024) * 1024;
  p.tx_bytes = p.a_stage_bytes + bn * sw;
  // statistics flavour (see the kernel): lane-local sums where a thread keeps seeing the same 32 columns, else the
  // shared-memory column reduction
  TD3D_REQUIRE(!(g.stats && g.out_f32), "gemm_nt_tc: the statistics epilogue describes the stored bf16 values (no fp32 output)");
  const int st_flavour = !g.stats ? TC_ST_NONE : (bn <= 32 ? TC_ST_LOCAL : TC_ST_SMEM);
  const int k_blocks = ceil_div(g.K, p.block_k);
  const int wres_bytes = p.n_tiles * k_blocks * p.b_stage_bytes;
  p.w_resident = (wres_bytes <= (st_flavour == TC_ST_SMEM ? 48 : 96) * 1024 ) ? 1 : 0;
  int stage_bytes = p.a_stage_bytes + (p.w_resident ? 0 : p.b_stage_bytes);
  // (a shared-memory staged TMA store of the output was measured twice, in round 1 and again with the lean epilogues of
  // round 2, profiles/r02_gemm_bench4.txt: within +-3 % of the per-thread 16-byte stores on every layer, so it was removed)
  const int ystage_bytes = st_flavour == TC_ST_SMEM ? TC_EPI_GROUPS * 4 * (g.ysaved ? 2 : 1) * TC_RED_TILE + 1024 : 0;
  int budget = 176 * 1024 - (p.w_resident ? wres_bytes : 0) - ystage_bytes;
  p.stages = budget / stage_bytes;
  if (p.stages > TC_MAX_STAGES) p.stages = TC_MAX_STAGES;
  if (p.stages < 2) p.stages = 2;
  p.y = g.out_f32 ? nullptr : (bf16*)g.y;
  p.yf = g.out_f32 ? (float*)g.y : nullptr;
  p.addend = (const bf16*)g.addend; p.bias = g.bias; p.ysaved = (const bf16*)g.ysaved;
  p.stats = g.stats; p.slots = g.slots > 0 ? g.slots : 1;
  p.act = g.act;
  p.dbg = kn.dbg;
  p.acc_stride = 32;
  while (p.acc_stride < bn) p.acc_stride <<= 1;
  p.stage_cols = p.acc_stride * p.m_sub;
  p.n_acc = TC_TMEM_COLS / p.stage_cols;
  if (p.n_acc > TC_MAX_ACC) p.n_acc = TC_MAX_ACC;
  p.epi_groups = p.n_acc < TC_EPI_GROUPS ? p.n_acc : TC_EPI_GROUPS;
  CUtensorMap map_a, map_w;
  TD3D_TRY(make_map_2d(&map_a, g.a, g.M, g.K, TC_BLOCK_M * p.m_sub, p.block_k, sw));
  TD3D_TRY(make_map_2d(&map_w, g.w, g.N, g.K, bn, p.block_k, sw));
  size_t smem = (size_t)p.stages * stage_bytes + (p.w_resident ? wres_bytes : 0) + ystage_bytes + 1024;
  TD3D_REQUIRE(smem <= 190 * 1024, "gemm_nt_tc: %zu bytes of shared memory for M=%d N=%d K=%d", smem, g.M, g.N, g.K);
  // instance: the exact epilogue mask where it is one of the plan's combinations, else the all-run-time kernel
  int epi = (g.bias ? TC_EPI_BIAS : 0) | (g.addend ? TC_EPI_ADDEND : 0) | (g.ysaved ? TC_EPI_YSAVED : 0) | (g.act != TD3D_ACT_NONE ? TC_EPI_ACT : 0);
  if (g.out_f32 || kn.dbg) epi = TC_EPI_ANY;
  const NtInstance* inst = nullptr;
  for (const NtInstance& c : kNtInstances)
    if (c.stats == st_flavour && (c.epi == epi || (!inst && c.epi == TC_EPI_ANY))) { inst = &c; if (c.epi == epi) break; }
  TD3D_REQUIRE(inst != nullptr, "gemm_nt_tc: no kernel instance for statistics flavour %d", st_flavour);
  static bool attr_set = false;
  if (!attr_set) {
    for (const NtInstance& c : kNtInstances) TD3D_CUDA(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024));
    attr_set = true;
  }
  int grid = p.m_tiles * p.n_tiles;
  if (grid > num_sms()) grid = num_sms();
  // tile t covers N tile t % n_tiles and CTA c takes tiles c, c + grid, ...: with grid a multiple of n_tiles every tile of
  // a CTA lies in the same N tile, which lets the statistics epilogue keep one accumulator row per CTA (flushed once)
  grid -= grid % p.n_tiles;
  TD3D_REQUIRE(p.act == TD3D_ACT_NONE || !p.stats, "gemm_nt_tc: the activation epilogue (inference) has no statistics flavour");
  TD3D_CUDA(launch_kernel(inst->fn, grid, TC_NT_THREADS, smem, st, map_a, map_w, p));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

int launch_gemm_tn_tc(const GemmTN& g, cudaStream_t st) {
  TD3D_REQUIRE(g.M > 0 && g.N1 % 8 == 0 && g.N2 % 8 == 0, "gemm_tn_tc: unsupported shape M=%d N1=%d N2=%d", g.M, g.N1, g.N2);
  TD3D_REQUIRE(((uintptr_t)g.c & 15) == 0, "gemm_tn_tc: the output must be 16-byte aligned (vector reductions)");
  TcTnParams p;
  p.M = g.M; p.N1 = g.N1; p.N2 = g.N2;
  int n2_tiles = ceil_div(g.N2, 256);
  p.n2_block = ceil_div(ceil_div(g.N2, n2_tiles), 16) * 16;
  n2_tiles = ceil_div(g.N2, p.n2_block);
  p.n2_boxes = ceil_div(p.n2_block, 64);
  p.a_boxes = g.N1 <= 64 ? 1 : 2;
  int n1_tiles = ceil_div(g.N1, 128);
  int tiles = n1_tiles * n2_tiles;
  // M split: every part ends with 128 x n2_block fp32 atomics into the same tile, so more parts buy
  // streaming parallelism at the price of atomic traffic (measured: 196 parts of 256 rows spent 60 us on a
  // 27 MB problem).  Pick the split that minimises a simple cost model: operand streaming at
  // ring-bytes / latency per CTA and ~5 TB/s per chip, plus ~300 G reduced elements/s, plus waves of 1 CTA per SM.
  int parts = 1;
  {
    const int max_parts = ceil_div(g.M, TN_BK * 4);
    const double row_bytes = 2.0 * ((g.N1 < 128 ? g.N1 : 128) + (g.N2 < p.n2_block ? g.N2 : p.n2_block));
    double best = 1e300;
    for (int cand = 1; cand <= max_parts && tiles * cand <= num_sms() * 2; cand = cand < 8 ? cand + 1 : cand + cand / 4) {
      const int rows = ceil_div(ceil_div(g.M, cand), TN_BK) * TN_BK;
      const int n_ctas = tiles * ceil_div(g.M, rows);
      const int conc = n_ctas < num_sms() ? n_ctas : num_sms();
      double rate = 5000.0 / conc;                  // GB/s per CTA: chip share, capped by the bytes the
      const double cap = TN_STAGES * TN_BK * row_bytes / 1.5e3;   // TMA ring keeps in flight over ~1.5 us of latency
      if (rate > cap) rate = cap;
      const double stream_us = (double)ceil_div(n_ctas, num_sms()) * rows * row_bytes / (rate * 1e3);
      const double atom_us = (double)n_ctas * 128.0 * p.n2_block / 3e5;      // fp32 elements per us through 16-byte reductions (the
                                                                             // total is flat between 1.5e5 and 1.2e6: 570-598 us)
      const double cost = stream_us + atom_us;
      if (cost < best) { best = cost; parts = cand; }
    }
  }
  p.m_per_part = ceil_div(ceil_div(g.M, parts), TN_BK) * TN_BK;
  parts = ceil_div(g.M, p.m_per_part);
  p.c = g.c;
  CUtensorMap map_a, map_b;
  TD3D_TRY(make_map_2d(&map_a, g.a, g.M, g.N1, TN_BK, 64, 128));
  TD3D_TRY(make_map_2d(&map_b, g.b, g.M, g.N2, TN_BK, 64, 128));
  size_t smem = (size_t)TN_STAGES * (p.a_boxes + p.n2_boxes) * TN_BOX_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    TD3D_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
    attr_set = true;
  }
  dim3 grid(n1_tiles, n2_tiles, parts);
  TD3D_CUDA(launch_kernel(gemm_tn_tc_kernel, grid, TC_THREADS, smem, st, map_a, map_b, p));
  TD3D_LAUNCH_CHECK();
  return TD3D_OK;
}

}  // namespace td3d
