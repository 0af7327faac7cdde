// Internal launcher declarations (host side). Every launcher is asynchronous on `st`.
#pragma once
#include "td3d_common.cuh"

namespace td3d {

// ---- k_elementwise.cu ----
int launch_apply_xform(const void* y, const XForm& xf, const void* res, void* out, float* pool_stats,
                       int B, int HW, int C, int dtype, cudaStream_t st);
int launch_affine2(const void* g, const void* y, const float* alpha, const float* beta, const float* gamma,
                   void* out, int B, int HW, int C, int dtype, cudaStream_t st);
int launch_act_bwd_stats(const void* g, const float* g_pooled, float g_scale, const void* y, const XForm& xf,
                         void* gu, float* stats, int B, int HW, int C, int dtype, cudaStream_t st,
                         const void* addend = nullptr);
int launch_pool_finalize(const float* stats, float scale, void* out, int B, int C, int dtype, cudaStream_t st);

// ---- k_bn.cu ----
struct BnFwdArgs {
  const float* stats; int slots; double count;          // [slots][2][C]
  const float* gamma; const float* beta;                // param arena
  float* running_mean; float* running_var; int64_t* nbt;  // may be null (no update)
  float* scale; float* shift; float* mean; float* invstd;
  int C; float momentum; float eps;
};
int launch_bn_finalize_fwd(const BnFwdArgs& a, cudaStream_t st);
int launch_bn_eval_fold(const float* gamma, const float* beta, const float* rm, const float* rv,
                        float* scale, float* shift, int C, float eps, cudaStream_t st);
struct BnBwdArgs {
  const float* stats; int slots;                        // P1,P2 : [slots][2][C]
  const float* mean; const float* invstd; const float* gamma;
  const float* se; const float* g_pool; const float* fwd_pool;  // SE terms ([B,C],[B,C],[B][2][C]) or null
  float* alpha; float* beta; float* gammac;             // [B,C], [C], [B,C]
  float* dgamma; float* dbeta;                          // grads arena
  int B; int HW; int C;
};
int launch_bn_bwd_finalize(const BnBwdArgs& a, cudaStream_t st);

// ---- k_se.cu ----
struct SeArgs {
  const float* pool_stats;   // [B][2][C] (which=0 is the per-(b,c) sum over pixels)
  const float* scale; const float* shift;   // BN fold applied to the pooled mean (null: identity)
  float inv_hw;
  const float* w1; const float* b1; const float* w2; const float* b2;   // fc.0 [Ch,C], fc.2 [C,Ch]
  float* zbar; float* hid; float* pre; float* gate;    // [B,C] [B,Ch] [B,C] [B,C]
  int B; int C; int Ch;
  const float* w2t = nullptr;                          // general flavour: packed W2^T [Ch,C] (coalesced), optional
};
int launch_se_fwd(const SeArgs& a, cudaStream_t st);
int launch_se_gen_fwd(const SeArgs& a, cudaStream_t st);   // k_se_gen.cu: sigmoid gate / SiLU hidden, any Ch (hid holds the hidden PRE-activation)
struct SeBwdArgs {
  const float* bwd_stats;    // [B][2][C]: P1 = sum gu, P2 = sum gu*y
  const float* scale; const float* shift;   // gs = scale*P2 + shift*P1 (null: gs = P2)
  float inv_hw;
  const float* w1t; const float* w2t;                  // transposed copies: W1^T [C,Ch], W2^T [Ch,C]
  const float* zbar; const float* hid; const float* pre;
  float* g_pre; float* g_hid; float* g_pool;           // [B,C] [B,Ch] [B,C]
  float* dw1; float* db1; float* dw2; float* db2;      // grads arena
  int B; int C; int Ch;
  const float* w1 = nullptr;                           // general flavour only: W1 [Ch,C] as stored
};
int launch_se_bwd(const SeBwdArgs& a, cudaStream_t st);
int launch_se_gen_bwd(const SeBwdArgs& a, cudaStream_t st);

// ---- k_stem.cu ----
// out_bias != null (inference): y = out_act(conv + out_bias[c]) with BatchNorm folded into w27xC / out_bias
int launch_stem_fwd(const float* img, const float* w27xC, void* y, float* stats, int B, int H, int W, int C,
                    int dtype, cudaStream_t st, const float* out_bias = nullptr, int out_act = TD3D_ACT_NONE);
// dW[C,3,3,3] (reference layout) += sum_pixels gy[p,c] * patch ; gy = alpha*g + beta*y + gamma
int launch_stem_wgrad(const float* img, const void* g, const void* y, const float* alpha, const float* beta,
                      const float* gamma, float* dw, int B, int H, int W, int C, int dtype, cudaStream_t st);

// ---- k_dwconv.cu ----
struct DwArgs {
  const void* x; XForm xf;          // input [B,H,W,C] + lazy transform
  const float* w_taps;              // [k*k][C] fp32
  void* y; float* stats;            // output raw [B,Ho,Wo,C]; stats [B][2][C] (sum, sumsq) or null
  int B, H, W, C, k, stride;
  const float* out_bias = nullptr;  // inference (BatchNorm folded into taps + bias): y = out_act(dw(...) + out_bias[c])
  int out_act = TD3D_ACT_NONE;      //   (column-walker kernel only)
};
int launch_dw_fwd(const DwArgs& a, int dtype, cudaStream_t st);
struct DwBwdArgs {
  const void* g; const void* y_out;                 // gy = alpha[b,c]*g + beta[c]*y_out + gamma[b,c]  [B,Ho,Wo,C]
  const float* alpha; const float* beta; const float* gamma;
  const void* x; XForm xf;                          // forward input (raw) + its transform
  const float* w_taps;
  void* gx;                                         // out: gu_in = (dgrad) * act'(u(x))   [B,H,W,C] (null: skip)
  float* stats;                                     // out: [B][2][C] sum gu_in, sum gu_in*x
  float* dw;                                        // out: grads arena, reference layout [C,1,k,k] (+=)
  int B, H, W, C, k, stride;
};
int launch_dw_bwd(const DwBwdArgs& a, int dtype, cudaStream_t st);      // = launch_dw_bwd_fused after argument checks
int launch_dw_fwd_v2(const DwArgs& a, int dtype, cudaStream_t st);      // k_dw2.cu: tiled persistent kernels
bool dw_walker_supported(int H, int W, int C, int k, int stride);       // k_dww.cu: small planes (W <= 32), stride 1
int launch_dw_fwd_walker(const DwArgs& a, int dtype, cudaStream_t st);
int launch_dw_bwd_fused(const DwBwdArgs& a, int dtype, cudaStream_t st);  // k_dwc.cu: one pass (data + weight gradient + sums)
int launch_dw_fwd_cw(const DwArgs& a, int dtype, cudaStream_t st);        // k_dwc.cu: column walker (optional bias + activation epilogue)

// ---- k_iou.cu (evaluation: 2D-keypoint based 3D IoU, EPnP lift) ----
int launch_iou_2d_based(const float* pred_kp, const float* gt_kp, int n, int portrait, const double* cam_ndc, double* iou,
                        cudaStream_t st);
int launch_lift_2d(const float* kp, int n, int portrait, const double* cam_ndc, double* out, cudaStream_t st);

// ---- k_gemm_simple.cu / k_gemm_tc.cu ----
struct GemmNT {              // Y[M,N] = A[M,K] * W[N,K]^T (+bias[n]) (+addend[m,n])
  const void* a; const void* w; void* y;
  const void* addend;        // T [M,N] or null
  const float* bias;         // [N] or null
  const void* ysaved;        // T [M,N] or null: stats second moment is v*ysaved instead of v*v
  float* stats; int slots;   // [slots][2][N] or null
  int M, N, K;
  int out_f32;               // write Y as float regardless of dtype
  int act;                   // TD3D_ACT_*: y = act(acc + bias) (+ addend)   (inference: BatchNorm folded into w / bias)
};
int launch_gemm_nt_simt(const GemmNT& g, int dtype, cudaStream_t st);
int launch_gemm_nt_tc(const GemmNT& g, cudaStream_t st);       // bf16 only
struct GemmTN {              // C[N1,N2] (f32) = A[M,N1]^T * B[M,N2]
  const void* a; const void* b; float* c; int M, N1, N2;
};
int launch_gemm_tn_simt(const GemmTN& g, int dtype, cudaStream_t st);
int launch_gemm_tn_tc(const GemmTN& g, cudaStream_t st);       // bf16 only
bool tc_gemm_supported(int M, int N, int K);
int tc_timeline_read(unsigned long long* out, int n);   // debug: pipeline event times of CTA 0 (TD3D_TC_DBG & 32)

// ---- k_roi.cu ----
int launch_roi_crop_resize(const uint8_t* frames, int n_frames, int fh, int fw, const int32_t* boxes, int n_boxes, int oh,
                           int ow, const float* mean255, const float* inv_std255, int swap_rb, float* out, cudaStream_t st);

// ---- k_heads.cu ----
struct HeadsArgs {
  const void* feat;                 // T [B,C]
  const int64_t* cats;
  const float* w_reg; int64_t reg_stride;   // head k at w_reg + k*reg_stride: weight [P,C] then bias [P]
  const float* w_cls; const float* b_cls;   // [nc,C], [nc]
  const float* keep; uint64_t seed; int training;
  const int32_t* step_ptr;          // optional device counter mixed into the dropout seed
  float* kp; float* logits;         // [B,P], [B,nc]
  int B, C, P, nc, max_classes;
};
int launch_heads_fwd(const HeadsArgs& a, int dtype, cudaStream_t st);
int launch_heads_all(const HeadsArgs& a, float* kp_all, int dtype, cudaStream_t st);   // export: [max_classes,B,P]
int launch_select_argmax(const float* kp_all, const float* logits, float* kp_sel, int64_t* labels, int B, int P,
                         int nc, int max_classes, cudaStream_t st);
struct HeadsBwdArgs {
  HeadsArgs f;
  const float* d_kp; const float* d_logits;
  float* g_pre;                     // scratch [B,P]
  float* g_feat;                    // out f32 [B,C]
  float* dw_reg; float* dw_cls; float* db_cls;   // grads arena (same strides as params)
  int32_t* present;                 // [max_classes]
};
int launch_heads_bwd(const HeadsBwdArgs& a, int dtype, cudaStream_t st);

// ---- k_loss.cu ----
int launch_loss(const td3d_loss_desc& d, const float* kp, const float* gt, const float* logits, const int64_t* cats,
                int B, int nc, float* loss_out, float* d_kp, float* d_logits, cudaStream_t st);
int launch_metrics(const float* kp, const float* gt, const float* logits, const int64_t* cats, int B, int nc,
                   int max_classes, double* acc, cudaStream_t st);

// ---- k_optim.cu ----
struct OptimArgs {
  td3d_optim_desc d;
  float* p; const float* g; float* s0; float* s1; int64_t n;
  int64_t head_off, head_stride; int n_heads;
  int32_t* steps; const int32_t* present;
  int64_t skip_begin, skip_end;     // floats [skip_begin, skip_end) are left untouched (cls_fc when num_classes == 1: the reference
                                    // never runs it, so its grad stays None and torch optimizers skip it, model_builder.py:140-144)
};
int launch_optim(const OptimArgs& a, cudaStream_t st);
struct PackSeg { const float* src; void* dst; int rows, cols; int transpose; int out_dtype;
                 const float* row_scale; };   // optional [rows]: dst = src[r][c] * row_scale[r] (eval-mode BatchNorm folded into the weights)
struct PackTable { int n; PackSeg seg[256]; };     // EfficientNet-B3: 26 blocks x 7 segments + 3
struct BnFoldSeg { const float *gamma, *beta, *rm, *rv; float *scale, *shift; int C;
                   const float* lin_bias; };     // optional [C]: bias of the Linear in front of this BatchNorm, folded into shift
struct BnFoldTable { int n; BnFoldSeg seg[128]; };
int launch_pack_table(const PackTable& t, cudaStream_t st);
int launch_bn_fold_table(const BnFoldTable& t, float eps, cudaStream_t st);
int launch_cast(const float* src, void* dst, int64_t n, int dtype, cudaStream_t st);
int launch_transpose_cast(const float* src, void* dst, int rows, int cols, int dtype, cudaStream_t st);  // dst[c][r] = src[r][c]
int launch_cast_f32(const void* src, float* dst, int64_t n, int dtype, cudaStream_t st);

}  // namespace td3d
