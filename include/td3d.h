/* td3d -- C ABI of the B200-native (sm_100a) second-stage 3D box regressor of torchdet3d.
 *
 * Drop-in boundary for the hot path of sovrasov/3d-object-detection.pytorch:
 *   model.forward(img, cats) -> (kp, logits)      torchdet3d/builders/model_builder.py:126-146
 *   model.forward_to_onnx(img)                    torchdet3d/builders/model_builder.py:112-124
 *   LossManager.parse_losses(...)                 torchdet3d/losses/regression_losses.py:79-115
 *   loss.backward(); optimizer.step()             torchdet3d/trainer/train.py:50-52
 *   compute_average_distance / compute_accuracy   torchdet3d/evaluation/metrics.py:10-37
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++ / torch types cross this boundary.
 *   - every pointer named d_* / documented "device" is a CUDA device pointer owned by the CALLER
 *     (the PyTorch host code allocates params, grads, optimizer state, workspaces as ordinary
 *     tensors); the library never allocates device memory after td3d_plan_create.
 *   - every call takes an explicit cudaStream_t (passed as void*), is asynchronous, never
 *     synchronises the device, and is CUDA-graph capturable.
 *   - return value: TD3D_OK (0) or a negative TD3D_E* code; td3d_last_error() gives the text.
 *   - there is no CPU fallback: every compute entry point fails with TD3D_ENODEVICE-style CUDA
 *     errors when no sm_100 device is present.
 */
#ifndef TD3D_H_
#define TD3D_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TD3D_ABI_VERSION 2

enum { TD3D_OK = 0, TD3D_EINVAL = -1, TD3D_ECUDA = -2, TD3D_ENOMEM = -3, TD3D_ESTATE = -4 };
enum { TD3D_F32 = 0, TD3D_BF16 = 1 };                       /* activation / packed-weight dtype  */
enum { TD3D_ACT_NONE = 0, TD3D_ACT_RELU = 1, TD3D_ACT_HSWISH = 2, TD3D_ACT_SILU = 3 };
/* backbone family: parameter naming, SE flavour / position, tail */
enum { TD3D_ARCH_MOBILENETV3 = 0,      /* torchdet3d/models/mobilenetv3.py: h_sigmoid/ReLU SE between BN and activation, classifier Linear+BN1d */
       TD3D_ARCH_EFFICIENTNET = 1 };   /* torchvision efficientnet .features (BASELINE configs 3, 5): SiLU, sigmoid/SiLU SE after the activation, no classifier */
enum { TD3D_GEMM_AUTO = 0, TD3D_GEMM_SIMT = 1, TD3D_GEMM_TCGEN05 = 2 };
enum { TD3D_OPT_SGD = 0, TD3D_OPT_ADAMW = 1, TD3D_OPT_RMSPROP = 2, TD3D_OPT_ADADELTA = 3 };

/* One InvertedResidual block (torchdet3d/models/mobilenetv3.py:126-166). */
typedef struct td3d_block_desc {
  int kernel;     /* 3 or 5                                   */
  int stride;     /* 1 or 2                                   */
  int in_ch;      /* block input channels                     */
  int exp_ch;     /* hidden (expanded) channels               */
  int out_ch;     /* block output channels                    */
  int use_se;     /* SELayer present                          */
  int se_hidden;  /* MobileNetV3: _make_divisible(exp_ch // 4, 8); EfficientNet: max(1, in_ch // 4) */
  int use_hs;     /* activation: 0 ReLU, 1 h_swish, 2 SiLU    */
  int name_stage; /* EfficientNet: features.<stage>.<index>.block.* parameter names (MobileNetV3: unused) */
  int name_index;
} td3d_block_desc;

/* Whole regressor: MobileNetV3 backbone (mobilenetv3.py:169-203) + ModelWrapper heads
 * (model_builder.py:76-87). */
typedef struct td3d_net_desc {
  int arch;               /* TD3D_ARCH_*                                          */
  int stem_ch;            /* 16 (MobileNetV3) / 32, 40 (EfficientNet-B0, B3)      */
  int n_blocks;
  const td3d_block_desc* blocks;
  int last_ch;            /* final 1x1 conv width (exp_size: 576 / 960)           */
  int head_ch;            /* classifier width (1024 / 1280); EfficientNet: == last_ch (no classifier) */
  int num_classes;        /* cls_fc outputs                                       */
  int max_classes;        /* number of regressor heads (9)                        */
  int num_points;         /* outputs per head (18)                                */
} td3d_net_desc;

typedef struct td3d_sizes {
  int64_t n_param_tensors;   /* trainable tensors, reference state_dict order                 */
  int64_t param_floats;      /* flat fp32 parameter arena (== gradient arena) length          */
  int64_t n_bn;              /* BatchNorm layers                                              */
  int64_t bn_floats;         /* running_mean|running_var arena: for each BN, mean[C] var[C]   */
  int64_t packed_bytes;      /* compute-layout weight copies (dtype T)                        */
  int64_t workspace_bytes;   /* activations + scratch for the planned batch                   */
  int64_t head_param_offset; /* first float of regressors.0.0.weight in the param arena       */
  int64_t head_param_stride; /* floats per head (num_points*head_ch + num_points)             */
} td3d_sizes;

typedef struct td3d_param_info {
  char name[96];       /* reference state_dict key, e.g. "features.3.conv.7.weight"           */
  int64_t offset;      /* float offset in the parameter / gradient arena                      */
  int64_t numel;
  int32_t ndim;
  int64_t shape[4];    /* reference (OIHW / [out,in]) shape                                    */
  int32_t bn_index;    /* >=0: this tensor is the weight (gamma) or bias (beta) of BN #bn_index */
} td3d_param_info;

typedef struct td3d_bn_info {
  char name[96];       /* key prefix, e.g. "features.3.conv.8"                                */
  int64_t offset;      /* float offset of running_mean in the bn arena; running_var follows   */
  int32_t channels;
} td3d_bn_info;

/* Loss recipe: sum_i coef_i * term_i (loss_builder.py:7-28, regression_losses.py:79-92). */
typedef struct td3d_loss_desc {
  float w_l1, w_smoothl1, w_mse, w_add, w_diag, w_wing, w_ce;   /* 0 => term not used         */
  float smoothl1_beta, wing_w, wing_eps;
} td3d_loss_desc;

typedef struct td3d_optim_desc {
  int kind;                       /* TD3D_OPT_*                                               */
  float lr, weight_decay;
  float momentum; int nesterov;   /* SGD                                                      */
  float beta1, beta2, eps;        /* AdamW (optim_builder.py:11: name 'adam' builds AdamW)    */
  float alpha, rho;               /* RMSprop / Adadelta                                       */
  float grad_scale;               /* multiplies every gradient first (1/world_size for DDP)   */
} td3d_optim_desc;

typedef struct td3d_plan td3d_plan;

/* ---- library ---------------------------------------------------------------------------- */
int         td3d_abi_version(void);
const char* td3d_last_error(void);
int         td3d_device_check(void);     /* TD3D_OK iff current device is sm_100 (B200)          */

/* ---- plan: replaces build_model (model_builder.py:25-71) --------------------------------- */
int  td3d_plan_create(const td3d_net_desc* net, int batch, int height, int width, int dtype,
                      int gemm_impl, td3d_plan** out);
void td3d_plan_destroy(td3d_plan* plan);
int  td3d_plan_sizes(const td3d_plan* plan, td3d_sizes* out);
int  td3d_plan_param_info(const td3d_plan* plan, int64_t index, td3d_param_info* out);
int  td3d_plan_bn_info(const td3d_plan* plan, int64_t index, td3d_bn_info* out);
/* Attach caller-owned device buffers (sizes from td3d_plan_sizes). */
int  td3d_plan_bind(td3d_plan* plan, float* params, float* grads, float* bn_stats,
                    int64_t* bn_num_batches_tracked, void* packed, void* workspace);
/* Optional device i32 counter mixed into the Philox dropout seed (the fused optimizer's global
 * step): a CUDA-graph replay of the train step then draws a fresh cls_fc dropout mask each time. */
int  td3d_plan_set_dropout_counter(td3d_plan* plan, const int32_t* counter);
/* Per-launch profiling (CUDA events on the launching stream, grouped by kernel kind, with the
 * algorithmic bytes each launch must move). Enabling clears previous records. profile_read returns
 * 1 when `kind` is past the last kind. Not for use inside CUDA-graph capture. */
int  td3d_plan_profile(td3d_plan* plan, int enable);
int  td3d_plan_profile_read(td3d_plan* plan, int kind, char* name, int name_cap, double* ms,
                            double* bytes, int64_t* launches);
/* One recorded launch, in issue order: kernel kind, layer tag (forward: 0 stem, i+1 block i,
 * n_blocks+1 tail; backward: the same + 1000), its time and algorithmic bytes. Returns 1 past
 * the last record. */
int  td3d_plan_profile_launch(td3d_plan* plan, int64_t index, int* kind, int* tag, double* ms,
                              double* bytes);
/* params -> compute-layout copies (+ eval-mode BN folding tables); call after any param change */
int  td3d_pack_weights(td3d_plan* plan, void* stream);

/* ---- forward: ModelWrapper.forward (model_builder.py:126-146) ------------------------------
 * img   device f32 [B,3,H,W] (NCHW, as the reference feeds it)
 * cats  device i64 [B] in [0,max_classes)
 * dropout_keep device f32 [B,head_ch] of {0,1} or NULL (NULL in training => Philox mask from seed)
 * kp    device f32 [B,num_points/2,2];  logits device f32 [B,num_classes]                      */
int td3d_forward(td3d_plan* plan, const float* img, const int64_t* cats, const float* dropout_keep,
                 uint64_t seed, int training, float* kp, float* logits, void* stream);
/* forward_to_onnx (model_builder.py:112-124): eval mode, all heads. kp_all f32 [9,B,9,2].
 * If select != 0 additionally applies the deployment consumer (utils/ie_wrappers.py:138-142):
 * label = argmax(logits), kp_sel[b] = kp_all[label_b, b]; labels i64 [B].                      */
int td3d_forward_export(td3d_plan* plan, const float* img, float* kp_all, float* logits,
                        int select, float* kp_sel, int64_t* labels, void* stream);

/* ---- backward of the last training-mode td3d_forward (train.py:51) ------------------------
 * Overwrites the bound gradient arena (== optimizer.zero_grad(); loss.backward()).
 * head_present device i32 [max_classes]: 1 if the head received a gradient (class present in
 * `cats`), 0 => the reference leaves grad=None and optimizers skip the tensor.
 * stage_begin/stage_end select a slice of the backward pass [0, td3d_backward_stages) so the
 * caller can interleave gradient all-reduce buckets; (0,-1) runs everything.                   */
int td3d_backward_stages(const td3d_plan* plan);
int td3d_backward(td3d_plan* plan, const float* d_kp, const float* d_logits, int32_t* head_present,
                  int stage_begin, int stage_end, void* stream);
/* float range [*begin,*end) of the gradient arena that is final once stages [0,stage) ran      */
int td3d_backward_ready_range(const td3d_plan* plan, int stage, int64_t* begin, int64_t* end);

/* ---- loss forward + gradient (regression_losses.py, loss_builder.py) -----------------------
 * loss_out device f32 [8]: total, l1, smoothl1, mse, add, diag, wing, ce (already weighted)
 * d_kp / d_logits may be NULL (forward only).                                                  */
int td3d_loss_fwd_bwd(const td3d_loss_desc* desc, const float* kp, const float* gt_kp,
                      const float* logits, const int64_t* cats, int batch, int num_classes,
                      float* loss_out, float* d_kp, float* d_logits, void* stream);

/* ---- metrics (evaluation/metrics.py:10-68) --------------------------------------------------
 * acc device f64 [4 + 4*max_classes]: sum_ADD(per-sample mean over 9 pts), sum_SADD, hits, count,
 * then per class the same four. Accumulates (+=); caller zeroes.                               */
int td3d_metrics_accum(const float* kp, const float* gt_kp, const float* logits, const int64_t* cats,
                       int batch, int num_classes, int max_classes, double* acc, void* stream);

/* ---- optimizer over the flat arenas (optim_builder.py:5-19) --------------------------------
 * state0/state1 device f32 arenas of param_floats (exp_avg/exp_avg_sq; momentum_buffer/unused;
 * square_avg/unused; square_avg/acc_delta). steps device i32 [1+max_classes]: step counters
 * (global, per head), updated by the call. head_present as produced by td3d_backward or NULL.   */
int td3d_optim_step(td3d_plan* plan, const td3d_optim_desc* desc, float* state0, float* state1,
                    int32_t* steps, const int32_t* head_present, void* stream);

/* ---- ROI front-end: the step before the path ----------------------------------
 * Replaces, per detector box, Regressor.crop + IEModel._preprocess (torchdet3d/utils/ie_wrappers.py:155-158,18-21:
 * frame[y0:y1, x0:x1] -> cv.resize(img, (w, h)) -> transpose(2,0,1)), ConvertColor (utils/transforms.py:10-17) and the
 * test-time Normalize (configs/default_config.py:9-10), bit-exact with OpenCV's uint8 INTER_LINEAR.
 * frames  device u8  [n_frames, frame_h, frame_w, 3]
 * boxes   device i32 [n_boxes, 5]: frame index, x0, y0, x1, y1 (x1 / y1 exclusive; clamped to the frame)
 * mean255 / inv_std255  HOST f32 [3]: mean*255 and 1/(std*255) per output channel
 * out     device f32 [n_boxes, 3, out_h, out_w] -- the `img` argument of td3d_forward / td3d_forward_export       */
int td3d_roi_crop_resize(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* boxes,
                         int n_boxes, int out_h, int out_w, const float* mean255, const float* inv_std255,
                         int swap_rb, float* out, void* stream);

/* ---- evaluation post-processing: the step after the path ------------------------
 * td3d_lift_2d replaces torchdet3d/utils/geometry.py:51-108 (lift_2d: EPnP-style lift of nine normalised 2D keypoints to
 * 3D box points in camera coordinates, up to scale; 16 x 12 system, eigenvector of the smallest eigenvalue).
 * td3d_iou_2d_based replaces the per-sample loop of torchdet3d/evaluation/metrics.py:70-89 (compute_2d_based_iou: lift both
 * keypoint sets in portrait mode, fit boxes and intersect them -- Objectron box.py:123-156,207-225 and iou.py:22-35,74-211
 * vendored under the reference's 3rdparty/); a pair the reference drops on a Qhull / LinAlg error yields 0 here too.
 * kp / pred_kp / gt_kp  device f32 [n, 9, 2];  cam_ndc  HOST f64 {fx, fy, cx, cy} of the NDC camera matrix or NULL for the
 * reference's default (2, 2, 0, 0);  out device f64 [n, 9, 3];  iou device f64 [n] (the caller sums or averages).     */
int td3d_lift_2d(const float* kp, int n, int portrait, const double* cam_ndc, double* out, void* stream);
int td3d_iou_2d_based(const float* pred_kp, const float* gt_kp, int n, int portrait, const double* cam_ndc,
                      double* iou, void* stream);

/* ---- per-kernel entry points (unit tests / profiling) -------------------------------------- */
int td3d_k_stem_fwd(const float* img, const float* w27x16, void* y, float* stats, int B, int H, int W,
                    int C, int dtype, void* stream);
int td3d_k_stem_wgrad(const float* img, const void* g, const void* y, const float* alpha, const float* beta,
                      const float* gamma, float* dw, int B, int H, int W, int C, int dtype, void* stream);
int td3d_k_dw_fwd(const void* x, const float* scale, const float* shift, const float* se, int act,
                  const float* w_taps, void* y, float* stats, int B, int H, int W, int C, int k,
                  int stride, int dtype, void* stream);
int td3d_k_dw_bwd(const void* g, const void* y_out, const float* alpha, const float* beta,
                  const float* gamma, const void* x, const float* scale, const float* shift,
                  const float* se, int act, const float* w_taps, void* gx, float* dw, float* stats,
                  int B, int H, int W, int C, int k, int stride, int dtype, void* stream);
/* The forward with an explicit implementation choice (A/B measurements, kernel tests; the backward has a single
 * implementation, the one-pass column walker):
 *   impl 0 = dispatch as the plan does, 1 = row walker / tiled kernels (k_dww.cu, k_dw2.cu), 2 = column walker (k_dwc.cu);
 *   out_bias / out_act (column walker only): y = out_act(dw(...) + out_bias[c]) -- the inference epilogue with eval-mode
 *   BatchNorm folded into the taps */
int td3d_k_dw_fwd_ex(const void* x, const float* scale, const float* shift, const float* se, int act,
                     const float* w_taps, const float* out_bias, int out_act, void* y, float* stats, int B, int H,
                     int W, int C, int k, int stride, int dtype, int impl, void* stream);
/* Y[M,N] = A[M,K] W[N,K]^T (+ bias[n]) (+ addend[m,n]);  stats (optional): [stat_slots][2][N] += column sums of y and of
 * y * ysaved (y * y when ysaved is null), taken on the stored values.  impl = TD3D_GEMM_TCGEN05: bf16 only, and the statistics
 * need the bf16 output (out_f32 = 0); the slot a sum lands in is unspecified (consumers add all slots). */
int td3d_k_gemm_nt(const void* a, const void* w, void* y, const void* addend, const float* bias,
                   const void* ysaved, float* stats, int stat_slots, int M, int N, int K, int dtype,
                   int out_f32, int impl, void* stream);
int td3d_k_gemm_tn(const void* a, const void* b, float* c, int M, int N1, int N2, int dtype,
                   int impl, void* stream);
int td3d_k_apply_xform(const void* y, const float* scale, const float* shift, const float* se,
                       int act, const void* res, void* out, float* pool_stats, int B, int HW, int C,
                       int dtype, void* stream);
int td3d_k_affine2(const void* g, const void* y, const float* alpha, const float* beta,
                   const float* gamma, void* out, int B, int HW, int C, int dtype, void* stream);
int td3d_k_act_bwd_stats(const void* g, const void* y, const float* scale, const float* shift,
                         const float* se, int act, void* gu, float* stats, int B, int HW, int C,
                         int dtype, void* stream);

/* Debug aid: with env TD3D_TC_DBG=32 CTA 0 of the tcgen05 NT GEMM records %globaltimer (ns) at 8 pipeline
 * events of its first 64 tiles; this copies the [8][64] table to the host. */
int td3d_debug_tc_timeline(uint64_t* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* TD3D_H_ */
