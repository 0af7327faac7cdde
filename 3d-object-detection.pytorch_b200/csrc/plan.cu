// Plan = the native runtime of the regressor hot path: layer table, memory planner for the
// caller-owned arenas, and the forward / backward / optimizer orchestration that walks the network
// launching the sm_100a kernels on one stream (CUDA-graph capturable, no host synchronisation).
//
// Mirrors, as one object, what the reference spreads over
//   build_model            torchdet3d/builders/model_builder.py:25-71
//   MobileNetV3.__init__   torchdet3d/models/mobilenetv3.py:169-197
//   ModelWrapper.forward   torchdet3d/builders/model_builder.py:126-146
//   loss.backward()/step() torchdet3d/trainer/train.py:50-52
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "td3d_kernels.h"

namespace td3d {

// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_last_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return TD3D_ECUDA;
}

static const float BN_EPS = 1e-5f, BN_MOMENTUM = 0.1f;

struct Bn {
  std::string name;
  int C = 0;
  int64_t gamma = 0, beta = 0;          // param arena offsets
  int64_t rm = 0;                       // bn arena offset (running_mean; running_var = rm + C)
  size_t scale = 0, shift = 0, mean = 0, invstd = 0;   // workspace (float) train-mode constants
  size_t escale = 0, eshift = 0;                       // packed arena (float) eval-mode constants
  size_t fstats = 0, bstats = 0;                       // workspace float [B][2][C]
  int fslots = 0, bslots = 0;                          // statistic slots the last producer actually used (0 = all B):
                                                       // GEMM epilogues spread their sums over GS slots only, so the
                                                       // finalize kernels read 8x fewer partial sums for those
};

struct Block {
  td3d_block_desc d;
  bool expand, residual;
  bool se_post;                         // SE after the activation (dw-first MobileNetV3 block; every EfficientNet block)
  int act;                              // TD3D_ACT_* of this block
  int Hin, Win, Hout, Wout;
  int bn1, bn2, bn3;                    // indices into bns (-1 if absent)
  int64_t w1 = -1, wdw = -1, w3 = -1;   // param offsets
  int64_t se_w1 = -1, se_b1 = -1, se_w2 = -1, se_b2 = -1;
  size_t pw1 = 0, pw1t = 0, pdw = 0, pw3 = 0, pw3t = 0;   // packed offsets
  size_t pse_w1t = 0, pse_w2t = 0;                        // fp32 transposed SE weights
  size_t pi1 = 0, pidw = 0, pi3 = 0;                      // inference copies: eval-mode BatchNorm folded into the weights
  size_t y1 = 0, y2 = 0, h = 0, h2 = 0, y3 = 0, out = 0;  // workspace activations (T)
  size_t hstats = 0, hbstats = 0;       // dw-first + SE: pool sums of H and their backward twin
  size_t zbar = 0, hid = 0, pre = 0, gate = 0;
  int64_t first_param = 0;
};

struct Param {
  td3d_param_info info;
};

// Optional per-launch timing (CUDA events on the launching stream) + algorithmic byte counts, by
// kernel kind. Used by bench.py for the roofline of the dominant kernel; off by default.
enum ProfKind { PK_STEM_FWD, PK_STEM_WGRAD, PK_GEMM_FWD, PK_GEMM_DGRAD, PK_GEMM_WGRAD, PK_DW_FWD, PK_DW_BWD,
                PK_BN, PK_XFORM, PK_AFFINE2, PK_ACT_BWD, PK_SE, PK_HEADS, PK_POOL, PK_OPTIM, PK_PACK, PK_COUNT };
static const char* const kProfNames[PK_COUNT] = {
    "stem_fwd", "stem_wgrad", "gemm_fwd", "gemm_dgrad", "gemm_wgrad", "dw_fwd", "dw_bwd", "bn_finalize",
    "apply_xform", "affine2", "act_bwd_stats", "se_fc", "heads", "pool", "optimizer", "pack_weights"};
struct ProfRec { int kind; int tag; double bytes; cudaEvent_t e0, e1; };
struct Prof {
  bool enabled = false;
  int tag = 0;                           // layer tag of the launches being issued (see td3d_plan_profile_launch)
  std::vector<ProfRec> recs;
};

}  // namespace td3d

using namespace td3d;

struct td3d_plan {
  td3d_net_desc net;
  std::vector<td3d_block_desc> block_descs;
  int B, H, W, dtype, gemm_impl;
  size_t esz;                            // bytes per activation element
  std::vector<Param> params;
  std::vector<Bn> bns;
  std::vector<Block> blocks;
  int64_t param_floats = 0, bn_floats = 0;
  size_t packed_bytes = 0, ws_bytes = 0;
  // stem / tail
  int H1, W1;                            // stem output resolution
  int stem_act, se_kind;                 // stem / last-conv activation; 0 = h_sigmoid/ReLU SE (k_se.cu), 1 = sigmoid/SiLU SE (k_se_gen.cu)
  bool has_fc;                           // classifier Linear + BatchNorm1d (MobileNetV3 only, model_builder.py:117-118)
  int bn_stem, bn_last, bn_fc;
  int64_t w_stem, w_last, w_fc, b_fc, w_reg0, w_cls, b_cls;
  int64_t head_stride;
  size_t p_stem, p_last, p_lastt, p_fc, p_fct;
  size_t pi_stem, pi_last, pi_fc;        // inference copies (BatchNorm folded)
  size_t y0, x0, yc, pooled, yfc, feat, kp_saved, logits_saved, pool_stats, pool_bstats_unused;
  int Hl, Wl;                            // final resolution
  int64_t first_param_tail;
  // backward scratch
  size_t g_narrow[2], g_y3, g_wide_a, g_wide_b, g_feat, g_pool_f, g_pre_heads;
  size_t alpha, beta, gammac, zeros_c, se_gpre, se_ghid, se_gpool, se_gpool_scaled;
  size_t fstats_begin = 0, fstats_end = 0, bstats_begin = 0, bstats_end = 0;
  // bound buffers
  float* P = nullptr; float* G = nullptr; float* BNB = nullptr; int64_t* NBT = nullptr;
  uint8_t* PK = nullptr; uint8_t* WS = nullptr;
  // last forward
  const float* last_img = nullptr; const int64_t* last_cats = nullptr; const float* last_keep = nullptr;
  uint64_t last_seed = 0; int last_training = 0;
  const int32_t* dropout_counter = nullptr;
  Prof prof;
  // side stream of backward (created at bind time, outside any graph capture): runs the weight-gradient GEMMs
  // concurrently with the data-gradient chain on the caller's stream
  cudaStream_t side[1] = {nullptr};
  cudaEvent_t ev_fork[1] = {nullptr}, ev_join[1] = {nullptr};
  bool side_pending[1] = {false};
  int overlap = 1;
  td3d::PackTable pack_table;
  td3d::PackTable pack_table_eval;       // weights * eval-mode BatchNorm scale (rebuilt with the fold, td3d_pack_weights)
  td3d::BnFoldTable fold_table;
};

namespace td3d {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Bump {
  size_t off = 0;
  size_t take(size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

static int64_t add_param(td3d_plan* pl, const std::string& name, std::vector<int64_t> shape, int bn_index = -1) {
  Param p;
  memset(&p.info, 0, sizeof(p.info));
  snprintf(p.info.name, sizeof(p.info.name), "%s", name.c_str());
  int64_t n = 1;
  for (size_t i = 0; i < shape.size(); ++i) { p.info.shape[i] = shape[i]; n *= shape[i]; }
  p.info.ndim = (int)shape.size();
  p.info.numel = n;
  p.info.offset = pl->param_floats;
  p.info.bn_index = bn_index;
  pl->param_floats = (int64_t)align_up((size_t)(pl->param_floats + n), 4);
  pl->params.push_back(p);
  return p.info.offset;
}

static int add_bn(td3d_plan* pl, const std::string& prefix, int C) {
  Bn b;
  b.name = prefix;
  b.C = C;
  int idx = (int)pl->bns.size();
  b.gamma = add_param(pl, prefix + ".weight", {C}, idx);
  b.beta = add_param(pl, prefix + ".bias", {C}, idx);
  b.rm = pl->bn_floats;
  pl->bn_floats += 2 * (int64_t)C;
  pl->bns.push_back(b);
  return idx;
}

static int build(td3d_plan* pl) {
  const td3d_net_desc& n = pl->net;
  const int B = pl->B;
  pl->H1 = (pl->H - 1) / 2 + 1;
  pl->W1 = (pl->W - 1) / 2 + 1;
  // ---- parameter table in reference state_dict order ----
  const bool eff = n.arch == TD3D_ARCH_EFFICIENTNET;
  pl->stem_act = eff ? TD3D_ACT_SILU : TD3D_ACT_HSWISH;
  pl->se_kind = eff ? 1 : 0;
  pl->has_fc = !eff;
  pl->w_stem = add_param(pl, "features.0.0.weight", {n.stem_ch, 3, 3, 3});
  pl->bn_stem = add_bn(pl, "features.0.1", n.stem_ch);
  int h = pl->H1, w = pl->W1;
  int cin = n.stem_ch;
  int last_stage = 0;
  for (int i = 0; i < n.n_blocks; ++i) {
    Block b;
    b.d = pl->block_descs[i];
    TD3D_REQUIRE(b.d.in_ch == cin, "block %d: in_ch=%d does not chain from %d", i, b.d.in_ch, cin);
    TD3D_REQUIRE(b.d.in_ch % 8 == 0 && b.d.exp_ch % 8 == 0 && b.d.out_ch % 8 == 0, "block %d: channels must be multiples of 8", i);
    TD3D_REQUIRE((b.d.kernel == 3 || b.d.kernel == 5) && (b.d.stride == 1 || b.d.stride == 2), "block %d: bad kernel/stride", i);
    TD3D_REQUIRE(b.d.use_hs >= 0 && b.d.use_hs <= 2, "block %d: bad activation code %d", i, b.d.use_hs);
    b.expand = b.d.in_ch != b.d.exp_ch;
    b.residual = b.d.stride == 1 && b.d.in_ch == b.d.out_ch;
    b.act = b.d.use_hs == 2 ? TD3D_ACT_SILU : (b.d.use_hs == 1 ? TD3D_ACT_HSWISH : TD3D_ACT_RELU);
    b.se_post = b.d.use_se && (eff || !b.expand);
    b.Hin = h; b.Win = w;
    b.Hout = (h - 1) / b.d.stride + 1; b.Wout = (w - 1) / b.d.stride + 1;
    b.first_param = pl->param_floats;
    b.bn1 = -1;
    if (!eff) {
      // torchdet3d/models/mobilenetv3.py:133-160: features.<i+1>.conv.<j>
      std::string pre = "features." + std::to_string(i + 1) + ".conv.";
      int j = 0;
      if (b.expand) {
        b.w1 = add_param(pl, pre + "0.weight", {b.d.exp_ch, b.d.in_ch, 1, 1});
        b.bn1 = add_bn(pl, pre + "1", b.d.exp_ch);
        j = 3;
      }
      b.wdw = add_param(pl, pre + std::to_string(j) + ".weight", {b.d.exp_ch, 1, b.d.kernel, b.d.kernel});
      b.bn2 = add_bn(pl, pre + std::to_string(j + 1), b.d.exp_ch);
      int se_idx = b.expand ? j + 2 : j + 3, pw_idx = b.expand ? 7 : 4;
      if (b.d.use_se) {
        std::string sp = pre + std::to_string(se_idx) + ".fc.";
        b.se_w1 = add_param(pl, sp + "0.weight", {b.d.se_hidden, b.d.exp_ch});
        b.se_b1 = add_param(pl, sp + "0.bias", {b.d.se_hidden});
        b.se_w2 = add_param(pl, sp + "2.weight", {b.d.exp_ch, b.d.se_hidden});
        b.se_b2 = add_param(pl, sp + "2.bias", {b.d.exp_ch});
      }
      b.w3 = add_param(pl, pre + std::to_string(pw_idx) + ".weight", {b.d.out_ch, b.d.exp_ch, 1, 1});
      b.bn3 = add_bn(pl, pre + std::to_string(pw_idx + 1), b.d.out_ch);
    } else {
      // torchvision MBConv: features.<stage>.<index>.block.<k> = [expand Conv2dNormActivation], depthwise
      // Conv2dNormActivation, SqueezeExcitation (fc1 / fc2 are 1x1 convs), project Conv2dNormActivation
      TD3D_REQUIRE(b.d.use_se && b.d.se_hidden >= 1, "block %d: EfficientNet blocks carry an SE layer", i);
      std::string pre = "features." + std::to_string(b.d.name_stage) + "." + std::to_string(b.d.name_index) + ".block.";
      int k = 0;
      if (b.expand) {
        b.w1 = add_param(pl, pre + "0.0.weight", {b.d.exp_ch, b.d.in_ch, 1, 1});
        b.bn1 = add_bn(pl, pre + "0.1", b.d.exp_ch);
        k = 1;
      }
      b.wdw = add_param(pl, pre + std::to_string(k) + ".0.weight", {b.d.exp_ch, 1, b.d.kernel, b.d.kernel});
      b.bn2 = add_bn(pl, pre + std::to_string(k) + ".1", b.d.exp_ch);
      std::string sp = pre + std::to_string(k + 1) + ".";
      b.se_w1 = add_param(pl, sp + "fc1.weight", {b.d.se_hidden, b.d.exp_ch, 1, 1});
      b.se_b1 = add_param(pl, sp + "fc1.bias", {b.d.se_hidden});
      b.se_w2 = add_param(pl, sp + "fc2.weight", {b.d.exp_ch, b.d.se_hidden, 1, 1});
      b.se_b2 = add_param(pl, sp + "fc2.bias", {b.d.exp_ch});
      b.w3 = add_param(pl, pre + std::to_string(k + 2) + ".0.weight", {b.d.out_ch, b.d.exp_ch, 1, 1});
      b.bn3 = add_bn(pl, pre + std::to_string(k + 2) + ".1", b.d.out_ch);
      if (b.d.name_stage > last_stage) last_stage = b.d.name_stage;
    }
    pl->blocks.push_back(b);
    h = b.Hout; w = b.Wout; cin = b.d.out_ch;
  }
  pl->Hl = h; pl->Wl = w;
  pl->first_param_tail = pl->param_floats;
  pl->w_fc = pl->b_fc = -1; pl->bn_fc = -1;
  if (!eff) {
    pl->w_last = add_param(pl, "conv.0.weight", {n.last_ch, cin, 1, 1});
    pl->bn_last = add_bn(pl, "conv.1", n.last_ch);
    pl->w_fc = add_param(pl, "classifier.0.weight", {n.head_ch, n.last_ch});
    pl->b_fc = add_param(pl, "classifier.0.bias", {n.head_ch});
    pl->bn_fc = add_bn(pl, "classifier.1", n.head_ch);
  } else {
    TD3D_REQUIRE(n.head_ch == n.last_ch, "EfficientNet: the heads read the pooled last conv (head_ch %d != last_ch %d)", n.head_ch, n.last_ch);
    std::string pre = "features." + std::to_string(last_stage + 1) + ".";
    pl->w_last = add_param(pl, pre + "0.weight", {n.last_ch, cin, 1, 1});
    pl->bn_last = add_bn(pl, pre + "1", n.last_ch);
  }
  pl->head_stride = 0;
  for (int k = 0; k < n.max_classes; ++k) {
    int64_t wo = add_param(pl, "regressors." + std::to_string(k) + ".0.weight", {n.num_points, n.head_ch});
    add_param(pl, "regressors." + std::to_string(k) + ".0.bias", {n.num_points});
    if (k == 0) pl->w_reg0 = wo;
    if (k == 1) pl->head_stride = wo - pl->w_reg0;
  }
  if (n.max_classes == 1) pl->head_stride = pl->param_floats - pl->w_reg0;
  pl->w_cls = add_param(pl, "cls_fc.1.weight", {n.num_classes, n.head_ch});
  pl->b_cls = add_param(pl, "cls_fc.1.bias", {n.num_classes});

  // ---- packed weights ----
  Bump pk;
  const size_t e = pl->esz;
  pl->p_stem = pk.take(sizeof(float) * 27 * n.stem_ch);
  for (auto& b : pl->blocks) {
    if (b.expand) {
      b.pw1 = pk.take(e * b.d.exp_ch * b.d.in_ch);
      b.pw1t = pk.take(e * b.d.exp_ch * b.d.in_ch);
    }
    b.pdw = pk.take(sizeof(float) * b.d.kernel * b.d.kernel * b.d.exp_ch);
    b.pw3 = pk.take(e * b.d.out_ch * b.d.exp_ch);
    b.pw3t = pk.take(e * b.d.out_ch * b.d.exp_ch);
    if (b.d.use_se) {
      b.pse_w1t = pk.take(sizeof(float) * b.d.exp_ch * b.d.se_hidden);
      b.pse_w2t = pk.take(sizeof(float) * b.d.exp_ch * b.d.se_hidden);
    }
  }
  pl->p_last = pk.take(e * n.last_ch * cin);
  pl->p_lastt = pk.take(e * n.last_ch * cin);
  pl->p_fc = pl->p_fct = 0;
  if (pl->has_fc) {
    pl->p_fc = pk.take(e * n.head_ch * n.last_ch);
    pl->p_fct = pk.take(e * n.head_ch * n.last_ch);
  }
  for (auto& bn : pl->bns) {
    bn.escale = pk.take(sizeof(float) * bn.C);
    bn.eshift = pk.take(sizeof(float) * bn.C);
  }
  pl->pi_stem = pk.take(sizeof(float) * 27 * n.stem_ch);
  for (auto& b : pl->blocks) {
    if (b.expand) b.pi1 = pk.take(e * b.d.exp_ch * b.d.in_ch);
    b.pidw = pk.take(sizeof(float) * b.d.kernel * b.d.kernel * b.d.exp_ch);
    b.pi3 = pk.take(e * b.d.out_ch * b.d.exp_ch);
  }
  pl->pi_last = pk.take(e * n.last_ch * cin);
  pl->pi_fc = pl->has_fc ? pk.take(e * n.head_ch * n.last_ch) : 0;
  pl->packed_bytes = pk.off;

  // ---- workspace ----
  Bump ws;
  auto act = [&](int64_t rows, int C) { return ws.take(e * (size_t)rows * C); };
  auto f32 = [&](int64_t nelem) { return ws.take(sizeof(float) * (size_t)nelem); };
  // statistics first (two contiguous regions so one memset clears each)
  pl->fstats_begin = ws.off;
  for (auto& bn : pl->bns) bn.fstats = f32((int64_t)B * 2 * bn.C);
  for (auto& b : pl->blocks)
    if (b.se_post) b.hstats = f32((int64_t)B * 2 * b.d.exp_ch);
  pl->pool_stats = f32((int64_t)B * 2 * n.last_ch);
  pl->fstats_end = ws.off;
  pl->bstats_begin = ws.off;
  for (auto& bn : pl->bns) bn.bstats = f32((int64_t)B * 2 * bn.C);
  for (auto& b : pl->blocks)
    if (b.se_post) b.hbstats = f32((int64_t)B * 2 * b.d.exp_ch);
  pl->bstats_end = ws.off;
  for (auto& bn : pl->bns) {
    bn.scale = f32(bn.C); bn.shift = f32(bn.C); bn.mean = f32(bn.C); bn.invstd = f32(bn.C);
  }
  const int64_t M1 = (int64_t)B * pl->H1 * pl->W1;
  pl->y0 = act(M1, n.stem_ch);
  pl->x0 = act(M1, n.stem_ch);
  int64_t max_narrow = M1 * n.stem_ch, max_wide = 0;
  int maxC = n.stem_ch;
  for (auto& b : pl->blocks) {
    const int64_t Mi = (int64_t)B * b.Hin * b.Win, Mo = (int64_t)B * b.Hout * b.Wout;
    if (b.expand) b.y1 = act(Mi, b.d.exp_ch);
    b.y2 = act(Mo, b.d.exp_ch);
    b.h2 = act(Mo, b.d.exp_ch);
    b.y3 = act(Mo, b.d.out_ch);
    b.out = act(Mo, b.d.out_ch);
    if (b.d.use_se) {
      b.zbar = f32((int64_t)B * b.d.exp_ch); b.hid = f32((int64_t)B * b.d.se_hidden);
      b.pre = f32((int64_t)B * b.d.exp_ch); b.gate = f32((int64_t)B * b.d.exp_ch);
    }
    if (Mi * b.d.in_ch > max_narrow) max_narrow = Mi * b.d.in_ch;
    if (Mo * b.d.out_ch > max_narrow) max_narrow = Mo * b.d.out_ch;
    if (Mi * b.d.exp_ch > max_wide) max_wide = Mi * b.d.exp_ch;
    if (b.d.exp_ch > maxC) maxC = b.d.exp_ch;
  }
  const int64_t Ml = (int64_t)B * pl->Hl * pl->Wl;
  pl->yc = act(Ml, n.last_ch);
  pl->pooled = act(B, n.last_ch);
  pl->yfc = act(B, n.head_ch);
  pl->feat = pl->has_fc ? act(B, n.head_ch) : pl->pooled;     // no classifier: the heads read the pooled last conv
  pl->kp_saved = f32((int64_t)B * n.num_points);
  pl->logits_saved = f32((int64_t)B * n.num_classes);
  if (Ml * n.last_ch > max_wide) max_wide = Ml * n.last_ch;
  if ((int64_t)B * n.head_ch > max_wide) max_wide = (int64_t)B * n.head_ch;
  if (n.last_ch > maxC) maxC = n.last_ch;
  if (n.head_ch > maxC) maxC = n.head_ch;
  pl->g_narrow[0] = act(max_narrow, 1);
  pl->g_narrow[1] = act(max_narrow, 1);
  pl->g_y3 = act(max_narrow, 1);
  pl->g_wide_a = act(max_wide, 1);
  pl->g_wide_b = act(max_wide, 1);
  pl->g_feat = f32((int64_t)B * n.head_ch);
  pl->g_pool_f = f32((int64_t)B * n.last_ch);
  pl->g_pre_heads = f32((int64_t)B * n.num_points);
  pl->alpha = f32((int64_t)B * maxC);
  pl->gammac = f32((int64_t)B * maxC);
  pl->beta = f32(maxC);
  pl->zeros_c = f32(maxC);
  pl->se_gpre = f32((int64_t)B * maxC);
  pl->se_ghid = f32((int64_t)B * maxC);
  pl->se_gpool = f32((int64_t)B * maxC);
  pl->se_gpool_scaled = f32((int64_t)B * maxC);
  pl->ws_bytes = ws.off;
  return TD3D_OK;
}

// ---- helpers --------------------------------------------------------------------------------
struct Ctx {
  td3d_plan* pl;
  cudaStream_t st;
  template <typename T = void> T* ws(size_t off) const { return reinterpret_cast<T*>(pl->WS + off); }
  float* wsf(size_t off) const { return reinterpret_cast<float*>(pl->WS + off); }
  void* pk(size_t off) const { return pl->PK + off; }
  float* pkf(size_t off) const { return reinterpret_cast<float*>(pl->PK + off); }
  float* P(int64_t off) const { return pl->P + off; }
  float* G(int64_t off) const { return pl->G + off; }
  void pb(int kind, double bytes) const {
    if (!pl->prof.enabled) return;
    ProfRec r; r.kind = kind; r.bytes = bytes; r.tag = pl->prof.tag;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    pdl_break(st);                             // timed launches are plain launches
    cudaEventRecord(r.e0, st);
    pl->prof.recs.push_back(r);
  }
  void pe() const {
    if (!pl->prof.enabled) return;
    pdl_break(st);
    cudaEventRecord(pl->prof.recs.back().e1, st);
  }
  double esz() const { return (double)pl->esz; }
};

// ---- backward side streams ---------------------------------------------------------------------
// fork: work launched on the returned context starts after everything issued so far on c.st.
// join: c.st waits for everything issued on that side stream.  Both are plain event record / wait
// pairs, so a CUDA-graph capture of the caller's stream records them as graph edges.
static bool side_on(const Ctx& c, int which) {
  return c.pl->overlap && !c.pl->prof.enabled && c.pl->side[which] != nullptr;
}
static int side_fork(const Ctx& c, int which, Ctx* out) {
  *out = c;
  if (!side_on(c, which)) return TD3D_OK;
  td3d_plan* pl = c.pl;
  pdl_break(c.st);                             // the launches on either side of an event record / wait are plain launches
  pdl_break(pl->side[which]);
  TD3D_CUDA(cudaEventRecord(pl->ev_fork[which], c.st));
  TD3D_CUDA(cudaStreamWaitEvent(pl->side[which], pl->ev_fork[which], 0));
  out->st = pl->side[which];
  return TD3D_OK;
}
static int side_done(const Ctx& c, int which) {        // call after the side launches of one fork
  if (!side_on(c, which)) return TD3D_OK;
  pdl_break(c.pl->side[which]);
  TD3D_CUDA(cudaEventRecord(c.pl->ev_join[which], c.pl->side[which]));
  c.pl->side_pending[which] = true;
  return TD3D_OK;
}
static int side_join(const Ctx& c, int which) {
  if (!c.pl->side_pending[which]) return TD3D_OK;
  pdl_break(c.st);
  TD3D_CUDA(cudaStreamWaitEvent(c.st, c.pl->ev_join[which], 0));
  c.pl->side_pending[which] = false;
  return TD3D_OK;
}

// run a launcher under the profiler: K = kind, BYTES = algorithmic bytes this launch must move
#define TD3D_K(K, BYTES, expr)        \
  do {                                \
    c.pb((K), (double)(BYTES));       \
    int _rc = (expr);                 \
    c.pe();                           \
    if (_rc != TD3D_OK) return _rc;   \
  } while (0)


// ---- profiled launch wrappers (algorithmic bytes = what the op must move at minimum) -----------
static int p_xform(const Ctx& c, const void* y, const XForm& xf, const void* res, void* out, float* pool, int B, int HW,
                   int C, int dt, cudaStream_t st) {
  double n = (double)B * HW * C;
  TD3D_K(out ? PK_XFORM : PK_POOL, n * c.esz() * (1 + (res ? 1 : 0) + (out ? 1 : 0)),
         launch_apply_xform(y, xf, res, out, pool, B, HW, C, dt, st));
  return TD3D_OK;
}
static int p_affine2(const Ctx& c, const void* g, const void* y, const float* al, const float* be, const float* ga,
                     void* out, int B, int HW, int C, int dt, cudaStream_t st) {
  TD3D_K(PK_AFFINE2, 3.0 * B * HW * C * c.esz(), launch_affine2(g, y, al, be, ga, out, B, HW, C, dt, st));
  return TD3D_OK;
}
static int p_actbwd(const Ctx& c, const void* g, const float* gp, float gs, const void* y, const XForm& xf, void* gu,
                    float* stats, int B, int HW, int C, int dt, cudaStream_t st, const void* addend = nullptr) {
  double n = (double)B * HW * C;
  TD3D_K(PK_ACT_BWD, n * c.esz() * (2 + (g ? 1 : 0) + (addend ? 1 : 0)),
         launch_act_bwd_stats(g, gp, gs, y, xf, gu, stats, B, HW, C, dt, st, addend));
  return TD3D_OK;
}
static int p_dwf(const Ctx& c, const DwArgs& a, int dt, cudaStream_t st) {
  int Ho = (a.H - 1) / a.stride + 1, Wo = (a.W - 1) / a.stride + 1;
  double in = (double)a.B * a.H * a.W * a.C, out = (double)a.B * Ho * Wo * a.C;
  TD3D_K(PK_DW_FWD, (in + out) * c.esz(), launch_dw_fwd(a, dt, st));
  return TD3D_OK;
}
static int p_dwb(const Ctx& c, const DwBwdArgs& a, int dt, cudaStream_t st) {
  int Ho = (a.H - 1) / a.stride + 1, Wo = (a.W - 1) / a.stride + 1;
  double in = (double)a.B * a.H * a.W * a.C, out = (double)a.B * Ho * Wo * a.C;
  // algorithmic bytes: read g, y_out, x once; write gx once (what the one-pass kernel does)
  TD3D_K(PK_DW_BWD, (2 * out + 2 * in) * c.esz(), launch_dw_bwd(a, dt, st));
  return TD3D_OK;
}

static bool use_tc(const td3d_plan* pl, int M, int N, int K) {
  if (pl->dtype != TD3D_BF16) return false;
  if (pl->gemm_impl == TD3D_GEMM_SIMT) return false;
  return tc_gemm_supported(M, N, K);
}

static int gemm_nt(const Ctx& c, GemmNT g, int kind = PK_GEMM_FWD) {
  double e = c.esz();
  double bytes = ((double)g.M * g.K + (double)g.N * g.K) * e + (double)g.M * g.N * (g.out_f32 ? 4.0 : e) +
                 (double)g.M * g.N * e * ((g.addend ? 1 : 0) + (g.ysaved ? 1 : 0));
  if (use_tc(c.pl, g.M, g.N, g.K)) { TD3D_K(kind, bytes, launch_gemm_nt_tc(g, c.st)); }
  else { TD3D_K(kind, bytes, launch_gemm_nt_simt(g, c.pl->dtype, c.st)); }
  return TD3D_OK;
}
static int gemm_tn_main(const Ctx& c, GemmTN g);
// weight-gradient GEMM on side stream 0; the caller joins (side_join(c, 0)) before the next kernel on
// c.st that overwrites an operand it reads
static int gemm_tn(const Ctx& c, GemmTN g) {
  Ctx sc;
  TD3D_TRY(side_fork(c, 0, &sc));
  TD3D_TRY(gemm_tn_main(sc, g));
  if (sc.st != c.st) TD3D_TRY(side_done(c, 0));
  return TD3D_OK;
}
static int gemm_tn_main(const Ctx& c, GemmTN g) {
  double bytes = ((double)g.M * g.N1 + (double)g.M * g.N2) * c.esz() + 4.0 * g.N1 * g.N2;
  if (c.pl->dtype == TD3D_BF16 && c.pl->gemm_impl != TD3D_GEMM_SIMT && g.N1 % 8 == 0 && g.N2 % 8 == 0) {
    TD3D_K(PK_GEMM_WGRAD, bytes, launch_gemm_tn_tc(g, c.st));
  } else {
    TD3D_K(PK_GEMM_WGRAD, bytes, launch_gemm_tn_simt(g, c.pl->dtype, c.st));
  }
  return TD3D_OK;
}

// BatchNorm after a conv: training -> finalize batch statistics (and update running stats);
// eval -> folded running statistics. Returns the (scale, shift) the consumer must apply.
static int bn_forward(const Ctx& c, int idx, double count, int training, const float** scale, const float** shift) {
  Bn& bn = c.pl->bns[idx];
  if (training) {
    BnFwdArgs a;
    a.stats = c.wsf(bn.fstats); a.slots = bn.fslots ? bn.fslots : c.pl->B; a.count = count;
    a.gamma = c.P(bn.gamma); a.beta = c.P(bn.beta);
    a.running_mean = c.pl->BNB + bn.rm; a.running_var = c.pl->BNB + bn.rm + bn.C;
    a.nbt = c.pl->NBT ? c.pl->NBT + idx : nullptr;
    a.scale = c.wsf(bn.scale); a.shift = c.wsf(bn.shift); a.mean = c.wsf(bn.mean); a.invstd = c.wsf(bn.invstd);
    a.C = bn.C; a.momentum = BN_MOMENTUM; a.eps = BN_EPS;
    TD3D_K(PK_BN, 8.0 * c.pl->B * bn.C, launch_bn_finalize_fwd(a, c.st));
    *scale = a.scale; *shift = a.shift;
  } else {
    *scale = c.pkf(bn.escale); *shift = c.pkf(bn.eshift);
  }
  return TD3D_OK;
}

static int se_forward(const Ctx& c, const SeArgs& s) {
  return c.pl->se_kind ? launch_se_gen_fwd(s, c.st) : launch_se_fwd(s, c.st);
}
static int se_backward(const Ctx& c, const SeBwdArgs& s) {
  return c.pl->se_kind ? launch_se_gen_bwd(s, c.st) : launch_se_bwd(s, c.st);
}

static XForm xf_make(const float* scale, const float* shift, const float* se, int act, int se_post = 0) {
  XForm x; x.scale = scale; x.shift = shift; x.se = se; x.act = act; x.se_post = se_post;
  return x;
}

static int forward_backbone(const Ctx& c, const float* img, int training, const void** feat_out) {
  td3d_plan* pl = c.pl;
  const td3d_net_desc& n = pl->net;
  const int B = pl->B, dt = pl->dtype;
  const int GS = B < 32 ? B : 32;      // statistic slots of the GEMM epilogues (tile m_tile lands in slot m_tile % GS)
  pdl_break(c.st);
  TD3D_CUDA(cudaMemsetAsync(pl->WS + pl->fstats_begin, 0, pl->fstats_end - pl->fstats_begin, c.st));
  const float *sc, *sh;
  pl->prof.tag = 0;
  // stem
  Bn& b0 = pl->bns[pl->bn_stem];
  TD3D_K(PK_STEM_FWD, (double)B * 3 * pl->H * pl->W * 4 + (double)B * pl->H1 * pl->W1 * n.stem_ch * c.esz(), launch_stem_fwd(img, c.pkf(pl->p_stem), c.ws(pl->y0), c.wsf(b0.fstats), B, pl->H, pl->W, n.stem_ch, dt, c.st));
  TD3D_TRY(bn_forward(c, pl->bn_stem, (double)B * pl->H1 * pl->W1, training, &sc, &sh));
  TD3D_TRY(p_xform(c, c.ws(pl->y0), xf_make(sc, sh, nullptr, pl->stem_act), nullptr, c.ws(pl->x0), nullptr,
                              B, pl->H1 * pl->W1, n.stem_ch, dt, c.st));
  const void* cur = c.ws(pl->x0);
  for (auto& b : pl->blocks) {
    pl->prof.tag = (int)(&b - pl->blocks.data()) + 1;
    const int act = b.act;
    const int HWi = b.Hin * b.Win, HWo = b.Hout * b.Wout;
    const int Mi = B * HWi, Mo = B * HWo;
    const int E = b.d.exp_ch;
    DwArgs dw;
    if (b.expand) {
      GemmNT g = {};
      g.a = cur; g.w = c.pk(b.pw1); g.y = c.ws(b.y1);
      g.stats = c.wsf(pl->bns[b.bn1].fstats); g.slots = pl->bns[b.bn1].fslots = GS;
      g.M = Mi; g.N = E; g.K = b.d.in_ch;
      TD3D_TRY(gemm_nt(c, g));
      TD3D_TRY(bn_forward(c, b.bn1, (double)Mi, training, &sc, &sh));
      dw.x = c.ws(b.y1); dw.xf = xf_make(sc, sh, nullptr, act);
    } else {
      dw.x = cur; dw.xf = xf_make(nullptr, nullptr, nullptr, TD3D_ACT_NONE);
    }
    dw.w_taps = c.pkf(b.pdw); dw.y = c.ws(b.y2); dw.stats = c.wsf(pl->bns[b.bn2].fstats);
    dw.B = B; dw.H = b.Hin; dw.W = b.Win; dw.C = E; dw.k = b.d.kernel; dw.stride = b.d.stride;
    TD3D_TRY(p_dwf(c, dw, dt, c.st));
    TD3D_TRY(bn_forward(c, b.bn2, (double)Mo, training, &sc, &sh));
    if (b.d.use_se) {
      SeArgs s;
      s.w1 = c.P(b.se_w1); s.b1 = c.P(b.se_b1); s.w2 = c.P(b.se_w2); s.b2 = c.P(b.se_b2);
      s.zbar = c.wsf(b.zbar); s.hid = c.wsf(b.hid); s.pre = c.wsf(b.pre); s.gate = c.wsf(b.gate);
      s.B = B; s.C = E; s.Ch = b.d.se_hidden; s.inv_hw = 1.f / (float)HWo; s.w2t = c.pkf(b.pse_w2t);
      if (!b.se_post) {    // BN -> SE -> act (mobilenetv3.py:153-156): squeeze from the dw epilogue sums
        s.pool_stats = c.wsf(pl->bns[b.bn2].fstats); s.scale = sc; s.shift = sh;
        TD3D_K(PK_SE, 8.0 * b.d.exp_ch * b.d.se_hidden, se_forward(c, s));
        TD3D_TRY(p_xform(c, c.ws(b.y2), xf_make(sc, sh, s.gate, act), nullptr, c.ws(b.h2), nullptr, B, HWo, E, dt, c.st));
      } else {             // BN -> act -> SE (mobilenetv3.py:137-140; every torchvision MBConv)
        // squeeze = mean of act(BN(y2)): a read-only pooling pass (act(BN(y2)) itself is never materialised), then ONE
        // apply pass h2 = act(BN(y2)) * gate
        TD3D_TRY(p_xform(c, c.ws(b.y2), xf_make(sc, sh, nullptr, act), nullptr, nullptr, c.wsf(b.hstats), B, HWo, E, dt, c.st));
        s.pool_stats = c.wsf(b.hstats); s.scale = nullptr; s.shift = nullptr;
        TD3D_K(PK_SE, 8.0 * b.d.exp_ch * b.d.se_hidden, se_forward(c, s));
        TD3D_TRY(p_xform(c, c.ws(b.y2), xf_make(sc, sh, s.gate, act, 1), nullptr, c.ws(b.h2), nullptr, B, HWo, E, dt, c.st));
      }
    } else {
      TD3D_TRY(p_xform(c, c.ws(b.y2), xf_make(sc, sh, nullptr, act), nullptr, c.ws(b.h2), nullptr, B, HWo, E, dt, c.st));
    }
    GemmNT g = {};
    g.a = c.ws(b.h2); g.w = c.pk(b.pw3); g.y = c.ws(b.y3);
    g.stats = c.wsf(pl->bns[b.bn3].fstats); g.slots = pl->bns[b.bn3].fslots = GS;
    g.M = Mo; g.N = b.d.out_ch; g.K = E;
    TD3D_TRY(gemm_nt(c, g));
    TD3D_TRY(bn_forward(c, b.bn3, (double)Mo, training, &sc, &sh));
    TD3D_TRY(p_xform(c, c.ws(b.y3), xf_make(sc, sh, nullptr, TD3D_ACT_NONE), b.residual ? cur : nullptr,
                                c.ws(b.out), nullptr, B, HWo, b.d.out_ch, dt, c.st));
    cur = c.ws(b.out);
  }
  // final 1x1 conv + BN + h_swish + global average pool (mobilenetv3.py:188,199-203; model_builder.py:98)
  pl->prof.tag = (int)pl->blocks.size() + 1;
  const int HWl = pl->Hl * pl->Wl, Ml = B * HWl;
  const int Cl = pl->blocks.back().d.out_ch;
  {
    GemmNT g = {};
    g.a = cur; g.w = c.pk(pl->p_last); g.y = c.ws(pl->yc);
    g.stats = c.wsf(pl->bns[pl->bn_last].fstats); g.slots = pl->bns[pl->bn_last].fslots = GS;
    g.M = Ml; g.N = n.last_ch; g.K = Cl;
    TD3D_TRY(gemm_nt(c, g));
    TD3D_TRY(bn_forward(c, pl->bn_last, (double)Ml, training, &sc, &sh));
    TD3D_TRY(p_xform(c, c.ws(pl->yc), xf_make(sc, sh, nullptr, pl->stem_act), nullptr, nullptr,
                                c.wsf(pl->pool_stats), B, HWl, n.last_ch, dt, c.st));
    TD3D_K(PK_POOL, 8.0 * B * n.last_ch, launch_pool_finalize(c.wsf(pl->pool_stats), 1.f / (float)HWl, c.ws(pl->pooled), B, n.last_ch, dt, c.st));
  }
  // classifier: Linear -> BatchNorm1d -> h_swish (mobilenetv3.py:191-195; applied only when the backbone is the in-repo
  // MobileNetV3, model_builder.py:117-118,130-131 -- otherwise the heads read the pooled features)
  if (pl->has_fc) {
    GemmNT g = {};
    g.a = c.ws(pl->pooled); g.w = c.pk(pl->p_fc); g.y = c.ws(pl->yfc); g.bias = c.P(pl->b_fc);
    g.stats = c.wsf(pl->bns[pl->bn_fc].fstats); g.slots = B;
    g.M = B; g.N = n.head_ch; g.K = n.last_ch;
    TD3D_TRY(gemm_nt(c, g));
    TD3D_TRY(bn_forward(c, pl->bn_fc, (double)B, training, &sc, &sh));
    TD3D_TRY(p_xform(c, c.ws(pl->yfc), xf_make(sc, sh, nullptr, TD3D_ACT_HSWISH), nullptr, c.ws(pl->feat), nullptr,
                                B, 1, n.head_ch, dt, c.st));
  }
  *feat_out = c.ws(pl->feat);
  return TD3D_OK;
}

// ---- inference forward (eval mode): BatchNorm folded into the weights, activation / residual in the producer's
// epilogue -- every layer reads its input once and writes its activated output once (SURVEY.md 8d forward bytes).
// Only SE blocks keep one extra pass: the gate needs the whole plane's squeeze before act(gate * z) can be applied.
// Replaces eval-mode ModelWrapper.forward / forward_to_onnx (model_builder.py:112-146) over MobileNetV3.extract_features
// + classifier (mobilenetv3.py:188-203) with running-statistic BatchNorm.
static int forward_infer(const Ctx& c, const float* img, const void** feat_out) {
  td3d_plan* pl = c.pl;
  const td3d_net_desc& n = pl->net;
  const int B = pl->B, dt = pl->dtype;
  auto eshift = [&](int bn_idx) { return (const float*)c.pkf(pl->bns[bn_idx].eshift); };
  pl->prof.tag = 0;
  TD3D_K(PK_STEM_FWD, (double)B * 3 * pl->H * pl->W * 4 + (double)B * pl->H1 * pl->W1 * n.stem_ch * c.esz(),
         launch_stem_fwd(img, c.pkf(pl->pi_stem), c.ws(pl->x0), nullptr, B, pl->H, pl->W, n.stem_ch, dt, c.st,
                         eshift(pl->bn_stem), pl->stem_act));
  const void* cur = c.ws(pl->x0);
  for (auto& b : pl->blocks) {
    pl->prof.tag = (int)(&b - pl->blocks.data()) + 1;
    const int act = b.act;
    const int HWi = b.Hin * b.Win, HWo = b.Hout * b.Wout;
    const int Mi = B * HWi, Mo = B * HWo;
    const int E = b.d.exp_ch;
    DwArgs dw;
    dw.xf = xf_make(nullptr, nullptr, nullptr, TD3D_ACT_NONE);
    if (b.expand) {
      GemmNT g = {};
      g.a = cur; g.w = c.pk(b.pi1); g.y = c.ws(b.y1); g.bias = eshift(b.bn1); g.act = act;
      g.M = Mi; g.N = E; g.K = b.d.in_ch;
      TD3D_TRY(gemm_nt(c, g));
      dw.x = c.ws(b.y1);
    } else {
      dw.x = cur;
    }
    float* sq = c.wsf(pl->bns[b.bn2].fstats);          // per-sample sums of the depthwise output: the SE squeeze
    pdl_break(c.st);
    if (b.d.use_se) TD3D_CUDA(cudaMemsetAsync(sq, 0, sizeof(float) * 2 * (size_t)B * E, c.st));
    // dw-first layout is BN -> act -> SE (mobilenetv3.py:137-140), expanded layout BN -> SE -> act (:153-156)
    const bool act_in_dw = !b.d.use_se || b.se_post;
    dw.w_taps = c.pkf(b.pidw); dw.y = c.ws(b.d.use_se ? b.y2 : b.h2); dw.stats = b.d.use_se ? sq : nullptr;
    dw.out_bias = eshift(b.bn2); dw.out_act = act_in_dw ? act : TD3D_ACT_NONE;
    dw.B = B; dw.H = b.Hin; dw.W = b.Win; dw.C = E; dw.k = b.d.kernel; dw.stride = b.d.stride;
    {
      double in = (double)B * HWi * E, out = (double)B * HWo * E;
      TD3D_K(PK_DW_FWD, (in + out) * c.esz(), launch_dw_fwd_cw(dw, dt, c.st));
    }
    if (b.d.use_se) {
      SeArgs s;
      s.w1 = c.P(b.se_w1); s.b1 = c.P(b.se_b1); s.w2 = c.P(b.se_w2); s.b2 = c.P(b.se_b2);
      s.zbar = c.wsf(b.zbar); s.hid = c.wsf(b.hid); s.pre = c.wsf(b.pre); s.gate = c.wsf(b.gate);
      s.B = B; s.C = E; s.Ch = b.d.se_hidden; s.inv_hw = 1.f / (float)HWo; s.w2t = c.pkf(b.pse_w2t);
      s.pool_stats = sq; s.scale = nullptr; s.shift = nullptr;
      TD3D_K(PK_SE, 8.0 * b.d.exp_ch * b.d.se_hidden, se_forward(c, s));
      TD3D_TRY(p_xform(c, c.ws(b.y2), xf_make(nullptr, nullptr, s.gate, act_in_dw ? TD3D_ACT_NONE : act), nullptr, c.ws(b.h2),
                       nullptr, B, HWo, E, dt, c.st));
    }
    GemmNT g = {};
    g.a = c.ws(b.h2); g.w = c.pk(b.pi3); g.y = c.ws(b.out); g.bias = eshift(b.bn3);
    g.addend = b.residual ? cur : nullptr;
    g.M = Mo; g.N = b.d.out_ch; g.K = E;
    TD3D_TRY(gemm_nt(c, g));
    cur = c.ws(b.out);
  }
  pl->prof.tag = (int)pl->blocks.size() + 1;
  const int HWl = pl->Hl * pl->Wl, Ml = B * HWl;
  const int Cl = pl->blocks.back().d.out_ch;
  {
    GemmNT g = {};
    g.a = cur; g.w = c.pk(pl->pi_last); g.y = c.ws(pl->yc); g.bias = eshift(pl->bn_last); g.act = pl->stem_act;
    g.M = Ml; g.N = n.last_ch; g.K = Cl;
    TD3D_TRY(gemm_nt(c, g));
    pdl_break(c.st);
    TD3D_CUDA(cudaMemsetAsync(c.wsf(pl->pool_stats), 0, sizeof(float) * 2 * (size_t)B * n.last_ch, c.st));
    TD3D_TRY(p_xform(c, c.ws(pl->yc), xf_make(nullptr, nullptr, nullptr, TD3D_ACT_NONE), nullptr, nullptr,
                     c.wsf(pl->pool_stats), B, HWl, n.last_ch, dt, c.st));
    TD3D_K(PK_POOL, 8.0 * B * n.last_ch, launch_pool_finalize(c.wsf(pl->pool_stats), 1.f / (float)HWl, c.ws(pl->pooled), B, n.last_ch, dt, c.st));
  }
  if (pl->has_fc) {
    GemmNT g = {};
    g.a = c.ws(pl->pooled); g.w = c.pk(pl->pi_fc); g.y = c.ws(pl->feat); g.bias = eshift(pl->bn_fc); g.act = TD3D_ACT_HSWISH;
    g.M = B; g.N = n.head_ch; g.K = n.last_ch;
    TD3D_TRY(gemm_nt(c, g));
  }
  *feat_out = c.ws(pl->feat);
  return TD3D_OK;
}

static HeadsArgs heads_args(const Ctx& c, const void* feat, const int64_t* cats, const float* keep, uint64_t seed,
                            int training) {
  td3d_plan* pl = c.pl;
  HeadsArgs h;
  h.feat = feat; h.cats = cats;
  h.w_reg = c.P(pl->w_reg0); h.reg_stride = pl->head_stride;
  h.w_cls = c.P(pl->w_cls); h.b_cls = c.P(pl->b_cls);
  h.keep = keep; h.seed = seed; h.training = training; h.step_ptr = pl->dropout_counter;
  h.kp = c.wsf(pl->kp_saved); h.logits = c.wsf(pl->logits_saved);
  h.B = pl->B; h.C = pl->net.head_ch; h.P = pl->net.num_points; h.nc = pl->net.num_classes;
  h.max_classes = pl->net.max_classes;
  return h;
}

// ---- backward ---------------------------------------------------------------------------------
static int bn_backward(const Ctx& c, int idx, int HW, const float* se, const float* g_pool, const float* fwd_pool) {
  td3d_plan* pl = c.pl;
  Bn& bn = pl->bns[idx];
  BnBwdArgs a;
  a.stats = c.wsf(bn.bstats); a.slots = (bn.bslots && !se) ? bn.bslots : pl->B;
  a.mean = c.wsf(bn.mean); a.invstd = c.wsf(bn.invstd); a.gamma = c.P(bn.gamma);
  a.se = se; a.g_pool = g_pool; a.fwd_pool = fwd_pool;
  a.alpha = c.wsf(pl->alpha); a.beta = c.wsf(pl->beta); a.gammac = c.wsf(pl->gammac);
  a.dgamma = c.G(bn.gamma); a.dbeta = c.G(bn.beta);
  a.B = pl->B; a.HW = HW; a.C = bn.C;
  TD3D_K(PK_BN, 8.0 * pl->B * bn.C, launch_bn_bwd_finalize(a, c.st));
  return TD3D_OK;
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ gammac,
              const float* __restrict__ bstats, const float* __restrict__ fstats, float* __restrict__ out, int B, int C) {
  pdl_entry();
  // sum_b (alpha[b,c]*g_u + beta[c]*y + gammac[b,c]) from the per-sample sums (HW = 1)
  __shared__ double s[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int cc = blockIdx.x * 32 + lane;
  double acc = 0.0;
  if (cc < C)
    for (int b = w; b < B; b += 8)
      acc += (double)alpha[(size_t)b * C + cc] * bstats[((size_t)b * 2) * C + cc] +
             (double)beta[cc] * fstats[((size_t)b * 2) * C + cc] + (double)gammac[(size_t)b * C + cc];
  s[w][lane] = acc;
  __syncthreads();
  if (w == 0 && cc < C) {
    for (int i = 1; i < 8; ++i) acc += s[i][lane];
    out[cc] = (float)acc;
  }
}

__global__ void scale_kernel(const float* __restrict__ src, float s, float* __restrict__ dst, int n) {
  pdl_entry();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i] * s;
}

static int n_stages(const td3d_plan* pl) { return (int)pl->blocks.size() + 2; }

static int backward_impl(const Ctx& c, const float* d_kp, const float* d_logits, int32_t* present, int s_begin,
                         int s_end) {
  td3d_plan* pl = c.pl;
  const td3d_net_desc& n = pl->net;
  const int B = pl->B, dt = pl->dtype;
  const int nblk = (int)pl->blocks.size();
  TD3D_REQUIRE(pl->last_training, "backward: no training-mode forward recorded on this plan");
  auto in_range = [&](int s) { return s >= s_begin && s < s_end; };
  const int HWl = pl->Hl * pl->Wl, Ml = B * HWl;
  const int Cl = pl->blocks.back().d.out_ch;

  if (in_range(0)) {
    pl->prof.tag = 1000 + nblk + 1;
    pdl_break(c.st);
    TD3D_CUDA(cudaMemsetAsync(pl->G, 0, sizeof(float) * (size_t)pl->param_floats, c.st));
    pdl_break(c.st);
    TD3D_CUDA(cudaMemsetAsync(pl->WS + pl->bstats_begin, 0, pl->bstats_end - pl->bstats_begin, c.st));
    pdl_break(c.st);
    TD3D_CUDA(cudaMemsetAsync(pl->WS + pl->zeros_c, 0, sizeof(float) * 16, c.st));
    // heads
    HeadsBwdArgs hb;
    hb.f = heads_args(c, c.ws(pl->feat), pl->last_cats, pl->last_keep, pl->last_seed, 1);
    hb.d_kp = d_kp; hb.d_logits = d_logits;
    hb.g_pre = c.wsf(pl->g_pre_heads); hb.g_feat = c.wsf(pl->g_feat);
    hb.dw_reg = c.G(pl->w_reg0); hb.dw_cls = c.G(pl->w_cls); hb.db_cls = c.G(pl->b_cls);
    hb.present = present;
    TD3D_K(PK_HEADS, 8.0 * B * n.head_ch, launch_heads_bwd(hb, dt, c.st));
    const float* g_pooled = c.wsf(pl->g_feat);          // no classifier: the heads' input gradient IS the pooled gradient
    if (pl->has_fc) {
    // classifier: feat = h_swish(BN1d(yfc))
    Bn& bfc = pl->bns[pl->bn_fc];
    XForm xfc = xf_make(c.wsf(bfc.scale), c.wsf(bfc.shift), nullptr, TD3D_ACT_HSWISH);
    TD3D_TRY(p_actbwd(c, nullptr, c.wsf(pl->g_feat), 1.f, c.ws(pl->yfc), xfc, c.ws(pl->g_wide_a), c.wsf(bfc.bstats),
                                  B, 1, n.head_ch, dt, c.st));
    TD3D_TRY(bn_backward(c, pl->bn_fc, 1, nullptr, nullptr, nullptr));
    TD3D_CUDA(launch_kernel(colsum_kernel, ceil_div(n.head_ch, 32), 256, 0, c.st, c.wsf(pl->alpha), c.wsf(pl->beta), c.wsf(pl->gammac),
                                                              c.wsf(bfc.bstats), c.wsf(bfc.fstats), c.G(pl->b_fc), B, n.head_ch));
    TD3D_LAUNCH_CHECK();
    TD3D_TRY(p_affine2(c, c.ws(pl->g_wide_a), c.ws(pl->yfc), c.wsf(pl->alpha), c.wsf(pl->beta), c.wsf(pl->gammac),
                            c.ws(pl->g_wide_a), B, 1, n.head_ch, dt, c.st));
    {
      GemmTN t = {c.ws(pl->g_wide_a), c.ws(pl->pooled), c.G(pl->w_fc), B, n.head_ch, n.last_ch};
      TD3D_TRY(gemm_tn(c, t));
      GemmNT g = {};
      g.a = c.ws(pl->g_wide_a); g.w = c.pk(pl->p_fct); g.y = c.wsf(pl->g_pool_f); g.out_f32 = 1;
      g.M = B; g.N = n.last_ch; g.K = n.head_ch;
      TD3D_TRY(gemm_nt(c, g, PK_GEMM_DGRAD));
    }
    TD3D_TRY(side_join(c, 0));        // the classifier weight-gradient GEMM reads g_wide_a, rewritten next
    g_pooled = c.wsf(pl->g_pool_f);
    }
    // avg-pool backward + activation + BN of the final conv
    Bn& bl = pl->bns[pl->bn_last];
    XForm xl = xf_make(c.wsf(bl.scale), c.wsf(bl.shift), nullptr, pl->stem_act);
    TD3D_TRY(p_actbwd(c, nullptr, g_pooled, 1.f / (float)HWl, c.ws(pl->yc), xl, c.ws(pl->g_wide_a),
                                  c.wsf(bl.bstats), B, HWl, n.last_ch, dt, c.st));
    TD3D_TRY(bn_backward(c, pl->bn_last, HWl, nullptr, nullptr, nullptr));
    TD3D_TRY(p_affine2(c, c.ws(pl->g_wide_a), c.ws(pl->yc), c.wsf(pl->alpha), c.wsf(pl->beta), c.wsf(pl->gammac),
                            c.ws(pl->g_wide_a), B, HWl, n.last_ch, dt, c.st));
    {
      Block& lb = pl->blocks.back();
      GemmTN t = {c.ws(pl->g_wide_a), c.ws(lb.out), c.G(pl->w_last), Ml, n.last_ch, Cl};
      TD3D_TRY(gemm_tn(c, t));
      GemmNT g = {};
      g.a = c.ws(pl->g_wide_a); g.w = c.pk(pl->p_lastt); g.y = c.ws(pl->g_narrow[nblk & 1]);
      g.ysaved = c.ws(lb.y3); g.stats = c.wsf(pl->bns[lb.bn3].bstats); g.slots = pl->bns[lb.bn3].bslots = B < 32 ? B : 32;
      g.M = Ml; g.N = Cl; g.K = n.last_ch;
      TD3D_TRY(gemm_nt(c, g, PK_GEMM_DGRAD));
    }
  }
  // blocks, last to first. Gradient w.r.t. the output of block i lives in g_narrow[(i+1)&1].
  for (int i = nblk - 1; i >= 0; --i) {
    if (!in_range(nblk - i)) continue;
    Block& b = pl->blocks[i];
    pl->prof.tag = 1000 + i + 1;
    TD3D_TRY(side_join(c, 0));        // weight-gradient GEMMs of the previous stage read g_wide_a/b and g_y3, rewritten below
    const int act = b.act;
    const int HWi = b.Hin * b.Win, HWo = b.Hout * b.Wout;
    const int Mi = B * HWi, Mo = B * HWo;
    const int E = b.d.exp_ch;
    void* g_out = c.ws(pl->g_narrow[(i + 1) & 1]);
    void* g_in = c.ws(pl->g_narrow[i & 1]);
    const void* x_in = i == 0 ? c.ws(pl->x0) : c.ws(pl->blocks[i - 1].out);
    // BN3 (linear): g_y3 = alpha*g_out + beta*y3 + gamma
    TD3D_TRY(bn_backward(c, b.bn3, HWo, nullptr, nullptr, nullptr));
    TD3D_TRY(p_affine2(c, g_out, c.ws(b.y3), c.wsf(pl->alpha), c.wsf(pl->beta), c.wsf(pl->gammac), c.ws(pl->g_y3), B,
                            HWo, b.d.out_ch, dt, c.st));
    {
      GemmTN t = {c.ws(pl->g_y3), c.ws(b.h2), c.G(b.w3), Mo, b.d.out_ch, E};
      TD3D_TRY(gemm_tn(c, t));
      GemmNT g = {};
      g.a = c.ws(pl->g_y3); g.w = c.pk(b.pw3t); g.y = c.ws(pl->g_wide_a);
      g.M = Mo; g.N = E; g.K = b.d.out_ch;
      TD3D_TRY(gemm_nt(c, g, PK_GEMM_DGRAD));
    }
    Bn& bn2 = pl->bns[b.bn2];
    const float* sc2 = c.wsf(bn2.scale);
    const float* sh2 = c.wsf(bn2.shift);
    void* gw = c.ws(pl->g_wide_a);
    auto se_bwd_args = [&](const float* stats, const float* scale, const float* shift) {
      SeBwdArgs s;
      s.bwd_stats = stats; s.scale = scale; s.shift = shift; s.inv_hw = 1.f / (float)HWo;
      s.w1t = c.pkf(b.pse_w1t); s.w2t = c.pkf(b.pse_w2t); s.w1 = c.P(b.se_w1);
      s.zbar = c.wsf(b.zbar); s.hid = c.wsf(b.hid); s.pre = c.wsf(b.pre);
      s.g_pre = c.wsf(pl->se_gpre); s.g_hid = c.wsf(pl->se_ghid); s.g_pool = c.wsf(pl->se_gpool);
      s.dw1 = c.G(b.se_w1); s.db1 = c.G(b.se_b1); s.dw2 = c.G(b.se_w2); s.db2 = c.G(b.se_b2);
      s.B = B; s.C = E; s.Ch = b.d.se_hidden;
      return s;
    };
    if (b.d.use_se && !b.se_post) {
      // x = act(gate * BN(y2)): the gate sits inside the activation (mobilenetv3.py:153-156)
      const float* gate = c.wsf(b.gate);
      TD3D_TRY(p_actbwd(c, gw, nullptr, 1.f, c.ws(b.y2), xf_make(sc2, sh2, gate, act), gw, c.wsf(bn2.bstats), B, HWo,
                                    E, dt, c.st));
      TD3D_K(PK_SE, 16.0 * b.d.exp_ch * b.d.se_hidden, se_backward(c, se_bwd_args(c.wsf(bn2.bstats), sc2, sh2)));
      TD3D_TRY(bn_backward(c, b.bn2, HWo, gate, c.wsf(pl->se_gpool), c.wsf(bn2.fstats)));
    } else {
      if (b.d.use_se) {
        // x = H * gate with H = act(BN(y2)), H recomputed from y2:
        //   pass 1 (read only): d loss / d gate = sum_HW g_x * H        -> SE backward -> g_pool (gradient of the squeeze)
        //   pass 2            : g_z = (gate*g_x + g_pool/HW) * act'(z)  + the BatchNorm-backward sums of bn2
        TD3D_TRY(p_actbwd(c, gw, nullptr, 1.f, c.ws(b.y2), xf_make(sc2, sh2, nullptr, act, 1), nullptr, c.wsf(b.hbstats), B, HWo,
                                      E, dt, c.st));
        TD3D_K(PK_SE, 16.0 * b.d.exp_ch * b.d.se_hidden, se_backward(c, se_bwd_args(c.wsf(b.hbstats), nullptr, nullptr)));
        TD3D_TRY(p_actbwd(c, gw, c.wsf(pl->se_gpool), 1.f / (float)HWo, c.ws(b.y2), xf_make(sc2, sh2, c.wsf(b.gate), act, 1), gw,
                                      c.wsf(bn2.bstats), B, HWo, E, dt, c.st));
      } else {
        // Fusing this pass into the dgrad GEMM epilogue was built and measured (r02 call H): act_bwd_stats -0.48 ms, dgrad GEMM
        // +0.37 ms, and the extra epilogue registers spilled and slowed EVERY tcgen05 GEMM (forward 1.77 -> 2.14 ms) -- removed.
        TD3D_TRY(p_actbwd(c, gw, nullptr, 1.f, c.ws(b.y2), xf_make(sc2, sh2, nullptr, act), gw, c.wsf(bn2.bstats), B, HWo, E,
                                      dt, c.st));
      }
      TD3D_TRY(bn_backward(c, b.bn2, HWo, nullptr, nullptr, nullptr));
    }
    // depthwise conv backward (data + weights)
    DwBwdArgs d;
    d.g = gw; d.y_out = c.ws(b.y2);
    d.alpha = c.wsf(pl->alpha); d.beta = c.wsf(pl->beta); d.gamma = c.wsf(pl->gammac);
    d.w_taps = c.pkf(b.pdw); d.dw = c.G(b.wdw);
    d.B = B; d.H = b.Hin; d.W = b.Win; d.C = E; d.k = b.d.kernel; d.stride = b.d.stride;
    const void* prev_y = i == 0 ? c.ws(pl->y0) : c.ws(pl->blocks[i - 1].y3);
    float* prev_bstats = c.wsf(pl->bns[i == 0 ? pl->bn_stem : pl->blocks[i - 1].bn3].bstats);
    if (b.expand) {
      Bn& bn1 = pl->bns[b.bn1];
      d.x = c.ws(b.y1); d.xf = xf_make(c.wsf(bn1.scale), c.wsf(bn1.shift), nullptr, act);
      d.gx = c.ws(pl->g_wide_b); d.stats = c.wsf(bn1.bstats);
      TD3D_TRY(p_dwb(c, d, dt, c.st));
      TD3D_TRY(bn_backward(c, b.bn1, HWi, nullptr, nullptr, nullptr));
      TD3D_TRY(p_affine2(c, c.ws(pl->g_wide_b), c.ws(b.y1), c.wsf(pl->alpha), c.wsf(pl->beta), c.wsf(pl->gammac),
                              c.ws(pl->g_wide_b), B, HWi, E, dt, c.st));
      GemmTN t = {c.ws(pl->g_wide_b), x_in, c.G(b.w1), Mi, E, b.d.in_ch};
      TD3D_TRY(gemm_tn(c, t));
      GemmNT g = {};
      g.a = c.ws(pl->g_wide_b); g.w = c.pk(b.pw1t); g.y = g_in;
      g.addend = b.residual ? g_out : nullptr;
      if (i > 0) { g.ysaved = prev_y; g.stats = prev_bstats; g.slots = pl->bns[pl->blocks[i - 1].bn3].bslots = B < 32 ? B : 32; }
      g.M = Mi; g.N = b.d.in_ch; g.K = E;
      TD3D_TRY(gemm_nt(c, g, PK_GEMM_DGRAD));
    } else {
      d.x = x_in; d.xf = xf_make(nullptr, nullptr, nullptr, TD3D_ACT_NONE);
      d.gx = g_in; d.stats = nullptr;
      TD3D_TRY(p_dwb(c, d, dt, c.st));
      if (i > 0) {
        // previous block output is linear in y3: g_u = g (+ residual), statistics for its BN3 (one slot per sample)
        pl->bns[pl->blocks[i - 1].bn3].bslots = 0;
        TD3D_TRY(p_actbwd(c, g_in, nullptr, 1.f, prev_y, xf_make(nullptr, nullptr, nullptr, TD3D_ACT_NONE), g_in,
                                      prev_bstats, B, HWi, b.d.in_ch, dt, c.st, b.residual ? g_out : nullptr));
      } else if (b.residual) {
        // folded into the stem stage below (needs act'); add the residual gradient now
        TD3D_TRY(p_actbwd(c, g_in, nullptr, 1.f, prev_y, xf_make(nullptr, nullptr, nullptr, TD3D_ACT_NONE), g_in,
                                      nullptr, B, HWi, b.d.in_ch, dt, c.st, g_out));
      }
    }
  }
  if (in_range(nblk + 1)) {
    pl->prof.tag = 1000;
    TD3D_TRY(side_join(c, 0));
    // stem: x0 = h_swish(BN(y0)); gradient w.r.t. x0 is in g_narrow[0]
    Bn& b0 = pl->bns[pl->bn_stem];
    const int HW1 = pl->H1 * pl->W1;
    void* g0 = c.ws(pl->g_narrow[0]);
    TD3D_TRY(p_actbwd(c, g0, nullptr, 1.f, c.ws(pl->y0), xf_make(c.wsf(b0.scale), c.wsf(b0.shift), nullptr, pl->stem_act),
                                  g0, c.wsf(b0.bstats), B, HW1, n.stem_ch, dt, c.st));
    TD3D_TRY(bn_backward(c, pl->bn_stem, HW1, nullptr, nullptr, nullptr));
    TD3D_K(PK_STEM_WGRAD, (double)B * 3 * pl->H * pl->W * 4 + 2.0 * B * pl->H1 * pl->W1 * n.stem_ch * c.esz(), launch_stem_wgrad(pl->last_img, g0, c.ws(pl->y0), c.wsf(pl->alpha), c.wsf(pl->beta), c.wsf(pl->gammac),
                               c.G(pl->w_stem), B, pl->H, pl->W, n.stem_ch, dt, c.st));
  }
  TD3D_TRY(side_join(c, 0));          // every gradient of this stage range is final on c.st when the call returns
  return TD3D_OK;
}

static int add_seg(PackTable& t, const float* src, void* dst, int rows, int cols, int transpose, int out_dtype,
                   const float* row_scale = nullptr) {
  TD3D_REQUIRE(t.n < 256, "pack table overflow");
  PackSeg& sg = t.seg[t.n++];
  sg.src = src; sg.dst = dst; sg.rows = rows; sg.cols = cols; sg.transpose = transpose; sg.out_dtype = out_dtype;
  sg.row_scale = row_scale;
  return TD3D_OK;
}

// (re)build the pack / BN-fold tables for the currently bound buffers
static int build_tables(td3d_plan* pl) {
  Ctx c = {pl, 0};
  const td3d_net_desc& n = pl->net;
  const int dt = pl->dtype;
  PackTable& t = pl->pack_table;
  t.n = 0;
  TD3D_TRY(add_seg(t, c.P(pl->w_stem), c.pk(pl->p_stem), n.stem_ch, 27, 1, TD3D_F32));
  for (auto& b : pl->blocks) {
    const int E = b.d.exp_ch, kk = b.d.kernel * b.d.kernel;
    if (b.expand) {
      TD3D_TRY(add_seg(t, c.P(b.w1), c.pk(b.pw1), E, b.d.in_ch, 0, dt));
      TD3D_TRY(add_seg(t, c.P(b.w1), c.pk(b.pw1t), E, b.d.in_ch, 1, dt));
    }
    TD3D_TRY(add_seg(t, c.P(b.wdw), c.pk(b.pdw), E, kk, 1, TD3D_F32));
    TD3D_TRY(add_seg(t, c.P(b.w3), c.pk(b.pw3), b.d.out_ch, E, 0, dt));
    TD3D_TRY(add_seg(t, c.P(b.w3), c.pk(b.pw3t), b.d.out_ch, E, 1, dt));
    if (b.d.use_se) {
      TD3D_TRY(add_seg(t, c.P(b.se_w1), c.pk(b.pse_w1t), b.d.se_hidden, E, 1, TD3D_F32));
      TD3D_TRY(add_seg(t, c.P(b.se_w2), c.pk(b.pse_w2t), E, b.d.se_hidden, 1, TD3D_F32));
    }
  }
  const int Cl = pl->blocks.back().d.out_ch;
  TD3D_TRY(add_seg(t, c.P(pl->w_last), c.pk(pl->p_last), n.last_ch, Cl, 0, dt));
  TD3D_TRY(add_seg(t, c.P(pl->w_last), c.pk(pl->p_lastt), n.last_ch, Cl, 1, dt));
  if (pl->has_fc) {
    TD3D_TRY(add_seg(t, c.P(pl->w_fc), c.pk(pl->p_fc), n.head_ch, n.last_ch, 0, dt));
    TD3D_TRY(add_seg(t, c.P(pl->w_fc), c.pk(pl->p_fct), n.head_ch, n.last_ch, 1, dt));
  }
  BnFoldTable& f = pl->fold_table;
  f.n = 0;
  TD3D_REQUIRE(pl->bns.size() <= 128, "too many BatchNorm layers for the fold table");
  for (auto& bn : pl->bns) {
    BnFoldSeg& sg = f.seg[f.n++];
    sg.gamma = c.P(bn.gamma); sg.beta = c.P(bn.beta);
    sg.rm = pl->BNB + bn.rm; sg.rv = pl->BNB + bn.rm + bn.C;
    sg.scale = c.pkf(bn.escale); sg.shift = c.pkf(bn.eshift); sg.C = bn.C;
    sg.lin_bias = (pl->has_fc && (&bn - pl->bns.data()) == pl->bn_fc) ? c.P(pl->b_fc) : nullptr;   // classifier Linear bias rides in the folded shift
  }
  // inference weights: W'[n][k] = W[n][k] * escale[n]; the consumer's bias is the BatchNorm's eshift
  PackTable& te = pl->pack_table_eval;
  te.n = 0;
  auto esc = [&](int bn_idx) { return (const float*)c.pkf(pl->bns[bn_idx].escale); };
  TD3D_TRY(add_seg(te, c.P(pl->w_stem), c.pk(pl->pi_stem), n.stem_ch, 27, 1, TD3D_F32, esc(pl->bn_stem)));
  for (auto& b : pl->blocks) {
    const int E = b.d.exp_ch, kk = b.d.kernel * b.d.kernel;
    if (b.expand) TD3D_TRY(add_seg(te, c.P(b.w1), c.pk(b.pi1), E, b.d.in_ch, 0, dt, esc(b.bn1)));
    TD3D_TRY(add_seg(te, c.P(b.wdw), c.pk(b.pidw), E, kk, 1, TD3D_F32, esc(b.bn2)));
    TD3D_TRY(add_seg(te, c.P(b.w3), c.pk(b.pi3), b.d.out_ch, E, 0, dt, esc(b.bn3)));
  }
  TD3D_TRY(add_seg(te, c.P(pl->w_last), c.pk(pl->pi_last), n.last_ch, Cl, 0, dt, esc(pl->bn_last)));
  if (pl->has_fc) TD3D_TRY(add_seg(te, c.P(pl->w_fc), c.pk(pl->pi_fc), n.head_ch, n.last_ch, 0, dt, esc(pl->bn_fc)));
  return TD3D_OK;
}

// weights only (every optimizer step) / weights + eval-mode BN fold (explicit td3d_pack_weights)
static int pack_impl(const Ctx& c, bool with_bn_fold) {
  TD3D_TRY(launch_pack_table(c.pl->pack_table, c.st));
  if (with_bn_fold) {
    TD3D_TRY(launch_bn_fold_table(c.pl->fold_table, BN_EPS, c.st));
    TD3D_TRY(launch_pack_table(c.pl->pack_table_eval, c.st));
  }
  return TD3D_OK;
}

}  // namespace td3d

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int td3d_abi_version(void) { return TD3D_ABI_VERSION; }
const char* td3d_last_error(void) { return g_err; }

int td3d_device_check(void) {
  int dev = 0;
  TD3D_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TD3D_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_last_error("device %d is sm_%d%d; td3d kernels are built for sm_100a (B200) only", dev, prop.major, prop.minor);
    return TD3D_ECUDA;
  }
  return TD3D_OK;
}

int td3d_plan_create(const td3d_net_desc* net, int batch, int height, int width, int dtype, int gemm_impl, td3d_plan** out) {
  TD3D_REQUIRE(net && out, "plan_create: null argument");
  TD3D_REQUIRE(batch > 0 && batch <= 65535 && height >= 8 && width >= 8, "plan_create: bad batch/resolution %d %dx%d", batch, height, width);
  TD3D_REQUIRE(dtype == TD3D_F32 || dtype == TD3D_BF16, "plan_create: bad dtype %d", dtype);
  TD3D_REQUIRE(net->n_blocks > 0 && net->blocks, "plan_create: no blocks");
  TD3D_REQUIRE(net->stem_ch >= 16 && net->stem_ch <= 48 && net->stem_ch % 8 == 0, "plan_create: stem_ch must be 16..48 in steps of 8");
  TD3D_REQUIRE(net->arch == TD3D_ARCH_MOBILENETV3 || net->arch == TD3D_ARCH_EFFICIENTNET, "plan_create: unknown arch %d", net->arch);
  TD3D_REQUIRE(net->num_points == 18, "plan_create: num_points must be 18 (9 keypoints)");
  TD3D_REQUIRE(net->num_classes >= 1 && net->num_classes <= 32 && net->max_classes >= 1 && net->max_classes <= 32, "plan_create: bad class counts");
  TD3D_REQUIRE(net->last_ch % 8 == 0 && net->head_ch % 8 == 0, "plan_create: widths must be multiples of 8");
  td3d_plan* pl = new td3d_plan();
  pl->net = *net;
  pl->block_descs.assign(net->blocks, net->blocks + net->n_blocks);
  pl->net.blocks = pl->block_descs.data();
  pl->B = batch; pl->H = height; pl->W = width; pl->dtype = dtype; pl->gemm_impl = gemm_impl;
  pl->esz = dtype == TD3D_BF16 ? 2 : 4;
  int rc = build(pl);
  if (rc != TD3D_OK) { delete pl; return rc; }
  *out = pl;
  return TD3D_OK;
}

void td3d_plan_destroy(td3d_plan* plan) {
  if (!plan) return;
  for (int i = 0; i < 1; ++i) {
    if (plan->side[i]) cudaStreamDestroy(plan->side[i]);
    if (plan->ev_fork[i]) cudaEventDestroy(plan->ev_fork[i]);
    if (plan->ev_join[i]) cudaEventDestroy(plan->ev_join[i]);
  }
  delete plan;
}

int td3d_plan_sizes(const td3d_plan* pl, td3d_sizes* out) {
  TD3D_REQUIRE(pl && out, "plan_sizes: null argument");
  out->n_param_tensors = (int64_t)pl->params.size();
  out->param_floats = pl->param_floats;
  out->n_bn = (int64_t)pl->bns.size();
  out->bn_floats = pl->bn_floats;
  out->packed_bytes = (int64_t)pl->packed_bytes;
  out->workspace_bytes = (int64_t)pl->ws_bytes;
  out->head_param_offset = pl->w_reg0;
  out->head_param_stride = pl->head_stride;
  return TD3D_OK;
}

int td3d_plan_param_info(const td3d_plan* pl, int64_t index, td3d_param_info* out) {
  TD3D_REQUIRE(pl && out && index >= 0 && index < (int64_t)pl->params.size(), "param_info: bad index");
  *out = pl->params[index].info;
  return TD3D_OK;
}

int td3d_plan_bn_info(const td3d_plan* pl, int64_t index, td3d_bn_info* out) {
  TD3D_REQUIRE(pl && out && index >= 0 && index < (int64_t)pl->bns.size(), "bn_info: bad index");
  memset(out, 0, sizeof(*out));
  snprintf(out->name, sizeof(out->name), "%s", pl->bns[index].name.c_str());
  out->offset = pl->bns[index].rm;
  out->channels = pl->bns[index].C;
  return TD3D_OK;
}

int td3d_plan_bind(td3d_plan* pl, float* params, float* grads, float* bn_stats, int64_t* nbt, void* packed, void* workspace) {
  TD3D_REQUIRE(pl && params && bn_stats && packed && workspace, "plan_bind: null buffer");
  TD3D_REQUIRE(((uintptr_t)params & 15) == 0 && ((uintptr_t)packed & 255) == 0 && ((uintptr_t)workspace & 255) == 0,
               "plan_bind: buffers must be 256-byte aligned");
  pl->P = params; pl->G = grads; pl->BNB = bn_stats; pl->NBT = nbt;
  pl->PK = (uint8_t*)packed; pl->WS = (uint8_t*)workspace;
  if (!pl->side[0]) {
    const char* e = getenv("TD3D_OVERLAP");
    pl->overlap = (e && atoi(e) == 0) ? 0 : 1;
    for (int i = 0; i < 1; ++i) {
      TD3D_CUDA(cudaStreamCreateWithFlags(&pl->side[i], cudaStreamNonBlocking));
      TD3D_CUDA(cudaEventCreateWithFlags(&pl->ev_fork[i], cudaEventDisableTiming));
      TD3D_CUDA(cudaEventCreateWithFlags(&pl->ev_join[i], cudaEventDisableTiming));
    }
  }
  return build_tables(pl);
}

int td3d_plan_set_dropout_counter(td3d_plan* pl, const int32_t* counter) {
  TD3D_REQUIRE(pl, "set_dropout_counter: null plan");
  pl->dropout_counter = counter;
  return TD3D_OK;
}

int td3d_plan_profile(td3d_plan* pl, int enable) {
  TD3D_REQUIRE(pl, "plan_profile: null plan");
  for (auto& r : pl->prof.recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  pl->prof.recs.clear();
  pl->prof.enabled = enable != 0;
  return TD3D_OK;
}

int td3d_plan_profile_read(td3d_plan* pl, int kind, char* name, int name_cap, double* ms, double* bytes, int64_t* launches) {
  TD3D_REQUIRE(pl && ms && bytes && launches, "profile_read: null argument");
  if (kind < 0 || kind >= PK_COUNT) return 1;   // end of table
  if (name && name_cap > 0) snprintf(name, (size_t)name_cap, "%s", kProfNames[kind]);
  *ms = 0; *bytes = 0; *launches = 0;
  for (auto& r : pl->prof.recs) {
    if (r.kind != kind) continue;
    TD3D_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    TD3D_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    *ms += t; *bytes += r.bytes; *launches += 1;
  }
  return TD3D_OK;
}

int td3d_plan_profile_launch(td3d_plan* pl, int64_t index, int* kind, int* tag, double* ms, double* bytes) {
  TD3D_REQUIRE(pl && kind && tag && ms && bytes, "profile_launch: null argument");
  if (index < 0 || index >= (int64_t)pl->prof.recs.size()) return 1;   // end of the record list
  ProfRec& r = pl->prof.recs[(size_t)index];
  TD3D_CUDA(cudaEventSynchronize(r.e1));
  float t = 0.f;
  TD3D_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
  *kind = r.kind; *tag = r.tag; *ms = t; *bytes = r.bytes;
  return TD3D_OK;
}

#define TD3D_BOUND(pl) TD3D_REQUIRE((pl) && (pl)->P && (pl)->WS, "plan is not bound (call td3d_plan_bind)")

int td3d_pack_weights(td3d_plan* pl, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  TD3D_BOUND(pl);
  Ctx c = {pl, (cudaStream_t)stream};
  return pack_impl(c, true);
}

int td3d_forward(td3d_plan* pl, const float* img, const int64_t* cats, const float* dropout_keep, uint64_t seed,
                 int training, float* kp, float* logits, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  TD3D_BOUND(pl);
  TD3D_REQUIRE(img && cats && kp && logits, "forward: null argument");
  Ctx c = {pl, (cudaStream_t)stream};
  const void* feat = nullptr;
  if (training) TD3D_TRY(forward_backbone(c, img, 1, &feat));
  else TD3D_TRY(forward_infer(c, img, &feat));
  HeadsArgs h = heads_args(c, feat, cats, dropout_keep, seed, training);
  TD3D_K(PK_HEADS, 4.0 * pl->B * pl->net.head_ch, launch_heads_fwd(h, pl->dtype, c.st));
  pdl_break(c.st);
  TD3D_CUDA(cudaMemcpyAsync(kp, h.kp, sizeof(float) * pl->B * pl->net.num_points, cudaMemcpyDeviceToDevice, c.st));
  pdl_break(c.st);
  TD3D_CUDA(cudaMemcpyAsync(logits, h.logits, sizeof(float) * pl->B * pl->net.num_classes, cudaMemcpyDeviceToDevice, c.st));
  pl->last_img = img; pl->last_cats = cats; pl->last_keep = dropout_keep; pl->last_seed = seed; pl->last_training = training;
  return TD3D_OK;
}

int td3d_forward_export(td3d_plan* pl, const float* img, float* kp_all, float* logits, int select, float* kp_sel,
                        int64_t* labels, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  TD3D_BOUND(pl);
  TD3D_REQUIRE(img && kp_all && logits, "forward_export: null argument");
  TD3D_REQUIRE(!select || (kp_sel && labels), "forward_export: select needs kp_sel and labels");
  Ctx c = {pl, (cudaStream_t)stream};
  const void* feat = nullptr;
  TD3D_TRY(forward_infer(c, img, &feat));
  HeadsArgs h = heads_args(c, feat, nullptr, nullptr, 0, 0);
  h.logits = logits;
  TD3D_TRY(launch_heads_all(h, kp_all, pl->dtype, c.st));
  if (select)
    TD3D_TRY(launch_select_argmax(kp_all, logits, kp_sel, labels, pl->B, pl->net.num_points, pl->net.num_classes,
                                  pl->net.max_classes, c.st));
  pl->last_training = 0;
  return TD3D_OK;
}

int td3d_backward_stages(const td3d_plan* pl) { return pl ? n_stages(pl) : 0; }

int td3d_backward(td3d_plan* pl, const float* d_kp, const float* d_logits, int32_t* head_present, int stage_begin,
                  int stage_end, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  TD3D_BOUND(pl);
  TD3D_REQUIRE(pl->G, "backward: no gradient arena bound");
  TD3D_REQUIRE(d_kp && d_logits && head_present, "backward: null argument");
  if (stage_end < 0) stage_end = n_stages(pl);
  Ctx c = {pl, (cudaStream_t)stream};
  return backward_impl(c, d_kp, d_logits, head_present, stage_begin, stage_end);
}

int td3d_backward_ready_range(const td3d_plan* pl, int stage, int64_t* begin, int64_t* end) {
  TD3D_REQUIRE(pl && begin && end, "ready_range: null argument");
  const int nblk = (int)pl->blocks.size();
  TD3D_REQUIRE(stage >= 1 && stage <= nblk + 2, "ready_range: stage out of range");
  // after stages [0,stage) : tail params, then blocks nblk-1 .. nblk-(stage-1)
  *end = pl->param_floats;
  if (stage >= nblk + 2) *begin = 0;
  else if (stage == 1) *begin = pl->first_param_tail;
  else *begin = pl->blocks[nblk - (stage - 1)].first_param;
  return TD3D_OK;
}

int td3d_loss_fwd_bwd(const td3d_loss_desc* desc, const float* kp, const float* gt_kp, const float* logits,
                      const int64_t* cats, int batch, int num_classes, float* loss_out, float* d_kp, float* d_logits,
                      void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  TD3D_REQUIRE(desc && kp && gt_kp && loss_out, "loss: null argument");
  return launch_loss(*desc, kp, gt_kp, logits, cats, batch, num_classes, loss_out, d_kp, d_logits, (cudaStream_t)stream);
}

int td3d_metrics_accum(const float* kp, const float* gt_kp, const float* logits, const int64_t* cats, int batch,
                       int num_classes, int max_classes, double* acc, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  TD3D_REQUIRE(kp && gt_kp && cats && acc, "metrics: null argument");
  return launch_metrics(kp, gt_kp, logits, cats, batch, num_classes, max_classes, acc, (cudaStream_t)stream);
}

int td3d_optim_step(td3d_plan* pl, const td3d_optim_desc* desc, float* state0, float* state1, int32_t* steps,
                    const int32_t* head_present, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  TD3D_BOUND(pl);
  TD3D_REQUIRE(desc && state0 && steps && pl->G, "optim_step: null argument");
  TD3D_REQUIRE(desc->kind >= TD3D_OPT_SGD && desc->kind <= TD3D_OPT_ADADELTA, "optim_step: unknown optimizer %d", desc->kind);
  TD3D_REQUIRE(!(desc->kind == TD3D_OPT_ADAMW || desc->kind == TD3D_OPT_ADADELTA) || state1, "optim_step: state1 required");
  OptimArgs a;
  a.d = *desc;
  a.p = pl->P; a.g = pl->G; a.s0 = state0; a.s1 = state1; a.n = pl->param_floats;
  a.head_off = pl->w_reg0; a.head_stride = pl->head_stride; a.n_heads = pl->net.max_classes;
  a.steps = steps; a.present = head_present;
  a.skip_begin = a.skip_end = 0;
  if (pl->net.num_classes == 1) { a.skip_begin = pl->w_cls; a.skip_end = pl->param_floats; }
  Ctx c = {pl, (cudaStream_t)stream};
  {
    double per = desc->kind == TD3D_OPT_ADAMW ? 28.0 : (desc->kind == TD3D_OPT_ADADELTA ? 28.0 : 20.0);
    TD3D_K(PK_OPTIM, per * pl->param_floats, launch_optim(a, c.st));
  }
  TD3D_K(PK_PACK, 12.0 * pl->param_floats, pack_impl(c, false));
  return TD3D_OK;
}

int td3d_roi_crop_resize(const uint8_t* frames, int n_frames, int frame_h, int frame_w, const int32_t* boxes, int n_boxes,
                         int out_h, int out_w, const float* mean255, const float* inv_std255, int swap_rb, float* out,
                         void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  return launch_roi_crop_resize(frames, n_frames, frame_h, frame_w, boxes, n_boxes, out_h, out_w, mean255, inv_std255, swap_rb,
                                out, (cudaStream_t)stream);
}

// ---- per-kernel entry points ----------------------------------------------------------------
int td3d_lift_2d(const float* kp, int n, int portrait, const double* cam_ndc, double* out, void* stream) {
  pdl_break_all();
  return launch_lift_2d(kp, n, portrait, cam_ndc, out, (cudaStream_t)stream);
}
int td3d_iou_2d_based(const float* pred_kp, const float* gt_kp, int n, int portrait, const double* cam_ndc, double* iou, void* stream) {
  pdl_break_all();
  return launch_iou_2d_based(pred_kp, gt_kp, n, portrait, cam_ndc, iou, (cudaStream_t)stream);
}
int td3d_k_stem_fwd(const float* img, const float* w27x16, void* y, float* stats, int B, int H, int W, int C, int dtype,
                    void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  return launch_stem_fwd(img, w27x16, y, stats, B, H, W, C, dtype, (cudaStream_t)stream);
}
int td3d_k_stem_wgrad(const float* img, const void* g, const void* y, const float* alpha, const float* beta,
                      const float* gamma, float* dw, int B, int H, int W, int C, int dtype, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  return launch_stem_wgrad(img, g, y, alpha, beta, gamma, dw, B, H, W, C, dtype, (cudaStream_t)stream);
}
int td3d_k_dw_fwd(const void* x, const float* scale, const float* shift, const float* se, int act, const float* w_taps,
                  void* y, float* stats, int B, int H, int W, int C, int k, int stride, int dtype, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  DwArgs a;
  a.x = x; a.xf = xf_make(scale, shift, se, act); a.w_taps = w_taps; a.y = y; a.stats = stats;
  a.B = B; a.H = H; a.W = W; a.C = C; a.k = k; a.stride = stride;
  return launch_dw_fwd(a, dtype, (cudaStream_t)stream);
}
int td3d_k_dw_bwd(const void* g, const void* y_out, const float* alpha, const float* beta, const float* gamma,
                  const void* x, const float* scale, const float* shift, const float* se, int act, const float* w_taps,
                  void* gx, float* dw, float* stats, int B, int H, int W, int C, int k, int stride, int dtype, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  DwBwdArgs a;
  a.g = g; a.y_out = y_out; a.alpha = alpha; a.beta = beta; a.gamma = gamma;
  a.x = x; a.xf = xf_make(scale, shift, se, act); a.w_taps = w_taps; a.gx = gx; a.stats = stats; a.dw = dw;
  a.B = B; a.H = H; a.W = W; a.C = C; a.k = k; a.stride = stride;
  return launch_dw_bwd(a, dtype, (cudaStream_t)stream);
}
int td3d_k_dw_fwd_ex(const void* x, const float* scale, const float* shift, const float* se, int act, const float* w_taps,
                     const float* out_bias, int out_act, void* y, float* stats, int B, int H, int W, int C, int k, int stride,
                     int dtype, int impl, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  DwArgs a;
  a.x = x; a.xf = xf_make(scale, shift, se, act); a.w_taps = w_taps; a.y = y; a.stats = stats;
  a.B = B; a.H = H; a.W = W; a.C = C; a.k = k; a.stride = stride; a.out_bias = out_bias; a.out_act = out_act;
  TD3D_REQUIRE((k == 3 || k == 5) && (stride == 1 || stride == 2) && C % 8 == 0, "dw_fwd_ex: bad kernel/stride/channels");
  if (impl == 2) return launch_dw_fwd_cw(a, dtype, (cudaStream_t)stream);
  if (impl == 1) {
    TD3D_REQUIRE(!out_bias && out_act == TD3D_ACT_NONE && act != TD3D_ACT_SILU, "dw_fwd_ex: epilogue / SiLU need the column walker");
    if (dw_walker_supported(H, W, C, k, stride)) return launch_dw_fwd_walker(a, dtype, (cudaStream_t)stream);
    return launch_dw_fwd_v2(a, dtype, (cudaStream_t)stream);
  }
  return launch_dw_fwd(a, dtype, (cudaStream_t)stream);
}
int td3d_k_gemm_nt(const void* a, const void* w, void* y, const void* addend, const float* bias, const void* ysaved,
                   float* stats, int stat_slots, int M, int N, int K, int dtype, int out_f32, int impl, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  GemmNT g = {};
  g.a = a; g.w = w; g.y = y; g.addend = addend; g.bias = bias; g.ysaved = ysaved; g.stats = stats; g.slots = stat_slots;
  g.M = M; g.N = N; g.K = K; g.out_f32 = out_f32;
  if (impl == TD3D_GEMM_TCGEN05) {
    TD3D_REQUIRE(dtype == TD3D_BF16, "tcgen05 GEMM is bf16 only");
    return launch_gemm_nt_tc(g, (cudaStream_t)stream);
  }
  return launch_gemm_nt_simt(g, dtype, (cudaStream_t)stream);
}
int td3d_k_gemm_tn(const void* a, const void* b, float* c, int M, int N1, int N2, int dtype, int impl, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  GemmTN g = {a, b, c, M, N1, N2};
  if (impl == TD3D_GEMM_TCGEN05) {
    TD3D_REQUIRE(dtype == TD3D_BF16, "tcgen05 GEMM is bf16 only");
    return launch_gemm_tn_tc(g, (cudaStream_t)stream);
  }
  return launch_gemm_tn_simt(g, dtype, (cudaStream_t)stream);
}
int td3d_debug_tc_timeline(uint64_t* out, int n) { return tc_timeline_read((unsigned long long*)out, n); }
int td3d_k_apply_xform(const void* y, const float* scale, const float* shift, const float* se, int act, const void* res,
                       void* out, float* pool_stats, int B, int HW, int C, int dtype, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  return launch_apply_xform(y, xf_make(scale, shift, se, act), res, out, pool_stats, B, HW, C, dtype, (cudaStream_t)stream);
}
int td3d_k_affine2(const void* g, const void* y, const float* alpha, const float* beta, const float* gamma, void* out,
                   int B, int HW, int C, int dtype, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  return launch_affine2(g, y, alpha, beta, gamma, out, B, HW, C, dtype, (cudaStream_t)stream);
}
int td3d_k_act_bwd_stats(const void* g, const void* y, const float* scale, const float* shift, const float* se, int act,
                         void* gu, float* stats, int B, int HW, int C, int dtype, void* stream) {
  pdl_break_all();                           // other libraries' work may precede this call on the stream
  return launch_act_bwd_stats(g, nullptr, 1.f, y, xf_make(scale, shift, se, act), gu, stats, B, HW, C, dtype,
                              (cudaStream_t)stream, nullptr);
}

}  // extern "C"
