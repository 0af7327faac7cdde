// CPU emulation of the one-pass depthwise backward: runs the exact per-thread body of csrc/dwc_core.cuh
// (host side of the __host__ __device__ code) over every (channel group, item lane) sequentially and compares with
// a naive double-precision restatement of conv2d backward.  Build + run: tests/test_host_emul.py (nvcc, no GPU).
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#include "../../3d-object-detection.pytorch_b200/csrc/dwc_core.cuh"

using namespace td3d;

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }
template <typename T> static T cvt(double v);
template <> float cvt<float>(double v) { return (float)v; }
template <> bf16 cvt<bf16>(double v) { return __float2bfloat16_rn((float)v); }
static double tof(float v) { return v; }
static double tof(bf16 v) { return __bfloat162float(v); }

struct HostSink {
  std::vector<double>* dwv; std::vector<double>* st; int KK, C;
  void dw(int c, int tap, float v) { (*dwv)[(size_t)c * KK + tap] += v; }
  void stat(int which, int c, float v) { (*st)[(size_t)which * C + c] += v; }
};

static double act_ref(double u, int act) {
  if (act == TD3D_ACT_RELU) return u > 0 ? u : 0;
  if (act == TD3D_ACT_HSWISH) { double r = u + 3; r = r < 0 ? 0 : (r > 6 ? 6 : r); return u * r / 6; }
  if (act == TD3D_ACT_SILU) return u / (1 + exp(-u));
  return u;
}
static double actd_ref(double u, int act) {
  if (act == TD3D_ACT_RELU) return u > 0 ? 1 : 0;
  if (act == TD3D_ACT_HSWISH) return u <= -3 ? 0 : (u >= 3 ? 1 : (2 * u + 3) / 6);
  if (act == TD3D_ACT_SILU) { double s = 1 / (1 + exp(-u)); return s * (1 + u * (1 - s)); }
  return 1;
}

template <typename T, int K, int S, int R, int CPT>
static int run_case(int B, int H, int W, int C, int act, bool with_xf, int item_lanes_req) {
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1, KK = K * K, PAD = (K - 1) / 2;
  std::vector<T> g((size_t)B * Ho * Wo * C), y(g.size()), x((size_t)B * H * W * C), gx(x.size());
  std::vector<float> alpha(B * C), gamma(B * C), beta(C), scale(C), shift(C), se(B * C), taps(KK * C);
  for (auto& v : g) v = cvt<T>(frand());
  for (auto& v : y) v = cvt<T>(frand() * 2);
  for (auto& v : x) v = cvt<T>(frand() * 3);
  for (auto& v : alpha) v = (float)frand();
  for (auto& v : gamma) v = (float)frand() * 0.1f;
  for (auto& v : beta) v = (float)frand() * 0.1f;
  for (auto& v : scale) v = (float)(frand() * 0.5 + 1.0);
  for (auto& v : shift) v = (float)frand() * 0.5f;
  for (auto& v : se) v = (float)(frand() * 0.25 + 0.75);
  for (auto& v : taps) v = (float)frand() * 0.3f;
  for (auto& v : gx) v = cvt<T>(777.0);
  DwcArgs a;
  a.g = g.data(); a.y_out = y.data(); a.alpha = alpha.data(); a.beta = beta.data(); a.gamma = gamma.data();
  a.x = x.data(); a.scale = with_xf ? scale.data() : nullptr; a.shift = with_xf ? shift.data() : nullptr;
  a.se = with_xf ? se.data() : nullptr; a.act = act; a.w_taps = taps.data(); a.gx = gx.data();
  std::vector<float> dummy_stats(1);
  a.stats = dummy_stats.data(); a.dw = nullptr;
  a.B = B; a.H = H; a.W = W; a.C = C; a.Ho = Ho; a.Wo = Wo; a.slots = B;
  a.n_bands = (Ho + R - 1) / R; a.n_items = B * a.n_bands;
  a.item_lanes = item_lanes_req < a.n_items ? item_lanes_req : a.n_items;
  a.cw = 0; a.n_cchunks = 0; a.ilb = 0;
  std::vector<double> dw((size_t)C * KK, 0.0), st(2 * (size_t)C, 0.0);
  HostSink sink = {&dw, &st, KK, C};
  for (int c = 0; c < C; c += CPT)
    for (int il = 0; il < a.item_lanes; ++il) {
      switch (act) {            // the kernels take the activation as a template parameter
        case TD3D_ACT_NONE: DwcBwd<T, K, S, R, CPT, TD3D_ACT_NONE>::thread_main(a, c, il, sink); break;
        case TD3D_ACT_RELU: DwcBwd<T, K, S, R, CPT, TD3D_ACT_RELU>::thread_main(a, c, il, sink); break;
        case TD3D_ACT_HSWISH: DwcBwd<T, K, S, R, CPT, TD3D_ACT_HSWISH>::thread_main(a, c, il, sink); break;
        default: DwcBwd<T, K, S, R, CPT, TD3D_ACT_SILU>::thread_main(a, c, il, sink); break;
      }
    }
  // ---- reference ----
  std::vector<double> gy(g.size()), rgx(x.size(), 0.0), rdw((size_t)C * KK, 0.0), rst(2 * (size_t)C, 0.0);
  for (int b = 0; b < B; ++b)
    for (size_t p = 0; p < (size_t)Ho * Wo; ++p)
      for (int c = 0; c < C; ++c) {
        size_t o = ((size_t)b * Ho * Wo + p) * C + c;
        gy[o] = (double)alpha[b * C + c] * tof(g[o]) + (double)beta[c] * tof(y[o]) + gamma[b * C + c];
      }
  double max_gx = 0, err_gx = 0;
  for (int b = 0; b < B; ++b)
    for (int qy = 0; qy < H; ++qy)
      for (int qx = 0; qx < W; ++qx)
        for (int c = 0; c < C; ++c) {
          size_t o = (((size_t)b * H + qy) * W + qx) * C + c;
          double xr = tof(x[o]);
          double t = with_xf ? (double)se[b * C + c] * ((double)scale[c] * xr + shift[c]) : xr;
          double xa = act_ref(t, act), da = actd_ref(t, act);
          double acc = 0;
          for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
              int ny = qy + PAD - i, nx = qx + PAD - j;
              if (ny < 0 || nx < 0 || ny % S || nx % S) continue;
              int py = ny / S, px = nx / S;
              if (py >= Ho || px >= Wo) continue;
              double gv = gy[(((size_t)b * Ho + py) * Wo + px) * C + c];
              acc += (double)taps[(i * K + j) * C + c] * gv;
              rdw[(size_t)c * KK + i * K + j] += xa * gv;
            }
          rgx[o] = acc * da;
          // near an activation-derivative jump either branch is legitimate
          bool safe = act != TD3D_ACT_HSWISH || (fabs(fabs(t) - 3.0) > 1e-3);
          if (act == TD3D_ACT_RELU && fabs(t) < 1e-4) safe = false;
          double got = tof(gx[o]);
          if (safe) {
            double e = fabs(got - rgx[o]);
            if (e > err_gx) err_gx = e;
          }
          if (fabs(rgx[o]) > max_gx) max_gx = fabs(rgx[o]);
          rst[c] += got;
          rst[C + c] += got * xr;
        }
  double err_dw = 0, max_dw = 0, err_st = 0, max_st = 0;
  for (size_t i = 0; i < dw.size(); ++i) { err_dw = fmax(err_dw, fabs(dw[i] - rdw[i])); max_dw = fmax(max_dw, fabs(rdw[i])); }
  for (size_t i = 0; i < st.size(); ++i) { err_st = fmax(err_st, fabs(st[i] - rst[i])); max_st = fmax(max_st, fabs(rst[i])); }
  const bool is_bf = sizeof(T) == 2;
  const double tol_gx = is_bf ? 6e-3 : 2e-5, tol_dw = is_bf ? 2e-3 : 2e-4, tol_st = 2e-4;
  const bool ok = err_gx / max_gx < tol_gx && err_dw / max_dw < tol_dw && err_st / fmax(max_st, 1e-9) < tol_st;
  printf("%s T=%s K=%d S=%d R=%d CPT=%d B=%d %dx%d C=%d act=%d xf=%d lanes=%d: gx %.2e dw %.2e st %.2e\n", ok ? "ok  " : "FAIL",
         is_bf ? "bf16" : "f32", K, S, R, CPT, B, H, W, C, act, (int)with_xf, a.item_lanes, err_gx / max_gx, err_dw / max_dw,
         err_st / fmax(max_st, 1e-9));
  return ok ? 0 : 1;
}

template <typename T>
static int run_all() {
  int bad = 0;
  const int shapes[][4] = {{2, 14, 14, 16}, {3, 7, 7, 24}, {2, 29, 23, 8}, {1, 56, 56, 8}, {2, 1, 5, 16}, {3, 2, 3, 8}, {1, 8, 8, 8}, {2, 5, 1, 8}};
  for (auto& s : shapes) {
    for (int act = 0; act < 4; ++act) {
      const bool xf = act != 0;
      bad += run_case<T, 3, 1, 4, 2>(s[0], s[1], s[2], s[3], act, xf, 5);
      bad += run_case<T, 3, 1, 2, 2>(s[0], s[1], s[2], s[3], act, xf, 1000);
      bad += run_case<T, 3, 2, 2, 2>(s[0], s[1], s[2], s[3], act, xf, 3);
      bad += run_case<T, 3, 2, 1, 2>(s[0], s[1], s[2], s[3], act, xf, 1000);
      bad += run_case<T, 5, 1, 2, 1>(s[0], s[1], s[2], s[3], act, xf, 4);
      bad += run_case<T, 5, 2, 1, 1>(s[0], s[1], s[2], s[3], act, xf, 1000);
      bad += run_case<T, 5, 1, 1, 2>(s[0], s[1], s[2], s[3], act, xf, 7);
    }
  }
  return bad;
}


// ---- forward ----------------------------------------------------------------------------------------------------
struct HostFwdSink {
  std::vector<double>* st; int C;
  void stat(int b, int which, int c, float v) { (*st)[((size_t)b * 2 + which) * C + c] += v; }
};

template <typename T, int K, int S, int R, int CPT>
static int run_fwd_case(int B, int H, int W, int C, int act, bool with_xf, int out_act, bool with_bias, int lanes_req) {
  const int Ho = (H - 1) / S + 1, Wo = (W - 1) / S + 1, KK = K * K, PAD = (K - 1) / 2;
  std::vector<T> x((size_t)B * H * W * C), y((size_t)B * Ho * Wo * C);
  std::vector<float> scale(C), shift(C), se(B * C), taps(KK * C), bias(C);
  for (auto& v : x) v = cvt<T>(frand() * 3);
  for (auto& v : scale) v = (float)(frand() * 0.5 + 1.0);
  for (auto& v : shift) v = (float)frand() * 0.5f;
  for (auto& v : se) v = (float)(frand() * 0.25 + 0.75);
  for (auto& v : taps) v = (float)frand() * 0.3f;
  for (auto& v : bias) v = (float)frand();
  for (auto& v : y) v = cvt<T>(777.0);
  DwcFwdArgs a;
  a.x = x.data(); a.scale = with_xf ? scale.data() : nullptr; a.shift = with_xf ? shift.data() : nullptr;
  a.se = with_xf ? se.data() : nullptr; a.act = act; a.w_taps = taps.data();
  a.out_bias = with_bias ? bias.data() : nullptr; a.out_act = out_act; a.y = y.data();
  std::vector<float> dummy(1);
  a.stats = dummy.data();
  a.B = B; a.H = H; a.W = W; a.C = C; a.Ho = Ho; a.Wo = Wo;
  a.n_bands = (Ho + R - 1) / R; a.n_items = B * a.n_bands;
  a.item_lanes = lanes_req < a.n_items ? lanes_req : a.n_items;
  a.cw = 0; a.n_cchunks = 0; a.ilb = 0;
  std::vector<double> st((size_t)B * 2 * C, 0.0), rst((size_t)B * 2 * C, 0.0);
  HostFwdSink sink = {&st, C};
  for (int c = 0; c < C; c += CPT)
    for (int il = 0; il < a.item_lanes; ++il) {
      // (input activation, output activation) pairs the launcher instantiates: (X, none) and (none, Y)
      if (out_act == TD3D_ACT_NONE) {
        switch (act) {
          case TD3D_ACT_NONE: DwcFwd<T, K, S, R, CPT, TD3D_ACT_NONE, TD3D_ACT_NONE>::thread_main(a, c, il, sink); break;
          case TD3D_ACT_RELU: DwcFwd<T, K, S, R, CPT, TD3D_ACT_RELU, TD3D_ACT_NONE>::thread_main(a, c, il, sink); break;
          case TD3D_ACT_HSWISH: DwcFwd<T, K, S, R, CPT, TD3D_ACT_HSWISH, TD3D_ACT_NONE>::thread_main(a, c, il, sink); break;
          default: DwcFwd<T, K, S, R, CPT, TD3D_ACT_SILU, TD3D_ACT_NONE>::thread_main(a, c, il, sink); break;
        }
      } else if (out_act == TD3D_ACT_HSWISH) {
        DwcFwd<T, K, S, R, CPT, TD3D_ACT_NONE, TD3D_ACT_HSWISH>::thread_main(a, c, il, sink);
      } else if (out_act == TD3D_ACT_SILU) {
        DwcFwd<T, K, S, R, CPT, TD3D_ACT_NONE, TD3D_ACT_SILU>::thread_main(a, c, il, sink);
      } else {
        DwcFwd<T, K, S, R, CPT, TD3D_ACT_NONE, TD3D_ACT_RELU>::thread_main(a, c, il, sink);
      }
    }
  double err = 0, mx = 0;
  for (int b = 0; b < B; ++b)
    for (int py = 0; py < Ho; ++py)
      for (int px = 0; px < Wo; ++px)
        for (int c = 0; c < C; ++c) {
          double acc = with_bias ? bias[c] : 0.0;
          for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) {
              int qy = S * py + i - PAD, qx = S * px + j - PAD;
              if (qy < 0 || qx < 0 || qy >= H || qx >= W) continue;
              double xr = tof(x[(((size_t)b * H + qy) * W + qx) * C + c]);
              double t = with_xf ? (double)se[b * C + c] * ((double)scale[c] * xr + shift[c]) : xr;
              acc += (double)taps[(i * K + j) * C + c] * act_ref(t, act);
            }
          double ref = act_ref(acc, out_act);
          double got = tof(y[(((size_t)b * Ho + py) * Wo + px) * C + c]);
          err = fmax(err, fabs(got - ref)); mx = fmax(mx, fabs(ref));
          rst[((size_t)b * 2) * C + c] += got;
          rst[((size_t)b * 2 + 1) * C + c] += got * got;
        }
  double err_st = 0, max_st = 0;
  for (size_t i = 0; i < st.size(); ++i) { err_st = fmax(err_st, fabs(st[i] - rst[i])); max_st = fmax(max_st, fabs(rst[i])); }
  const bool is_bf = sizeof(T) == 2;
  const bool ok = err / mx < (is_bf ? 6e-3 : 2e-5) && err_st / fmax(max_st, 1e-9) < 2e-4;
  printf("%s FWD T=%s K=%d S=%d R=%d CPT=%d B=%d %dx%d C=%d act=%d xf=%d oact=%d bias=%d lanes=%d: y %.2e st %.2e\n", ok ? "ok  " : "FAIL",
         is_bf ? "bf16" : "f32", K, S, R, CPT, B, H, W, C, act, (int)with_xf, out_act, (int)with_bias, a.item_lanes, err / mx,
         err_st / fmax(max_st, 1e-9));
  return ok ? 0 : 1;
}

template <typename T>
static int run_all_fwd() {
  int bad = 0;
  const int shapes[][4] = {{2, 14, 14, 16}, {3, 7, 7, 24}, {2, 29, 23, 8}, {1, 56, 56, 8}, {2, 1, 5, 16}, {3, 2, 3, 8}, {1, 8, 8, 8}, {2, 5, 1, 8}};
  for (auto& s : shapes) {
    for (int v = 0; v < 5; ++v) {
      // v = 0: inference epilogue (bias + h-swish); 1, 2: training inputs (ReLU / h-swish on load); 3: SiLU input;
      // 4: inference epilogue with SiLU
      const int act = v == 4 ? 0 : v, oact = v == 0 ? 2 : (v == 4 ? 3 : 0);
      const bool xf = act != 0, bias = oact != 0;
      bad += run_fwd_case<T, 3, 1, 4, 2>(s[0], s[1], s[2], s[3], act, xf, oact, bias, 5);
      bad += run_fwd_case<T, 3, 1, 2, 2>(s[0], s[1], s[2], s[3], act, xf, oact, bias, 1000);
      bad += run_fwd_case<T, 3, 2, 2, 2>(s[0], s[1], s[2], s[3], act, xf, oact, bias, 3);
      bad += run_fwd_case<T, 3, 2, 1, 2>(s[0], s[1], s[2], s[3], act, xf, oact, bias, 1000);
      bad += run_fwd_case<T, 5, 1, 2, 2>(s[0], s[1], s[2], s[3], act, xf, oact, bias, 4);
      bad += run_fwd_case<T, 5, 2, 2, 2>(s[0], s[1], s[2], s[3], act, xf, oact, bias, 1000);
      bad += run_fwd_case<T, 5, 2, 1, 1>(s[0], s[1], s[2], s[3], act, xf, oact, bias, 7);
    }
  }
  return bad;
}

int main() {
  srand(1234);
  int bad = run_all<float>() + run_all<bf16>() + run_all_fwd<float>() + run_all_fwd<bf16>();
  printf("%s (%d failing cases)\n", bad ? "DWC_EMUL FAILED" : "DWC_EMUL OK", bad);
  return bad ? 1 : 0;
}
