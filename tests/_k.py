"""Thin test-side wrappers around the per-kernel C-ABI entry points (td3d_k_*)."""
import ctypes as C

import torch

from torchdet3d_b200 import _lib as L


def _f(x):
    return C.c_float(float(x))


def dt(code):
    return torch.bfloat16 if code == L.BF16 else torch.float32


def nhwc(x_nchw, code):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(dt(code))


def nchw(x_nhwc):
    return x_nhwc.float().permute(0, 3, 1, 2).contiguous()


def stats_buf(B, Cn, dev):
    return torch.zeros(B, 2, Cn, device=dev)


def apply_xform(y, scale, shift, se, act, res, code, want_out=True, pool=False):
    B, H, W, Cn = y.shape
    out = torch.empty_like(y) if want_out else None
    st = stats_buf(B, Cn, y.device) if pool else None
    L.check(L.lib().td3d_k_apply_xform(L.ptr(y), L.ptr(scale), L.ptr(shift), L.ptr(se), act, L.ptr(res), L.ptr(out),
                                       L.ptr(st), B, H * W, Cn, code, L.stream()))
    return out, st


def affine2(g, y, alpha, beta, gamma, code):
    B, H, W, Cn = y.shape
    out = torch.empty_like(y)
    L.check(L.lib().td3d_k_affine2(L.ptr(g), L.ptr(y), L.ptr(alpha), L.ptr(beta), L.ptr(gamma), L.ptr(out), B, H * W, Cn,
                                   code, L.stream()))
    return out


def act_bwd_stats(g, y, scale, shift, se, act, code):
    B, H, W, Cn = y.shape
    gu = torch.empty_like(y)
    st = stats_buf(B, Cn, y.device)
    L.check(L.lib().td3d_k_act_bwd_stats(L.ptr(g), L.ptr(y), L.ptr(scale), L.ptr(shift), L.ptr(se), act, L.ptr(gu),
                                         L.ptr(st), B, H * W, Cn, code, L.stream()))
    return gu, st


def stem_fwd(img, w27xC, code):
    B, _, H, W = img.shape
    Cn = w27xC.shape[1]
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = torch.empty(B, Ho, Wo, Cn, device=img.device, dtype=dt(code))
    st = stats_buf(B, Cn, img.device)
    L.check(L.lib().td3d_k_stem_fwd(L.ptr(img), L.ptr(w27xC), L.ptr(y), L.ptr(st), B, H, W, Cn, code, L.stream()))
    return y, st


def stem_wgrad(img, g, y, alpha, beta, gamma, code):
    B, _, H, W = img.shape
    Cn = y.shape[3]
    dw = torch.zeros(Cn, 3, 3, 3, device=img.device)
    L.check(L.lib().td3d_k_stem_wgrad(L.ptr(img), L.ptr(g), L.ptr(y), L.ptr(alpha), L.ptr(beta), L.ptr(gamma), L.ptr(dw),
                                      B, H, W, Cn, code, L.stream()))
    return dw


def dw_fwd(x, scale, shift, se, act, w_taps, k, stride, code):
    B, H, W, Cn = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    y = torch.empty(B, Ho, Wo, Cn, device=x.device, dtype=dt(code))
    st = stats_buf(B, Cn, x.device)
    L.check(L.lib().td3d_k_dw_fwd(L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(se), act, L.ptr(w_taps), L.ptr(y), L.ptr(st),
                                  B, H, W, Cn, k, stride, code, L.stream()))
    return y, st


def dw_fwd_ex(x, scale, shift, se, act, w_taps, k, stride, code, impl, out_bias=None, out_act=0):
    B, H, W, Cn = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    y = torch.empty(B, Ho, Wo, Cn, device=x.device, dtype=dt(code))
    st = stats_buf(B, Cn, x.device)
    L.check(L.lib().td3d_k_dw_fwd_ex(L.ptr(x), L.ptr(scale), L.ptr(shift), L.ptr(se), act, L.ptr(w_taps), L.ptr(out_bias),
                                     out_act, L.ptr(y), L.ptr(st), B, H, W, Cn, k, stride, code, impl, L.stream()))
    return y, st


def dw_bwd(g, y_out, alpha, beta, gamma, x, scale, shift, se, act, w_taps, k, stride, code):
    B, H, W, Cn = x.shape
    gx = torch.empty_like(x)
    dw = torch.zeros(Cn, 1, k, k, device=x.device)
    st = stats_buf(B, Cn, x.device)
    L.check(L.lib().td3d_k_dw_bwd(L.ptr(g), L.ptr(y_out), L.ptr(alpha), L.ptr(beta), L.ptr(gamma), L.ptr(x), L.ptr(scale),
                                  L.ptr(shift), L.ptr(se), act, L.ptr(w_taps), L.ptr(gx), L.ptr(dw), L.ptr(st), B, H, W, Cn,
                                  k, stride, code, L.stream()))
    return gx, dw, st


def gemm_nt(a, w, code, impl, addend=None, bias=None, ysaved=None, slots=0, out_f32=False):
    M, K = a.shape
    N = w.shape[0]
    y = torch.empty(M, N, device=a.device, dtype=torch.float32 if out_f32 else dt(code))
    st = torch.zeros(slots, 2, N, device=a.device) if slots else None
    L.check(L.lib().td3d_k_gemm_nt(L.ptr(a), L.ptr(w), L.ptr(y), L.ptr(addend), L.ptr(bias), L.ptr(ysaved), L.ptr(st), slots,
                                   M, N, K, code, 1 if out_f32 else 0, impl, L.stream()))
    return y, st


def gemm_tn(a, b, code, impl):
    M, N1 = a.shape
    N2 = b.shape[1]
    c = torch.zeros(N1, N2, device=a.device)
    L.check(L.lib().td3d_k_gemm_tn(L.ptr(a), L.ptr(b), L.ptr(c), M, N1, N2, code, impl, L.stream()))
    return c


def act_ref(u, act):
    if act == L.ACT_RELU:
        return torch.relu(u)
    if act == L.ACT_HSWISH:
        return u * torch.nn.functional.relu6(u + 3) / 6
    if act == L.ACT_SILU:
        return torch.nn.functional.silu(u)
    return u
