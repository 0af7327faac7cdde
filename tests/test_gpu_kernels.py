"""GPU parity of the individual sm_100a kernels (through the C ABI) against plain PyTorch fp32
references of the same op.  fp32 kernels: tight tolerance; bf16 storage: tolerance of the
rounding of inputs/outputs (2^-8 relative), stated per test."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from torchdet3d_b200 import _lib as L   # noqa: E402
import _k as K                           # noqa: E402

DEV = "cuda"
CODES = [L.F32, L.BF16]


def tol(code):
    return dict(rtol=2e-5, atol=2e-5) if code == L.F32 else dict(rtol=2e-2, atol=2e-2)


def rel_err(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)
    L.require_b200()


@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("shape", [(3, 7, 7, 96), (2, 28, 28, 72), (5, 1, 1, 1024), (2, 56, 56, 16), (1, 5, 9, 2096)])
@pytest.mark.parametrize("act", [L.ACT_NONE, L.ACT_RELU, L.ACT_HSWISH, L.ACT_SILU])
def test_apply_xform_and_pool(code, shape, act):
    B, H, W, Cn = shape
    y = (torch.randn(shape, device=DEV) * 2).to(K.dt(code))
    res = torch.randn(shape, device=DEV).to(K.dt(code))
    scale, shift = torch.rand(Cn, device=DEV) + 0.5, torch.randn(Cn, device=DEV)
    se = torch.rand(B, Cn, device=DEV)
    out, st = K.apply_xform(y, scale, shift, se, act, res, code, pool=True)
    ref = K.act_ref(se[:, None, None, :] * (y.float() * scale + shift), act) + res.float()
    torch.testing.assert_close(out.float(), ref, **tol(code))
    torch.testing.assert_close(st[:, 0], out.float().sum(dim=(1, 2)), rtol=1e-4, atol=1e-3)
    # no transform, no residual, pooled only (out = NULL)
    _, st2 = K.apply_xform(y, None, None, None, act, None, code, want_out=False, pool=True)
    torch.testing.assert_close(st2[:, 0], K.act_ref(y.float(), act).sum(dim=(1, 2)), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("shape", [(3, 7, 7, 96), (2, 28, 28, 72), (4, 1, 1, 1280)])
def test_affine2_and_act_bwd_stats(code, shape):
    B, H, W, Cn = shape
    g = torch.randn(shape, device=DEV).to(K.dt(code))
    y = (torch.randn(shape, device=DEV) * 2).to(K.dt(code))
    alpha, gamma, beta = torch.randn(B, Cn, device=DEV), torch.randn(B, Cn, device=DEV), torch.randn(Cn, device=DEV)
    out = K.affine2(g, y, alpha, beta, gamma, code)
    ref = alpha[:, None, None] * g.float() + beta * y.float() + gamma[:, None, None]
    torch.testing.assert_close(out.float(), ref, **tol(code))
    scale, shift = torch.rand(Cn, device=DEV) + 0.5, torch.randn(Cn, device=DEV)
    se = torch.rand(B, Cn, device=DEV) + 0.2
    for act in (L.ACT_NONE, L.ACT_RELU, L.ACT_HSWISH, L.ACT_SILU):
        gu, st = K.act_bwd_stats(g, y, scale, shift, se, act, code)
        u = (se[:, None, None] * (y.float() * scale + shift)).requires_grad_(True)
        K.act_ref(u, act).backward(g.float())
        torch.testing.assert_close(gu.float(), u.grad, **tol(code))
        torch.testing.assert_close(st[:, 0], gu.float().sum(dim=(1, 2)), rtol=1e-4, atol=2e-3)
        torch.testing.assert_close(st[:, 1], (gu.float() * y.float()).sum(dim=(1, 2)), rtol=1e-4, atol=5e-3)


@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("hw", [(224, 224), (64, 64), (37, 51)])
@pytest.mark.parametrize("Cs", [16, 32, 40])          # MobileNetV3 / EfficientNet-B0 / EfficientNet-B3 stems
def test_stem(code, hw, Cs):
    B = 3
    img = torch.rand(B, 3, *hw, device=DEV)
    w = torch.randn(Cs, 3, 3, 3, device=DEV) * 0.3
    w27 = w.reshape(Cs, 27).t().contiguous()
    y, st = K.stem_fwd(img, w27, code)
    ref = F.conv2d(img, w, stride=2, padding=1)
    assert rel_err(K.nchw(y), ref) < (1e-5 if code == L.F32 else 6e-3)
    yf = y.float()
    torch.testing.assert_close(st[:, 0], yf.sum(dim=(1, 2)), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(st[:, 1], (yf * yf).sum(dim=(1, 2)), rtol=1e-4, atol=1e-2)
    # weight gradient with a lazily applied affine (BN backward) on the incoming gradient
    g = torch.randn_like(yf).to(K.dt(code))
    alpha, gamma, beta = torch.randn(B, Cs, device=DEV), torch.randn(B, Cs, device=DEV) * 0.1, torch.randn(Cs, device=DEV) * 0.1
    dw = K.stem_wgrad(img, g, y, alpha, beta, gamma, code)
    gy = alpha[:, None, None] * g.float() + beta * yf + gamma[:, None, None]
    wr = w.clone().requires_grad_(True)
    F.conv2d(img, wr, stride=2, padding=1).backward(gy.permute(0, 3, 1, 2))
    assert rel_err(dw, wr.grad) < 2e-4


@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (5, 1), (5, 2)])
@pytest.mark.parametrize("shape", [(2, 14, 14, 240), (3, 7, 7, 96), (2, 29, 23, 16), (1, 56, 56, 72)])
def test_depthwise(code, k, stride, shape, gx_tol_bf16=1e-2, act=L.ACT_HSWISH):
    B, H, W, Cn = shape
    x = (torch.randn(shape, device=DEV) * 2).to(K.dt(code))
    scale, shift = torch.rand(Cn, device=DEV) + 0.5, torch.randn(Cn, device=DEV) * 0.5
    se = torch.rand(B, Cn, device=DEV) + 0.5
    w = torch.randn(Cn, 1, k, k, device=DEV) * 0.3
    taps = w.reshape(Cn, k * k).t().contiguous()
    y, st = K.dw_fwd(x, scale, shift, se, act, taps, k, stride, code)
    xt = (se[:, None, None] * (x.float() * scale + shift)).requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    ref = F.conv2d(K.act_ref(xt, act).permute(0, 3, 1, 2), wr, stride=stride, padding=(k - 1) // 2, groups=Cn)
    assert rel_err(K.nchw(y), ref) < (1e-5 if code == L.F32 else 6e-3)
    yf = y.float()
    torch.testing.assert_close(st[:, 0], yf.sum(dim=(1, 2)), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(st[:, 1], (yf * yf).sum(dim=(1, 2)), rtol=1e-4, atol=1e-2)
    # backward: gy = alpha*g + beta*y + gamma
    g = torch.randn_like(yf).to(K.dt(code))
    alpha, gamma, beta = torch.randn(B, Cn, device=DEV), torch.randn(B, Cn, device=DEV) * 0.1, torch.randn(Cn, device=DEV) * 0.1
    gx, dw, bst = K.dw_bwd(g, y, alpha, beta, gamma, x, scale, shift, se, act, taps, k, stride, code)
    gy = alpha[:, None, None] * g.float() + beta * yf + gamma[:, None, None]
    ref.backward(gy.permute(0, 3, 1, 2))
    # h_swish' jumps at |u| = 3: an element whose pre-activation lies within rounding of the jump legitimately takes either
    # branch (seed 0 puts one at u = 3 - 1ulp for channel 5), so those elements are left out of the comparison
    safe = ((xt.detach().abs() - 3.0).abs() > 1e-4).float()
    assert rel_err(gx.float() * safe, xt.grad * safe) < (2e-5 if code == L.F32 else gx_tol_bf16)
    assert rel_err(dw, wr.grad) < (2e-4 if code == L.F32 else 2e-2)
    # the backward sums feed a BatchNorm backward over the whole batch: only their total over the slots is defined
    gxf = gx.float()
    torch.testing.assert_close(bst[:, 0].sum(0), gxf.sum(dim=(0, 1, 2)), rtol=1e-4, atol=2e-2 * B ** 0.5)
    torch.testing.assert_close(bst[:, 1].sum(0), (gxf * x.float()).sum(dim=(0, 1, 2)), rtol=1e-4, atol=5e-2 * B ** 0.5)


# row-walker edge cases (k_dww.cu): rows that fill all lanes of their segment (shuffle wrap needs the select path),
# single-row planes, channel spans that end inside a warp, more samples than sample lanes in the grid
@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("k", [3, 5])
@pytest.mark.parametrize("shape", [(5, 8, 8, 40), (3, 32, 32, 24), (2, 1, 5, 16), (9, 16, 16, 8), (300, 7, 7, 24), (2, 3, 31, 48)])
def test_depthwise_walker_edges(code, k, shape):
    test_depthwise(code, k, 1, shape)


# SiLU inputs (EfficientNet) take the column-walker forward kernel (k_dwc.cu) instead of the row walker / tiled kernels
@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (5, 1), (5, 2)])
@pytest.mark.parametrize("shape", [(2, 14, 14, 240), (3, 7, 7, 96), (2, 29, 23, 16), (1, 56, 56, 72), (2, 1, 5, 16), (3, 2, 3, 8)])
def test_depthwise_silu_column_walker(code, k, stride, shape):
    test_depthwise(code, k, stride, shape, act=L.ACT_SILU)


# both forward implementations and the inference epilogue (bias + activation) of the column-walker forward, explicitly
@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("k,stride", [(3, 1), (3, 2), (5, 1), (5, 2)])
@pytest.mark.parametrize("shape", [(2, 14, 14, 240), (3, 7, 7, 96), (2, 29, 23, 16), (1, 56, 56, 72)])
def test_depthwise_explicit_implementations(code, k, stride, shape):
    B, H, W, Cn = shape
    x = (torch.randn(shape, device=DEV) * 2).to(K.dt(code))
    scale, shift = torch.rand(Cn, device=DEV) + 0.5, torch.randn(Cn, device=DEV) * 0.5
    w = torch.randn(Cn, 1, k, k, device=DEV) * 0.3
    taps = w.reshape(Cn, k * k).t().contiguous()
    bias = torch.randn(Cn, device=DEV)
    # inference epilogue: y = h_swish(dw(x) + bias)
    y, st = K.dw_fwd_ex(x, None, None, None, L.ACT_NONE, taps, k, stride, code, 2, out_bias=bias, out_act=L.ACT_HSWISH)
    ref = K.act_ref(F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=stride, padding=(k - 1) // 2, groups=Cn) + bias[None, :, None, None], L.ACT_HSWISH)
    assert rel_err(K.nchw(y), ref) < (1e-5 if code == L.F32 else 6e-3)
    torch.testing.assert_close(st[:, 0], y.float().sum(dim=(1, 2)), rtol=1e-4, atol=1e-2)
    # forward: the two implementations agree
    y1, _ = K.dw_fwd_ex(x, scale, shift, None, L.ACT_RELU, taps, k, stride, code, 1)
    y2, _ = K.dw_fwd_ex(x, scale, shift, None, L.ACT_RELU, taps, k, stride, code, 2)
    assert rel_err(y2.float(), y1.float()) < (1e-5 if code == L.F32 else 8e-3)


GEMM_SHAPES = [(300, 64, 16), (1000, 24, 72), (257, 88, 24), (129, 960, 160), (64, 1280, 960), (5000, 16, 64),
               (130, 200, 80), (128, 184, 80), (77, 40, 120)]


@pytest.mark.parametrize("code", CODES)
@pytest.mark.parametrize("M,N,Kd", GEMM_SHAPES)
def test_gemm_simt(code, M, N, Kd):
    a = torch.randn(M, Kd, device=DEV).to(K.dt(code))
    w = (torch.randn(N, Kd, device=DEV) / Kd ** 0.5).to(K.dt(code))
    add = torch.randn(M, N, device=DEV).to(K.dt(code))
    bias = torch.randn(N, device=DEV)
    ys = torch.randn(M, N, device=DEV).to(K.dt(code))
    y, st = K.gemm_nt(a, w, code, L.GEMM_SIMT, slots=7)
    ref = a.float() @ w.float().t()
    assert rel_err(y.float(), ref) < (1e-5 if code == L.F32 else 6e-3)
    yf = y.float()
    torch.testing.assert_close(st.sum(0)[0], yf.sum(0), rtol=1e-4, atol=1e-2)
    torch.testing.assert_close(st.sum(0)[1], (yf * yf).sum(0), rtol=1e-4, atol=1e-2)
    y2, st2 = K.gemm_nt(a, w, code, L.GEMM_SIMT, addend=add, bias=bias, ysaved=ys, slots=3, out_f32=True)
    ref2 = ref + add.float() + bias
    assert rel_err(y2, ref2) < 1e-5
    torch.testing.assert_close(st2.sum(0)[1], (y2 * ys.float()).sum(0), rtol=1e-4, atol=2e-2)
    c = K.gemm_tn(add, a, code, L.GEMM_SIMT)          # [N, Kd] = add^T a
    assert rel_err(c, add.float().t() @ a.float()) < 2e-5
