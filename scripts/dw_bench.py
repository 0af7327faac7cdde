#!/usr/bin/env python
"""Per-layer depthwise micro-benchmark (MobileNetV3-large shapes, batch 256, bf16): row-walker / tiled kernels vs the
column walker, forward and backward, CUDA-event timed.   python scripts/dw_bench.py [batch] > table"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-object-detection.pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from torchdet3d_b200 import _lib as L  # noqa: E402
import _k as K  # noqa: E402

# (H, C, k, stride) of the 15 depthwise layers of mobilenetv3_large at 224x224
LAYERS = [(112, 16, 3, 1), (112, 64, 3, 2), (56, 72, 3, 1), (56, 72, 5, 2), (28, 120, 5, 1), (28, 120, 5, 1), (28, 240, 3, 2),
          (14, 200, 3, 1), (14, 184, 3, 1), (14, 184, 3, 1), (14, 480, 3, 1), (14, 672, 3, 1), (14, 672, 5, 2), (7, 960, 5, 1),
          (7, 960, 5, 1)]


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    only = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
    L.require_b200()
    dev, code = "cuda", L.BF16
    tot = [0.0] * 4
    print(f"{'layer':28s} {'fwd old us':>10s} {'GB/s':>7s} {'fwd cw us':>10s} {'GB/s':>7s} {'fwd plan us':>11s} {'GB/s':>7s} {'bwd us':>12s} {'GB/s':>7s}")
    for i, (H, C, k, s) in enumerate(LAYERS):
        if only and i not in only:
            continue
        Ho = (H - 1) // s + 1
        x = (torch.randn(B, H, H, C, device=dev) * 2).bfloat16()
        scale, shift = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.5
        taps = (torch.randn(k * k, C, device=dev) * 0.3).contiguous()
        y, _ = K.dw_fwd_ex(x, scale, shift, None, L.ACT_HSWISH, taps, k, s, code, 1)
        g = torch.randn_like(y.float()).bfloat16()
        alpha, gamma, beta = torch.randn(B, C, device=dev), torch.randn(B, C, device=dev) * 0.1, torch.randn(C, device=dev) * 0.1
        fb = 2.0 * B * C * (H * H + Ho * Ho)
        bb = 2.0 * B * C * (2 * H * H + 2 * Ho * Ho)
        t = [timed(lambda: K.dw_fwd_ex(x, scale, shift, None, L.ACT_HSWISH, taps, k, s, code, 1)),
             timed(lambda: K.dw_fwd_ex(x, scale, shift, None, L.ACT_HSWISH, taps, k, s, code, 2)),
             timed(lambda: K.dw_fwd_ex(x, scale, shift, None, L.ACT_HSWISH, taps, k, s, code, 0)),
             timed(lambda: K.dw_bwd(g, y, alpha, beta, gamma, x, scale, shift, None, L.ACT_HSWISH, taps, k, s, code))]
        for j in range(4):
            tot[j] += t[j]
        print(f"{i + 1:2d} {H:3d}x{H:<3d} C={C:<4d} k={k} s={s}     {t[0]:10.1f} {fb / t[0] / 1e3:7.0f} {t[1]:10.1f} {fb / t[1] / 1e3:7.0f} "
              f"{t[2]:11.1f} {fb / t[2] / 1e3:7.0f} {t[3]:12.1f} {bb / t[3] / 1e3:7.0f}")
    print(f"{'total':28s} {tot[0]:10.1f} {'':7s} {tot[1]:10.1f} {'':7s} {tot[2]:11.1f} {'':7s} {tot[3]:12.1f}")


if __name__ == "__main__":
    main()
