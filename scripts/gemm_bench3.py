#!/usr/bin/env python
"""tcgen05 NT GEMM, 128-row against 256-row tiles (TD3D_TC_MSUB=1 / 2) on the long-M MobileNetV3-large layers (batch 256), with and
without the BatchNorm-statistics epilogue and with the inference epilogue (bias + h-swish)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-object-detection.pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["TD3D_TC_LIVE_ENV"] = "1"

import torch  # noqa: E402

from torchdet3d_b200 import _lib as L  # noqa: E402

L.require_b200()
dev = "cuda"
B = 256
SHAPES = [(B * 12544, 16, 16), (B * 12544, 64, 16), (B * 3136, 24, 64), (B * 3136, 72, 24), (B * 3136, 24, 72), (B * 784, 40, 72),
          (B * 784, 120, 40), (B * 784, 40, 120), (B * 196, 80, 240), (B * 196, 112, 480), (B * 3136, 64, 64), (B * 784, 64, 128)]


def run(M, N, Kd, slots, msub):
    os.environ["TD3D_TC_MSUB"] = str(msub)
    a = torch.randn(M, Kd, device=dev).bfloat16()
    w = torch.randn(N, Kd, device=dev).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(slots, 2, N, device=dev) if slots else None

    def call():
        L.check(L.lib().td3d_k_gemm_nt(L.ptr(a), L.ptr(w), L.ptr(y), None, None, None, L.ptr(st), slots, M, N, Kd, L.BF16, 0,
                                       L.GEMM_TCGEN05, L.stream()))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    return us, (M * Kd + M * N) * 2 / 1e3 / us


print(f"{'M':>9s} {'N':>4s} {'K':>4s} {'stats':>5s} | {'128-row us':>10s} {'GB/s':>6s} | {'256-row us':>10s} {'GB/s':>6s} | {'default us':>10s}")
SHAPES += [(B * 3136, 72, 24), (B * 784, 120, 40), (B * 784, 240, 40), (B * 196, 480, 80), (B * 196, 672, 112), (B * 49, 960, 160)]
for M, N, Kd in SHAPES:
    for slots in (32, 0):
        u1, g1 = run(M, N, Kd, slots, 1)
        u2, g2 = run(M, N, Kd, slots, 2) if N <= 128 else (float("nan"), float("nan"))
        u3, _ = run(M, N, Kd, slots, 0)
        print(f"{M:9d} {N:4d} {Kd:4d} {slots:5d} | {u1:10.1f} {g1:6.0f} | {u2:10.1f} {g2:6.0f} | {u3:10.1f}", flush=True)
os.environ["TD3D_TC_MSUB"] = "0"
