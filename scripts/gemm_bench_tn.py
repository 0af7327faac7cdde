#!/usr/bin/env python
"""tcgen05 TN (weight-gradient) GEMM on the MobileNetV3-large layer shapes, batch 256: dW[N1,N2] += A[M,N1]^T B[M,N2]."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-object-detection.pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from torchdet3d_b200 import _lib as L  # noqa: E402

L.require_b200()
dev = "cuda"
B = 256
SHAPES = [(B * 12544, 16, 16), (B * 12544, 64, 16), (B * 3136, 24, 64), (B * 3136, 72, 24), (B * 3136, 24, 72), (B * 784, 40, 72),
          (B * 784, 120, 40), (B * 784, 40, 120), (B * 784, 240, 40), (B * 196, 80, 240), (B * 196, 200, 80), (B * 196, 480, 80),
          (B * 196, 112, 480), (B * 196, 672, 112), (B * 49, 160, 672), (B * 49, 960, 160), (B * 49, 160, 960)]
print(f"{'M':>9s} {'N1':>4s} {'N2':>4s} | {'us':>8s} {'GB/s':>6s}")
tot = 0.0
for M, N1, N2 in SHAPES:
    a = torch.randn(M, N1, device=dev).bfloat16()
    b = torch.randn(M, N2, device=dev).bfloat16()
    c = torch.zeros(N1, N2, device=dev)

    def call():
        L.check(L.lib().td3d_k_gemm_tn(L.ptr(a), L.ptr(b), L.ptr(c), M, N1, N2, L.BF16, L.GEMM_TCGEN05, L.stream()))
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
    tot += us
    print(f"{M:9d} {N1:4d} {N2:4d} | {us:8.1f} {(M * (N1 + N2)) * 2 / 1e3 / us:6.0f}", flush=True)
print(f"total {tot:.1f} us")
