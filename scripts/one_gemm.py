#!/usr/bin/env python
"""A few launches of the tcgen05 NT GEMM on one shape (for `ncu --set full`):  one_gemm.py M N K slots"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-object-detection.pytorch_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

from torchdet3d_b200 import _lib as L  # noqa: E402

L.require_b200()
M, N, Kd, slots = (int(v) for v in sys.argv[1:5])
dev = "cuda"
a = torch.randn(M, Kd, device=dev).bfloat16()
w = torch.randn(N, Kd, device=dev).bfloat16()
y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
st = torch.zeros(slots, 2, N, device=dev) if slots else None
for _ in range(3):
    L.check(L.lib().td3d_k_gemm_nt(L.ptr(a), L.ptr(w), L.ptr(y), None, None, None, L.ptr(st), slots, M, N, Kd, L.BF16, 0,
                                   L.GEMM_TCGEN05, L.stream()))
torch.cuda.synchronize()
