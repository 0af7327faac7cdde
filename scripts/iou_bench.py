#!/usr/bin/env python
"""2D-keypoint based 3D IoU: the CUDA kernel (one launch per batch, pairs/s with the keypoints resident on the device and
end to end from host tensors) next to the oracle port of the reference's per-sample CPU loop on a 256-pair sample."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "3d-object-detection.pytorch_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import iou_port  # noqa: E402
from torchdet3d_b200 import _lib as L  # noqa: E402
from torchdet3d_b200.evaluation import compute_2d_based_iou  # noqa: E402

L.require_b200()
gold = np.load(os.path.join(ROOT, "tests", "golden", "iou.npz"))
base = gold["gt"][0].astype(np.float64)
rng = np.random.default_rng(0)
for n in (128, 512, 4096, 32768):
    gt = np.clip(base + rng.normal(0, 0.03, (n,) + base.shape), 0, 1).astype(np.float32)
    pred = np.clip(gt + rng.normal(0, 0.02, gt.shape), 0, 1).astype(np.float32)
    p, g = torch.as_tensor(pred, device="cuda"), torch.as_tensor(gt, device="cuda")
    out = torch.empty(n, dtype=torch.float64, device="cuda")

    def launch():
        L.check(L.lib().td3d_iou_2d_based(L.ptr(p), L.ptr(g), n, 1, None, L.ptr(out), L.stream()))
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        launch()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ph, gh = torch.as_tensor(pred).pin_memory(), torch.as_tensor(gt).pin_memory()
    t0 = time.perf_counter()
    for _ in range(5):
        v = compute_2d_based_iou(ph.cuda(non_blocking=True), gh.cuda(non_blocking=True))
    e2e_ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"n={n:6d}  kernel {ms:8.3f} ms  {n / ms * 1e3:12.0f} pairs/s   end to end {e2e_ms:8.3f} ms  {n / e2e_ms * 1e3:12.0f} pairs/s   mean IoU {v:.4f}", flush=True)
m = 256
t0 = time.perf_counter()
ref = iou_port.compute_2d_based_iou(pred[:m], gt[:m])
sec = time.perf_counter() - t0
print(f"oracle port of the reference loop (numpy eigh + scipy Qhull, 1 core): {m} pairs in {sec * 1e3:.1f} ms = {m / sec:.0f} pairs/s   mean IoU {ref:.4f}")
