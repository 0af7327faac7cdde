"""ADD / symmetric-ADD / accuracy on the fused sm_100a metrics kernel.

Drop-in for torchdet3d/evaluation/metrics.py:10-68 (`compute_average_distance`, `compute_accuracy`,
`compute_metrics_per_cls`).  The reference runs a 9x9 Python loop of tiny kernels plus three host
syncs per call; here one launch (one warp per sample) accumulates everything -- totals and the
per-class breakdown -- into a device-side f64 accumulator.  The 3D-IoU branch (EPnP lift + Qhull,
metrics.py:70-89) is CPU numpy/scipy per sample and out of scope; `compute_iou=True` raises.
"""
import torch

from .. import _lib as L

MAX_CLASSES = 9


class MetricAccumulator:
    """Device-side running sums: [sum ADD_b, sum SADD_b, hits, count] + the same per class.
    `update()` launches one kernel and never synchronises; `read()` does one D2H copy."""

    def __init__(self, device, max_classes=MAX_CLASSES):
        self.max_classes = max_classes
        self.acc = torch.zeros(4 + 4 * max_classes, dtype=torch.float64, device=device)

    def reset(self):
        self.acc.zero_()

    def update(self, pred_kp, gt_kp, pred_cats, gt_cats):
        if not pred_kp.is_cuda:
            raise L.Td3dError("td3d metrics run on CUDA (sm_100a) only; there is no CPU fallback")
        B = pred_kp.shape[0]
        if B == 0:
            return
        kp = pred_kp.detach().contiguous().float()
        gt = gt_kp.detach().contiguous().float()
        lg = pred_cats.detach().contiguous().float() if pred_cats is not None else None
        ct = gt_cats.contiguous().to(torch.int64)
        nc = lg.shape[1] if lg is not None else 1
        L.check(L.lib().td3d_metrics_accum(L.ptr(kp), L.ptr(gt), L.ptr(lg), L.ptr(ct), B, nc, self.max_classes,
                                           L.ptr(self.acc), L.stream()))

    def read(self):
        return self.acc.cpu().tolist()


def _run(pred_kp, gt_kp, pred_cats, gt_cats):
    acc = MetricAccumulator(pred_kp.device)
    acc.update(pred_kp, gt_kp, pred_cats, gt_cats)
    return acc.read()


@torch.no_grad()
def compute_average_distance(pred_kp, gt_kp, num_keypoint=9, reduce_mean=True, **kwargs):
    """-> (ADD, symmetric ADD) as Python floats (metrics.py:10-29)."""
    assert num_keypoint == 9
    B = pred_kp.shape[0]
    if B == 0:
        return (float("nan"), float("nan")) if reduce_mean else (0.0, 0.0)
    cats = torch.zeros(B, dtype=torch.int64, device=pred_kp.device)
    a = _run(pred_kp, gt_kp, None, cats)
    if reduce_mean:
        return a[0] / B, a[1] / B
    return a[0], a[1]            # reference: sum(norm)/9 and sum(sym)/9 -- per-sample means summed


@torch.no_grad()
def compute_accuracy(pred_cats, gt_cats, reduce_mean=True, **kwargs):
    B = pred_cats.shape[0]
    if B == 0:
        return float("nan") if reduce_mean else 0.0
    dummy = torch.zeros(B, 9, 2, device=pred_cats.device)
    a = _run(dummy, dummy, pred_cats, gt_cats)
    return a[2] / B if reduce_mean else a[2]


@torch.no_grad()
def compute_metrics_per_cls(pred_kp, gt_kp, pred_cats, gt_cats, compute_iou=False, **kwargs):
    """-> ([(cls, ADD, SADD, IOU, acc)], ADD, SADD, IOU, acc) as metrics.py:39-68 (IOU == 0.)."""
    if compute_iou:
        raise NotImplementedError("3D IoU (CPU EPnP lift + Qhull per sample, metrics.py:70-89) is outside the "
                                  "B200 hot path; call the reference implementation for it")
    a = _run(pred_kp, gt_kp, pred_cats, gt_cats)
    B = pred_kp.shape[0]
    rows = []
    for k in range(MAX_CLASSES):
        s_add, s_sadd, hits, n = a[4 + 4 * k: 8 + 4 * k]
        if n > 0:
            rows.append((k, s_add / n, s_sadd / n, 0., hits / n))
    return rows, a[0] / B, a[1] / B, 0., a[2] / B
