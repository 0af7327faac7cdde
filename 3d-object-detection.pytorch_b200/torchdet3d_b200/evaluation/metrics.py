"""ADD / symmetric-ADD / accuracy on the fused sm_100a metrics kernel.

Drop-in for torchdet3d/evaluation/metrics.py:10-68 (`compute_average_distance`, `compute_accuracy`,
`compute_metrics_per_cls`).  The reference runs a 9x9 Python loop of tiny kernels plus three host
syncs per call; here one launch (one warp per sample) accumulates everything -- totals and the
per-class breakdown -- into a device-side f64 accumulator.  The 3D-IoU column (metrics.py:70-89: EPnP lift
+ Objectron box fit / clipping + Qhull volume, a Python loop over the batch on the CPU in the reference) is the
CUDA kernel `td3d_iou_2d_based` (csrc/k_iou.cu), one thread per sample pair, so
`compute_iou=True` of unmodified callers (scripts/main.py:105) stays on the device; `set_iou_backend(fn)`
swaps in another implementation (e.g. the reference's CPU function for a cross-check).
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
def compute_2d_based_iou(pred_kp, gt_kp, reduce_mean=True):
    """Reference metrics.py:70-89 (lift both keypoint sets in portrait mode, fit boxes, intersect them) as ONE launch of
    `td3d_iou_2d_based` over the batch: no device -> host copy of the keypoints, no per-sample Python loop, no scipy.
    pred_kp / gt_kp: CUDA tensors [n, 9, 2].  A pair the reference drops (Qhull / LinAlg error) contributes 0 here too."""
    assert pred_kp.dim() == 3
    n = pred_kp.shape[0]
    if n == 0:
        return 0
    L.require_b200()
    p = pred_kp.detach().float().contiguous()
    g = gt_kp.detach().float().contiguous()
    out = torch.empty(n, dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        L.check(L.lib().td3d_iou_2d_based(L.ptr(p), L.ptr(g), n, 1, None, L.ptr(out), L.stream()))
    total = float(out.sum().item())
    return total / n if reduce_mean else total


_iou_backend = None          # callable(pred_kp[n,9,2], gt_kp[n,9,2], reduce_mean=False) -> summed IoU; None = the CUDA kernel


def set_iou_backend(fn):
    """Replace the 3D-IoU implementation used when `compute_iou=True` (signature of the reference's
    `compute_2d_based_iou`, metrics.py:70-89), e.g. with the reference's own CPU function for a cross-check.
    `None` restores the built-in CUDA kernel."""
    global _iou_backend
    _iou_backend = fn


def _iou_fn():
    return _iou_backend if _iou_backend is not None else compute_2d_based_iou


@torch.no_grad()
def compute_metrics_per_cls(pred_kp, gt_kp, pred_cats, gt_cats, compute_iou=True, **kwargs):
    """-> ([(cls, ADD, SADD, IOU, acc)], ADD, SADD, IOU, acc) as metrics.py:39-68.  ADD / SADD / acc come
    from one kernel launch; the IOU column from `compute_2d_based_iou` (one more launch per class present)."""
    a = _run(pred_kp, gt_kp, pred_cats, gt_cats)
    B = pred_kp.shape[0]
    iou = _iou_fn() if compute_iou else None
    rows, tot_iou = [], 0.
    for k in range(MAX_CLASSES):
        s_add, s_sadd, hits, n = a[4 + 4 * k: 8 + 4 * k]
        if n > 0:
            s_iou = 0.
            if iou is not None:
                sel = gt_cats == k
                s_iou = float(iou(pred_kp[sel], gt_kp[sel], reduce_mean=False))
            tot_iou += s_iou
            rows.append((k, s_add / n, s_sadd / n, s_iou / n, hits / n))
    return rows, a[0] / B, a[1] / B, tot_iou / B, a[2] / B
