"""`Evaluator.val` -- validation hook of the reference (torchdet3d/evaluation/evaluate.py:73-149):
eval-mode forward with the GROUND-TRUTH class selecting the regressor head (:92), global and
per-class ADD / SADD / accuracy meters weighted by the full batch size (:96-106), TensorBoard
scalars and a per-class table.

Per-batch metrics come from one fused kernel launch; `visual_test` (drawing) and the 3D-IoU
column (CPU EPnP + Qhull per sample) are outside the B200 hot path.
"""
from dataclasses import dataclass

import torch

from ..utils import AverageMeter, OBJECTRON_CLASSES, put_on_device
from .metrics import compute_metrics_per_cls


@dataclass
class Evaluator:
    model: object
    val_loader: object
    test_loader: object
    cfg: dict
    writer: object
    max_epoch: int
    device: str = 'cuda'
    num_classes: int = len(OBJECTRON_CLASSES)
    samples: object = 'random'
    num_samples: int = 10
    path_to_save_imgs: str = './testing_images'
    debug: bool = False
    debug_steps: int = 30

    @torch.no_grad()
    def val(self, epoch=None, compute_iou=False):
        if compute_iou:
            raise NotImplementedError("3D IoU is outside the B200 hot path (evaluation/metrics.py:70-89 of the reference)")
        ADD_meter, SADD_meter, ACC_meter = AverageMeter(), AverageMeter(), AverageMeter()
        ADD_cls = [AverageMeter() for _ in range(self.num_classes)]
        SADD_cls = [AverageMeter() for _ in range(self.num_classes)]
        ACC_cls = [AverageMeter() for _ in range(self.num_classes)]
        self.model.eval()
        for it, (imgs, gt_kp, gt_cats) in enumerate(self.val_loader):
            imgs, gt_kp, gt_cats = put_on_device([imgs, gt_kp, gt_cats], self.device)
            pred_kp, pred_cats = self.model(imgs, gt_cats)
            per_cls, ADD, SADD, _, ACC = compute_metrics_per_cls(pred_kp, gt_kp, pred_cats, gt_cats, False)
            n = imgs.size(0)
            for cl, a, s, _, c in per_cls:        # weighted by the whole batch size, as the reference does
                ADD_cls[cl].update(a, n)
                SADD_cls[cl].update(s, n)
                ACC_cls[cl].update(c, n)
            ADD_meter.update(ADD, n)
            SADD_meter.update(SADD, n)
            ACC_meter.update(ACC, n)
            if self.debug and it == self.debug_steps:
                break
        if epoch is not None and self.writer is not None:
            self.writer.add_scalar('Val/ADD', ADD_meter.avg, global_step=epoch)
            self.writer.add_scalar('Val/SADD', SADD_meter.avg, global_step=epoch)
            self.writer.add_scalar('Val/ACC', ACC_meter.avg, global_step=epoch)
        rows = [("Average metrics", ADD_meter.avg, SADD_meter.avg, ACC_meter.avg)]
        rows += [(OBJECTRON_CLASSES[k] if k < len(OBJECTRON_CLASSES) else str(k), ADD_cls[k].avg, SADD_cls[k].avg,
                  ACC_cls[k].avg) for k in range(self.num_classes)]
        head = f"{'category name':<18}{'ADD':>10}{'SADD':>10}{'accuracy':>10}"
        body = "\n".join(f"{r[0]:<18}{r[1]:>10.4f}{r[2]:>10.4f}{r[3]:>10.4f}" for r in rows)
        print("\nComputed val metrics:\n" + (f"epoch: {epoch}\n" if epoch is not None else "") + head + "\n" + body)
        self.results = dict(ADD=ADD_meter.avg, SADD=SADD_meter.avg, ACC=ACC_meter.avg, per_class=rows[1:])
        return self.results

    def run_eval_pipe(self, visual_only=False):
        if not visual_only:
            self.val(compute_iou=False)
