"""`Trainer` -- same dataclass fields and `train(epoch, is_last_epoch)` contract as the reference
(torchdet3d/trainer/train.py:10-114).

With the B200 Regressor + FusedOptimizer + fused LossManager the loop body is a FusedTrainStep
(one CUDA-graph replay per iteration) and meters are fed from device-side accumulators that are
read back only every `print_freq` iterations; the reference performs >= B+6 device->host syncs
per iteration (train.py:57-76, model_builder.py:137).  Any other model/optimizer falls back to the
reference-style loop on the same hooks (compatibility, not the fast path).
"""
import datetime
import time
from dataclasses import dataclass

from ..builders.optim_builder import FusedOptimizer
from ..evaluation import compute_average_distance, compute_accuracy
from ..losses import LossManager
from ..models.regressor import Regressor
from ..utils import AverageMeter, save_snap, put_on_device
from .step import FusedTrainStep


@dataclass(init=True)
class Trainer:
    model: object
    train_loader: object
    optimizer: object
    scheduler: object
    loss_manager: object
    writer: object
    max_epoch: int
    log_path: str
    device: str = 'cuda'
    save_chkpt: bool = True
    debug: bool = False
    debug_steps: int = 30
    save_freq: int = 10
    print_freq: int = 10
    train_step: int = 0
    use_graph: bool = True
    allreduce: object = None

    def _fused(self):
        net = self.model.module if hasattr(self.model, "module") else self.model
        return (isinstance(net, Regressor) and isinstance(self.optimizer, FusedOptimizer)
                and isinstance(self.loss_manager, LossManager) and not self.loss_manager.use_alwa)

    def train(self, epoch, is_last_epoch):
        if self._fused():
            self._train_fused(epoch)
        else:
            self._train_generic(epoch)
        if self.save_chkpt and (epoch % self.save_freq == 0 or is_last_epoch) and not self.debug:
            save_snap(self.model, self.optimizer, self.scheduler, epoch, self.log_path)
        if self.scheduler is not None:
            self.scheduler.step()

    # ---- fast path ------------------------------------------------------------------------------
    def _train_fused(self, epoch):
        net = self.model.module if hasattr(self.model, "module") else self.model
        net.train()
        self.num_iters = len(self.train_loader)
        losses, ADD_meter, SADD_meter, ACC_meter, batch_time = (AverageMeter() for _ in range(5))
        step = None
        start = time.time()
        pending = 0
        for it, (imgs, gt_kp, gt_cats) in enumerate(self.train_loader):
            if step is None or step.imgs.shape != imgs.shape:
                first = step is None
                step = self._get_step(net, imgs.shape)      # all step objects of this trainer share the epoch accumulators
                if first:
                    step.reset_epoch()
            step(imgs, gt_kp, gt_cats)
            self.train_step += 1
            pending += 1
            last = it == self.num_iters - 1
            if (it % self.print_freq == 0) or last or (self.debug and it == self.debug_steps):
                r = step.read_epoch()                      # the only host sync
                n = int(r["count"])
                for meter, val in ((losses, r["loss"][0]), (ADD_meter, r["ADD"]), (SADD_meter, r["SADD"]),
                                   (ACC_meter, r["acc"])):
                    meter.val = meter.avg = val
                    meter.count, meter.sum = n, val * n
                if self.writer is not None:
                    gs = self.train_step - 1
                    self.writer.add_scalar('Train/loss', losses.avg, global_step=gs)
                    self.writer.add_scalar('Train/ADD', ADD_meter.avg, global_step=gs)
                    self.writer.add_scalar('Train/SADD', SADD_meter.avg, global_step=gs)
                    self.writer.add_scalar('Train/ACC', ACC_meter.avg, global_step=gs)
                batch_time.update((time.time() - start) / max(pending, 1), pending)
                start, pending = time.time(), 0
                eta = batch_time.avg * ((self.num_iters - (it + 1)) + (self.max_epoch - (epoch + 1)) * self.num_iters)
                print('epoch: [{0}/{1}][{2}/{3}]\t'
                      'time {bt.val:.3f} ({bt.avg:.3f})\teta {eta}\t'
                      'cls acc {acc.avg:.3f}\tADD {ADD.avg:.4f}\tSADD {SADD.avg:.4f}\t'
                      'loss {loss.avg:.5f}\tlr {lr:.6f}'.format(
                          epoch, self.max_epoch, it, self.num_iters, bt=batch_time,
                          eta=str(datetime.timedelta(seconds=int(eta))), acc=ACC_meter, ADD=ADD_meter,
                          SADD=SADD_meter, loss=losses, lr=self.optimizer.param_groups[0]['lr']))
            if self.debug and it == self.debug_steps:
                break
        self.meters = dict(loss=losses, ADD=ADD_meter, SADD=SADD_meter, ACC=ACC_meter)

    def _get_step(self, net, shape):
        cache = self.__dict__.setdefault("_steps", {})
        key = tuple(shape)
        if key not in cache:
            shared = next(iter(cache.values())) if cache else None
            cache[key] = FusedTrainStep(net, self.loss_manager, self.optimizer, shape[0], shape[2], shape[3],
                                        use_graph=self.use_graph, allreduce=self.allreduce, shared=shared)
        return cache[key]

    # ---- reference-style loop (train.py:42-108) on the same hooks ------------------------------
    def _train_generic(self, epoch):
        losses, ADD_meter, SADD_meter, ACC_meter, batch_time = (AverageMeter() for _ in range(5))
        self.model.train()
        self.num_iters = len(self.train_loader)
        start = time.time()
        for it, (imgs, gt_kp, gt_cats) in enumerate(self.train_loader):
            imgs, gt_kp, gt_cats = put_on_device([imgs, gt_kp, gt_cats], self.device)
            pred_kp, pred_cats = self.model(imgs, gt_cats)
            loss = self.loss_manager.parse_losses(pred_kp, gt_kp, pred_cats, gt_cats, it)
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
            ADD, SADD = compute_average_distance(pred_kp, gt_kp)
            acc = compute_accuracy(pred_cats, gt_cats)
            n = imgs.size(0)
            losses.update(loss.item(), n)
            ADD_meter.update(ADD, n)
            SADD_meter.update(SADD, n)
            ACC_meter.update(acc, n)
            if self.writer is not None:
                self.writer.add_scalar('Train/loss', loss.item(), global_step=self.train_step)
                self.writer.add_scalar('Train/ADD', ADD_meter.avg, global_step=self.train_step)
                self.writer.add_scalar('Train/SADD', SADD_meter.avg, global_step=self.train_step)
                self.writer.add_scalar('Train/ACC', ACC_meter.avg, global_step=self.train_step)
            self.train_step += 1
            batch_time.update(time.time() - start)
            if (it % self.print_freq == 0) or (it == self.num_iters - 1):
                print(f'epoch: [{epoch}/{self.max_epoch}][{it}/{self.num_iters}]\ttime {batch_time.val:.3f}\t'
                      f'cls acc {ACC_meter.val:.3f} ({ACC_meter.avg:.3f})\tADD {ADD_meter.val:.4f} ({ADD_meter.avg:.4f})\t'
                      f'SADD {SADD_meter.val:.4f} ({SADD_meter.avg:.4f})\tloss {losses.avg:.5f}\t'
                      f"lr {self.optimizer.param_groups[0]['lr']:.6f}")
            start = time.time()
            if self.debug and it == self.debug_steps:
                break
        self.meters = dict(loss=losses, ADD=ADD_meter, SADD=SADD_meter, ACC=ACC_meter)
