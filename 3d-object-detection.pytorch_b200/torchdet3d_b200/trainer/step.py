"""One fused training step: fwd -> loss -> bwd -> (grad all-reduce) -> optimizer -> metrics.

The hot loop body of the reference trainer (torchdet3d/trainer/train.py:44-55) as a fixed sequence
of C-ABI calls on static device buffers, with no host synchronisation; after a warm-up the whole
sequence is captured in a CUDA graph (>= 300 small kernels per step would otherwise be bound by
launch latency, SURVEY.md section 7 hard part 7).  Results stay on the device (`loss_terms`,
`kp`, `logits`, the metric accumulator) until the caller reads them.
"""
import ctypes as C

import torch

from .. import _lib as L
from ..builders.optim_builder import FusedOptimizer
from ..evaluation.metrics import MetricAccumulator
from ..models.regressor import Regressor, MAX_CLASSES, NUM_POINTS


class FusedTrainStep:
    def __init__(self, model, loss_manager, optimizer, batch, height, width, use_graph=True, allreduce=None,
                 metrics=True, shared=None):
        assert isinstance(model, Regressor) and isinstance(optimizer, FusedOptimizer)
        self.model, self.loss_manager, self.optimizer = model, loss_manager, optimizer
        self.use_graph, self.allreduce, self.with_metrics = use_graph, allreduce, metrics
        dev = model._flat.device
        if dev.type != "cuda":
            raise L.Td3dError("FusedTrainStep needs the model on a CUDA device (no CPU fallback)")
        self.device = dev
        self.imgs = torch.zeros(batch, 3, height, width, device=dev)
        self.gt_kp = torch.zeros(batch, NUM_POINTS // 2, 2, device=dev)
        self.cats = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.kp = torch.zeros(batch, NUM_POINTS // 2, 2, device=dev)
        self.logits = torch.zeros(batch, model.num_classes, device=dev)
        self.d_kp = torch.zeros(batch, NUM_POINTS, device=dev)
        self.d_logits = torch.zeros(batch, model.num_classes, device=dev)
        self.loss_terms = torch.zeros(8, device=dev)
        # epoch accumulators; `shared` = another FusedTrainStep whose accumulators this one adds to (the Trainer uses
        # one step object per batch shape: a partial last batch must land in the same epoch sums)
        if shared is not None:
            self.loss_sum, self.metrics, self._crops = shared.loss_sum, shared.metrics, shared._crops
        else:
            self.loss_sum = torch.zeros(8, dtype=torch.float64, device=dev)   # sum over steps of terms * B
            self.metrics = MetricAccumulator(dev)
            self._crops = [0]
        self.keep = None                 # optional injected cls_fc dropout mask [B, head_ch] (parity tests); None = in-kernel Philox
        self.batch = batch
        self.steps_done = 0
        self._graph = None
        self._graph_key = None
        self._warm = 0
        self.launches_per_step = None

    # -- the launch sequence ----------------------------------------------------------------------
    def _sequence(self):
        m, lib = self.model, L.lib()
        m.train()
        plan = m._plan_for(self.imgs)
        m._last_plan = plan
        L.check(lib.td3d_plan_set_dropout_counter(plan.handle, L.ptr(self.optimizer.steps)))
        m._dropout_counter_ref = self.optimizer.steps        # the plan keeps the raw pointer: keep the tensor alive with the model
        m.pack(plan)
        st = L.stream()
        L.check(lib.td3d_forward(plan.handle, L.ptr(self.imgs), L.ptr(self.cats), L.ptr(self.keep), C.c_uint64(m.dropout_seed), 1,
                                 L.ptr(self.kp), L.ptr(self.logits), st))
        desc = self.loss_manager.loss_desc()
        has_cls = bool(self.loss_manager.class_criterions)
        L.check(lib.td3d_loss_fwd_bwd(C.byref(desc), L.ptr(self.kp), L.ptr(self.gt_kp),
                                      L.ptr(self.logits) if has_cls else None, L.ptr(self.cats) if has_cls else None,
                                      self.batch, m.num_classes, L.ptr(self.loss_terms), L.ptr(self.d_kp),
                                      L.ptr(self.d_logits), st))
        if self.allreduce is None:
            m._backward_impl(plan, self.d_kp, self.d_logits)
        else:
            self.allreduce.backward_with_overlap(m, plan, self.d_kp, self.d_logits)
        d = self.optimizer.desc()
        L.check(lib.td3d_optim_step(plan.handle, C.byref(d), L.ptr(self.optimizer.state0), L.ptr(self.optimizer.state1),
                                    L.ptr(self.optimizer.steps), L.ptr(m.present), st))
        if self.with_metrics:
            self.metrics.update(self.kp, self.gt_kp, self.logits, self.cats)
        self.loss_sum.add_(self.loss_terms.double() * self.batch)

    def _key(self):
        g = self.optimizer.param_groups[0]
        lm = self.loss_manager
        return (float(g['lr']), float(g['weight_decay']), float(self.optimizer.grad_scale),
                float(lm.lam_cls) if lm.use_alwa else 1.0, self.model._flat.data_ptr())

    def load(self, imgs, gt_kp, cats, dropout_keep=None):
        """Stage one batch (pinned host or device tensors) into the static device buffers."""
        self.imgs.copy_(imgs, non_blocking=True)
        self.gt_kp.copy_(gt_kp, non_blocking=True)
        self.cats.copy_(cats, non_blocking=True)
        if dropout_keep is not None:
            if self.keep is None:
                if self._graph is not None:
                    raise L.Td3dError("dropout_keep must be injected from the first step on (the captured graph has no mask input)")
                self.keep = torch.ones(self.batch, self.model.head_ch, device=self.device)
            self.keep.copy_(dropout_keep, non_blocking=True)

    def run(self):
        """Execute one step on the staged batch. Asynchronous."""
        key = self._key()
        if self.use_graph and self._graph is not None and key == self._graph_key:
            self._graph.replay()
        elif self.use_graph and self._warm >= 2:
            # capture (also after an LR / loss-weight change: scalars are baked into kernel arguments)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread may touch CUDA events while this thread captures (N > 1)
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._sequence()
            self._graph, self._graph_key = g, key
            g.replay()
        else:
            self._sequence()
            self._warm += 1
        self.steps_done += 1
        self._crops[0] += self.batch
        self.model.mark_packed()
        # running statistics and BN affine parameters moved on EVERY branch (a graph replay never runs the Python
        # of _sequence): the eval-mode fold must be rebuilt before the next eval forward
        self.model._eval_fold_stale = True

    def __call__(self, imgs, gt_kp, cats, dropout_keep=None):
        self.load(imgs, gt_kp, cats, dropout_keep)
        self.run()
        return self.loss_terms

    def read_epoch(self):
        """One D2H sync: (mean loss terms[8], ADD, SADD, acc, per-class rows) since the last reset."""
        acc = self.metrics.read()
        n = max(acc[3], 1.0)
        loss = (self.loss_sum / max(self._crops[0], 1)).cpu().tolist()
        return dict(loss=loss, ADD=acc[0] / n, SADD=acc[1] / n, acc=acc[2] / n, count=acc[3], raw=acc)

    def reset_epoch(self):
        self.metrics.reset()
        self.loss_sum.zero_()
        self.steps_done = 0
        self._crops[0] = 0
