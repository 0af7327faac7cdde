"""Data-parallel training: one process per GPU, replicas kept identical by summing the flat fp32
gradient arena over NCCL (NVLink 5 / NVSwitch) -- the B200 replacement for the reference's
single-process `torch.nn.DataParallel` wrap (scripts/main.py:60-61).

Semantics kept from the reference: the global batch is split along dim 0, BatchNorm statistics
stay per replica (DataParallel does not sync BN), the loss is a mean over the local shard and
averaging gradients over equally sized shards reproduces the global-batch mean.  A regressor
head is skipped by the optimizer only if its class is absent on EVERY rank (`present` is
max-reduced), matching grad=None semantics of the single-process reference.

The path has exactly one exchange step (gradients), so that is the only collective.  Payload is
small (4.4 M params = 17.7 MB for MobileNetV3-large): it is latency-, not bandwidth-bound, hence
a handful of buckets issued from inside the backward pass on a side stream, each as soon as the
backward stages that produce it have been enqueued; 1/world is folded into the optimizer kernel.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L


def plan_buckets(n_stages, n_buckets):
    """Split backward stages [0, n_stages) into contiguous groups; earlier (tail-of-network) stages
    first.  Returns [(stage_begin, stage_end)]."""
    n_buckets = max(1, min(n_buckets, n_stages))
    edges = [round(i * n_stages / n_buckets) for i in range(n_buckets + 1)]
    return [(edges[i], edges[i + 1]) for i in range(n_buckets) if edges[i + 1] > edges[i]]


def reduce_ready_chunks(flat_grads, chunks, group=None):
    """All-reduce (sum) the given [begin, end) float ranges of the flat gradient arena in place."""
    for b, e in chunks:
        if e > b:
            dist.all_reduce(flat_grads[b:e], op=dist.ReduceOp.SUM, group=group)


class GradAllReduce:
    def __init__(self, model, optimizer, n_buckets=3, group=None):
        assert dist.is_initialized(), "torch.distributed must be initialised (one process per GPU)"
        self.model, self.optimizer, self.group, self.n_buckets = model, optimizer, group, n_buckets
        self.world = dist.get_world_size(group)
        optimizer.grad_scale = 1.0 / self.world
        self.comm_stream = torch.cuda.Stream(model._flat.device) if model._flat.is_cuda else None
        self.broadcast_state()

    @torch.no_grad()
    def broadcast_state(self, src=0):
        """Replicas start identical (params + BN buffers from rank `src`)."""
        m = self.model
        for t in (m._flat, m._bn, m._nbt):
            dist.broadcast(t, src=src, group=self.group)
        m._packed_version = None

    def ready_range(self, plan, stage):
        b, e = C.c_int64(), C.c_int64()
        L.check(L.lib().td3d_backward_ready_range(plan.handle, stage, C.byref(b), C.byref(e)))
        return b.value, e.value

    def backward_with_overlap(self, model, plan, d_kp, d_logits):
        cur = torch.cuda.current_stream()
        prev_begin = model._flat.numel()
        for s0, s1 in plan_buckets(plan.n_stages, self.n_buckets):
            model._backward_impl(plan, d_kp, d_logits, s0, s1)
            begin, _ = self.ready_range(plan, s1)
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                reduce_ready_chunks(model._gflat, [(begin, prev_begin)], self.group)
                if s0 == 0:
                    dist.all_reduce(model.present, op=dist.ReduceOp.MAX, group=self.group)
            prev_begin = begin
        cur.wait_stream(self.comm_stream)
