"""Deployment forward for a fixed batch of crops: `InferSession` -- the B200 replacement for running the exported
model behind `utils/ie_wrappers.py:138-142` of the reference (all nine regressor heads, then the keypoints of the
arg-max class), i.e. `ModelWrapper.forward_to_onnx` (builders/model_builder.py:112-124) + its consumer.

The batch (BASELINE config 4: 4096 crops from the 2D detector) lives in one static device buffer and is processed
in micro-batches that fit L2-friendly workspaces; eval-mode BatchNorm is folded into the weights
(`td3d_pack_weights`), activations / residuals run in the producers' epilogues, and the whole sequence of C-ABI calls
is captured once in a CUDA graph.  Outputs stay on the device until read.
"""
import ctypes as C

import torch

from . import _lib as L
from .models.regressor import Regressor, MAX_CLASSES, NUM_POINTS


class InferSession:
    def __init__(self, model, batch, height, width, chunk=256, use_graph=True):
        assert isinstance(model, Regressor)
        dev = model._flat.device
        if dev.type != "cuda":
            raise L.Td3dError("InferSession needs the model on a CUDA device (no CPU fallback)")
        self.model, self.batch, self.chunk, self.use_graph, self.device = model, batch, min(chunk, batch), use_graph, dev
        self.imgs = torch.zeros(batch, 3, height, width, device=dev)
        self.kp = torch.zeros(batch, NUM_POINTS // 2, 2, device=dev)
        self.labels = torch.zeros(batch, dtype=torch.int64, device=dev)
        self.logits = torch.zeros(batch, model.num_classes, device=dev)
        self._kp_all = torch.zeros(MAX_CLASSES, self.chunk, NUM_POINTS, device=dev)
        self._graph = None
        self._warm = 0
        self._version = None
        self._copy_stream = None         # host -> device staging stream of the pipelined path (run_from_host)
        self._staged, self._consumed = [], []
        self._buffer_busy_on_main = True  # the static input buffer was last touched by work on the caller's stream

    def _run_chunk(self, c0, c1):
        m = self.model
        x = self.imgs[c0:c1]
        plan = m._plan_for(x)
        L.check(L.lib().td3d_forward_export(plan.handle, L.ptr(x), L.ptr(self._kp_all), L.ptr(self.logits[c0:c1]), 1,
                                            L.ptr(self.kp[c0:c1]), L.ptr(self.labels[c0:c1]), L.stream()))
        m._last_plan = plan

    def _sequence(self):
        for c0 in range(0, self.batch, self.chunk):
            self._run_chunk(c0, min(self.batch, c0 + self.chunk))

    def load(self, imgs):
        """Stage the crops (pinned host or device tensor) into the static device buffer."""
        self.imgs.copy_(imgs, non_blocking=True)
        self._buffer_busy_on_main = True

    def load_rois(self, frames, boxes, **kw):
        """Fill the batch straight from uint8 frames + detector boxes (preprocess.crop_resize_normalize): the fp32 crops
        are produced on the device, only the frames cross PCIe."""
        from .preprocess import crop_resize_normalize
        crop_resize_normalize(frames, boxes, size=tuple(self.imgs.shape[2:]), out=self.imgs, **kw)
        self._buffer_busy_on_main = True

    def run(self):
        m = self.model
        m.eval()
        self._buffer_busy_on_main = True
        # weights / running statistics may have moved since the graph was captured: the graph reads the packed arena,
        # which `pack` refreshes in place, so a re-capture is only needed when the arena itself was re-allocated
        m.pack(m._plan_for(self.imgs[:self.chunk]), for_eval=True)
        key = (m._flat.data_ptr(), m._packed.data_ptr())
        if self.use_graph and self._graph is not None and key == self._version:
            self._graph.replay()
        elif self.use_graph and self._warm >= 1:
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._sequence()
            self._graph, self._version = g, key
            g.replay()
        else:
            self._sequence()
            self._warm += 1
        return self.kp, self.labels, self.logits

    @torch.no_grad()
    def run_from_host(self, imgs):
        """Crops in (pinned) host memory: the micro-batches are copied on a staging stream while the previous ones are computed,
        so a step costs max(PCIe copy, compute) instead of their sum (2.47 GB of fp32 crops at batch 4096 take as long to copy
        as to process).  Micro-batch i is computed as soon as its copy has landed and may be overwritten by the next call as
        soon as it has been consumed; the launches are eager (programmatic dependent launch keeps them back to back)."""
        assert not imgs.is_cuda and imgs.is_pinned() and tuple(imgs.shape) == tuple(self.imgs.shape)
        m = self.model
        m.eval()
        m.pack(m._plan_for(self.imgs[:self.chunk]), for_eval=True)
        main = torch.cuda.current_stream(self.device)
        starts = list(range(0, self.batch, self.chunk))
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._staged = [torch.cuda.Event() for _ in starts]
            self._consumed = [torch.cuda.Event() for _ in starts]
        cs = self._copy_stream
        if self._buffer_busy_on_main:       # run() / load() / load_rois() used the buffer on the caller's stream since the last
            for ev in self._consumed:       # pipelined call: the first copies must wait for all of that
                ev.record(main)
            self._buffer_busy_on_main = False
        with torch.cuda.stream(cs):
            for i, c0 in enumerate(starts):
                c1 = min(self.batch, c0 + self.chunk)
                cs.wait_event(self._consumed[i])
                self.imgs[c0:c1].copy_(imgs[c0:c1], non_blocking=True)
                self._staged[i].record(cs)
        for i, c0 in enumerate(starts):
            main.wait_event(self._staged[i])
            self._run_chunk(c0, min(self.batch, c0 + self.chunk))
            self._consumed[i].record(main)
        return self.kp, self.labels, self.logits

    @torch.no_grad()
    def __call__(self, imgs):
        if not imgs.is_cuda and imgs.is_pinned():      # only pinned memory copies asynchronously
            return self.run_from_host(imgs)
        self.load(imgs)
        return self.run()
