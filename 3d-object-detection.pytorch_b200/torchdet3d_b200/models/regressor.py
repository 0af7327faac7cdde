"""B200-native second-stage 3D box regressor: the drop-in for the object the reference's
`build_model` returns (ModelWrapper over MobileNetV3, torchdet3d/builders/model_builder.py:73-151,
torchdet3d/models/mobilenetv3.py:169-221).

Same nn.Module protocol as the reference object -- `model(img, cats) -> (kp[B,9,2], logits[B,nc])`,
`forward_to_onnx(img) -> (kp_all[9,B,9,2], logits)`, `.train()/.eval()`, `.to(device)`,
`.parameters()`, `state_dict()` with identical keys and shapes -- but there are no torch.nn layers
inside: parameters are views into one flat fp32 arena, and forward/backward are single calls into
the sm_100a kernels of libtd3d.so through the C ABI (include/td3d.h).  No CPU fallback.
"""
import ctypes as C
import math

import torch
import torch.nn as nn

from .. import _lib as L
from .arch import blocks_for, arch_of, AVAILABLE_MODELS

MAX_CLASSES = 9      # model_builder.py:78
NUM_POINTS = 18      # model_builder.py:73


class _Node(nn.Module):
    """Name-space node so that parameter paths equal the reference state_dict keys."""


class _Plan:
    """One td3d_plan (fixed batch/resolution) + its workspace."""

    def __init__(self, owner, batch, height, width):
        lib = L.lib()
        self.batch, self.height, self.width = batch, height, width
        self.handle = C.c_void_p()
        L.check(lib.td3d_plan_create(C.byref(owner._net_desc), batch, height, width, owner._dtype_code,
                                     owner._gemm_impl, C.byref(self.handle)))
        self.sizes = L.Sizes()
        L.check(lib.td3d_plan_sizes(self.handle, C.byref(self.sizes)))
        self.workspace = None
        self.n_stages = lib.td3d_backward_stages(self.handle)

    def bind(self, owner):
        if self.workspace is None or self.workspace.device != owner._flat.device:
            self.workspace = torch.empty(int(self.sizes.workspace_bytes), dtype=torch.uint8, device=owner._flat.device)
        L.check(L.lib().td3d_plan_bind(self.handle, L.ptr(owner._flat), L.ptr(owner._gflat), L.ptr(owner._bn),
                                       L.ptr(owner._nbt), L.ptr(owner._packed), L.ptr(self.workspace)))

    def __del__(self):
        try:
            if self.handle:
                L.lib().td3d_plan_destroy(self.handle)
        except Exception:
            pass


class _RegressorFn(torch.autograd.Function):
    """Autograd bridge: `loss.backward()` on the outputs runs td3d_backward."""

    @staticmethod
    def forward(ctx, model, img, cats, keep, anchor):
        kp, logits = model._forward_impl(img, cats, keep, training=True)
        ctx.model = model
        ctx.plan = model._last_plan
        return kp, logits

    @staticmethod
    def backward(ctx, d_kp, d_logits):
        model = ctx.model
        if d_kp is None:
            d_kp = torch.zeros(ctx.plan.batch, NUM_POINTS // 2, 2, device=model._flat.device)
        if d_logits is None:
            d_logits = torch.zeros(ctx.plan.batch, model.num_classes, device=model._flat.device)
        model._backward_impl(ctx.plan, d_kp.contiguous().float(), d_logits.contiguous().float())
        model._publish_grads()
        return None, None, None, None, torch.zeros_like(model._anchor)


class Regressor(nn.Module):
    def __init__(self, name, num_classes=9, export_mode=False, compute_dtype="fp32", gemm_impl=L.GEMM_AUTO):
        super().__init__()
        assert name in AVAILABLE_MODELS, f"Wrong model name parameter. Expected one of {AVAILABLE_MODELS}"
        self.model_name, self.num_classes, self.export_mode = name, num_classes, export_mode
        self._dtype_code = L.dtype_code(compute_dtype)
        self._gemm_impl = gemm_impl
        stem, blocks, last_ch, head_ch = blocks_for(name)
        self._blocks_py = blocks
        self._block_array = (L.BlockDesc * len(blocks))(*[L.BlockDesc(**b) for b in blocks])
        self._net_desc = L.NetDesc(arch_of(name), stem, len(blocks), self._block_array, last_ch, head_ch, num_classes,
                                   MAX_CLASSES, NUM_POINTS)
        self.head_ch = head_ch
        self._plans = {}
        self._last_plan = None
        self._packed_version = None
        self._eval_fold_stale = True
        self._dropout_calls = 0
        self.dropout_seed = 0x5DEECE66D
        self.infer_chunk = 512            # forward_to_onnx micro-batch (crops)
        self.present = None              # device i32[9]: heads that received a gradient (last backward)

        probe = _Plan(self, 1, 32, 32)   # shape-independent: parameter / BN tables
        self._sizes = probe.sizes
        lib = L.lib()
        self._param_table, self._bn_table = [], []
        for i in range(self._sizes.n_param_tensors):
            info = L.ParamInfo()
            L.check(lib.td3d_plan_param_info(probe.handle, i, C.byref(info)))
            self._param_table.append((info.name.decode(), int(info.offset), int(info.numel),
                                      tuple(int(info.shape[d]) for d in range(info.ndim))))
        for i in range(self._sizes.n_bn):
            info = L.BnInfo()
            L.check(lib.td3d_plan_bn_info(probe.handle, i, C.byref(info)))
            self._bn_table.append((info.name.decode(), int(info.offset), int(info.channels)))
        del probe

        self._flat = torch.zeros(int(self._sizes.param_floats))
        self._gflat = torch.zeros(int(self._sizes.param_floats))
        self._bn = torch.zeros(int(self._sizes.bn_floats))
        self._nbt = torch.zeros(int(self._sizes.n_bn), dtype=torch.int64)
        self._packed = torch.zeros(int(self._sizes.packed_bytes), dtype=torch.uint8)
        self._anchor = torch.zeros((), requires_grad=True)
        self._params = []
        for pname, off, numel, shape in self._param_table:
            p = nn.Parameter(self._flat[off:off + numel].view(shape))
            self._register(pname, p, is_param=True)
            self._params.append(p)
        for bname, off, ch in self._bn_table:
            self._register(bname + ".running_mean", self._bn[off:off + ch], is_param=False)
            self._register(bname + ".running_var", self._bn[off + ch:off + 2 * ch], is_param=False)
        for i, (bname, _, _) in enumerate(self._bn_table):
            self._register(bname + ".num_batches_tracked", self._nbt[i], is_param=False)
        self._initialize_weights()

    # ---- module tree with reference key names -------------------------------------------------
    def _register(self, dotted, tensor, is_param):
        node = self
        parts = dotted.split(".")
        for p in parts[:-1]:
            if p not in node._modules:
                node.add_module(p, _Node())
            node = node._modules[p]
        if is_param:
            node.register_parameter(parts[-1], tensor)
        else:
            node.register_buffer(parts[-1], tensor)

    @torch.no_grad()
    def _initialize_weights(self):
        """Same distributions as the reference: conv N(0, sqrt(2/(k*k*Cout))), BN 1/0, Linear
        N(0, .01) with zero bias (mobilenetv3.py:205-218); regressors / cls_fc are created after
        that pass and keep torch's default Linear init (model_builder.py:76-87)."""
        for (pname, off, numel, shape), p in zip(self._param_table, self._params):
            if pname.startswith("regressors") or pname.startswith("cls_fc"):
                bound = 1.0 / math.sqrt(self.head_ch)
                p.uniform_(-bound, bound)
            elif len(shape) == 4:
                p.normal_(0, math.sqrt(2.0 / (shape[2] * shape[3] * shape[0])))
            elif len(shape) == 2:
                p.normal_(0, 0.01)
            elif pname.endswith(".weight"):
                p.fill_(1.0)            # BN gamma (BatchNorm1d of the classifier defaults to 1/0 too)
            else:
                p.zero_()
        for bname, off, ch in self._bn_table:
            self._bn[off:off + ch].zero_()
            self._bn[off + ch:off + 2 * ch].fill_(1.0)
        self._nbt.zero_()

    # ---- device movement keeps the flat arenas -------------------------------------------------
    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        named = dict(self.named_parameters())
        self._params = [named[pname] for pname, _, _, _ in self._param_table]   # _apply may re-create Parameters
        flat = fn(self._flat)
        if flat.dtype != torch.float32:
            raise TypeError("td3d keeps fp32 master parameters; use compute_dtype='bf16' for bf16 compute")
        device = flat.device
        # re-flatten: parameters may have been replaced by independent copies
        new_flat = torch.empty_like(self._flat, device=device)
        new_bn = torch.empty_like(self._bn, device=device)
        new_nbt = torch.empty_like(self._nbt, device=device)
        new_flat.zero_()
        with torch.no_grad():
            for (pname, off, numel, shape), p in zip(self._param_table, self._params):
                new_flat[off:off + numel].copy_(p.data.reshape(-1))
                p.data = new_flat[off:off + numel].view(shape)
                p.grad = None
            for i, (bname, off, ch) in enumerate(self._bn_table):
                node = self.get_submodule(bname)
                new_bn[off:off + ch].copy_(node.running_mean)
                new_bn[off + ch:off + 2 * ch].copy_(node.running_var)
                new_nbt[i].copy_(node.num_batches_tracked)
                node._buffers["running_mean"] = new_bn[off:off + ch]
                node._buffers["running_var"] = new_bn[off + ch:off + 2 * ch]
                node._buffers["num_batches_tracked"] = new_nbt[i]
        self._flat, self._bn, self._nbt = new_flat, new_bn, new_nbt
        self._gflat = torch.zeros_like(new_flat)
        self._packed = torch.zeros(int(self._sizes.packed_bytes), dtype=torch.uint8, device=device)
        self._anchor = torch.zeros((), device=device, requires_grad=True)
        self.present = torch.ones(MAX_CLASSES, dtype=torch.int32, device=device)
        self._packed_version = None
        self._eval_fold_stale = True
        for plan in self._plans.values():
            plan.workspace = None
        return self

    # ---- plans / packing ---------------------------------------------------------------------
    def _plan_for(self, img):
        if not img.is_cuda:
            raise L.Td3dError("td3d regressor runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self._flat.device != img.device:
            raise L.Td3dError(f"model is on {self._flat.device} but input on {img.device}; call model.to(device)")
        key = (img.shape[0], img.shape[2], img.shape[3])
        plan = self._plans.get(key)
        if plan is None:
            L.require_b200()
            plan = _Plan(self, *key)
            self._plans[key] = plan
        plan.bind(self)
        return plan

    def _param_version(self):
        return (self._flat.data_ptr(), sum(p._version for p in self._params), int(self._bn._version))

    def pack(self, plan=None, force=False, for_eval=False):
        """Refresh compute-layout weight copies (after parameters changed) and the folded
        eval-mode BN constants (after running statistics moved)."""
        if force or self._param_version() != self._packed_version or (for_eval and self._eval_fold_stale):
            plan = plan or self._last_plan or next(iter(self._plans.values()))
            L.check(L.lib().td3d_pack_weights(plan.handle, L.stream()))
            self._packed_version = self._param_version()
            self._eval_fold_stale = False

    def mark_packed(self):
        """The fused optimizer re-packs the weights inside td3d_optim_step -- but not the eval-mode
        BatchNorm fold (gamma/beta moved too), which the next eval forward has to rebuild."""
        self._packed_version = self._param_version()
        self._eval_fold_stale = True

    # ---- forward / backward ------------------------------------------------------------------
    def _forward_impl(self, img, cats, keep, training):
        img = img.contiguous().float()
        cats = cats.contiguous().to(torch.int64)
        plan = self._plan_for(img)
        self.pack(plan, for_eval=not training)
        B = img.shape[0]
        kp = torch.empty(B, NUM_POINTS // 2, 2, device=img.device)
        logits = torch.empty(B, self.num_classes, device=img.device)
        seed = (self.dropout_seed + 0x9E3779B97F4A7C15 * self._dropout_calls) & 0xFFFFFFFFFFFFFFFF
        if training:
            self._dropout_calls += 1
        if keep is not None:
            keep = keep.contiguous().float()
            assert keep.shape == (B, self.head_ch)
        L.check(L.lib().td3d_forward(plan.handle, L.ptr(img), L.ptr(cats), L.ptr(keep), C.c_uint64(seed),
                                     1 if training else 0, L.ptr(kp), L.ptr(logits), L.stream()))
        self._last_plan = plan
        self._keepalive = (img, cats, keep)      # the C side keeps raw pointers until backward
        if training:
            self._eval_fold_stale = True         # running statistics moved
        return kp, logits

    def _backward_impl(self, plan, d_kp, d_logits, stage_begin=0, stage_end=-1):
        if self.present is None:
            self.present = torch.ones(MAX_CLASSES, dtype=torch.int32, device=self._flat.device)
        L.check(L.lib().td3d_backward(plan.handle, L.ptr(d_kp), L.ptr(d_logits), L.ptr(self.present),
                                      stage_begin, stage_end, L.stream()))

    def _publish_grads(self):
        """Expose the gradient arena as `.grad` views. Heads whose class was absent from the batch
        get `.grad = None`, exactly like the reference (one host sync; the fused optimizer path
        avoids it by consuming the device-side mask)."""
        present = self.present.tolist()
        for (pname, off, numel, shape), p in zip(self._param_table, self._params):
            if pname.startswith("regressors."):
                k = int(pname.split(".")[1])
                if not present[k]:
                    p.grad = None
                    continue
            if self.num_classes == 1 and pname.startswith("cls_fc."):
                p.grad = None              # the reference never runs cls_fc then (model_builder.py:140-144)
                continue
            p.grad = self._gflat[off:off + numel].view(shape)

    def forward(self, x, cats, dropout_keep=None):
        """ModelWrapper.forward (model_builder.py:126-146). `dropout_keep` ([B, head_ch] of {0,1})
        optionally injects the cls_fc dropout mask (parity tests); default is an in-kernel Philox
        mask."""
        if self.export_mode:
            return self.forward_to_onnx(x)
        if self.training and torch.is_grad_enabled():
            kp, logits = _RegressorFn.apply(self, x, cats, dropout_keep, self._anchor)
        else:
            kp, logits = self._forward_impl(x, cats, dropout_keep, training=self.training)
        if self.num_classes > 1:
            return kp, logits
        return kp, cats.unsqueeze(dim=1)            # model_builder.py:143-144

    @torch.no_grad()
    def forward_to_onnx(self, x, select=False):
        """All nine heads (model_builder.py:112-124). With select=True also applies the deployment
        consumer (utils/ie_wrappers.py:138-142) on device: returns (kp_sel[B,9,2], labels[B], logits)."""
        x = x.contiguous().float()
        B = x.shape[0]
        if B > self.infer_chunk:
            return self._export_chunked(x, select)
        plan = self._plan_for(x)
        self.pack(plan, for_eval=True)
        kp_all = torch.empty(MAX_CLASSES, B, NUM_POINTS // 2, 2, device=x.device)
        logits = torch.empty(B, self.num_classes, device=x.device)
        kp_sel = torch.empty(B, NUM_POINTS // 2, 2, device=x.device) if select else None
        labels = torch.empty(B, dtype=torch.int64, device=x.device) if select else None
        L.check(L.lib().td3d_forward_export(plan.handle, L.ptr(x), L.ptr(kp_all), L.ptr(logits), 1 if select else 0,
                                            L.ptr(kp_sel), L.ptr(labels), L.stream()))
        self._last_plan = plan
        if select:
            return kp_sel, labels, logits
        if self.num_classes > 1:
            return kp_all, logits
        return kp_all, torch.zeros(B, device=x.device)

    def _export_chunked(self, x, select):
        """Large batches (BASELINE config 4: 4096 crops) run as micro-batches of `infer_chunk` crops through one plan:
        activations of a micro-batch stay L2/HBM friendly and tensors stay below the kernels' 2^31-element index range."""
        B, n = x.shape[0], self.infer_chunk
        outs = [self.forward_to_onnx(x[c0:min(B, c0 + n)], select=select) for c0 in range(0, B, n)]
        if select:
            return tuple(torch.cat([o[i] for o in outs], dim=0) for i in range(3))
        return torch.cat([o[0] for o in outs], dim=1), torch.cat([o[1] for o in outs], dim=0)

    def train(self, mode=True):
        super().train(mode)
        return self

    def extra_repr(self):
        return (f"{self.model_name}, num_classes={self.num_classes}, compute={'bf16' if self._dtype_code else 'fp32'}, "
                f"params={sum(n for _, _, n, _ in self._param_table)}")
