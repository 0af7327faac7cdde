"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the EfficientNet-B0 / B3 regressors of BASELINE configs 3 and 5.

The reference has no B0 / B3 (`__AVAI_MODELS__`, builders/model_builder.py:14-17); SURVEY.md 8c defines the oracle as
torchvision's `efficientnet_b0/b3(...).features` (stochastic depth off: it is the only RNG use) passed through the
reference's own `model_wrapper` (model_builder.py:73-151) with output_channels 1280 / 1536.  torchvision IS present on
the GPU box, `/root/reference` is not, so the wrapper part (pool -> per-sample head -> sigmoid, cls_fc with dropout,
forward_to_onnx) is restated here, each function citing the wrapper lines it follows; tests/test_oracle_golden.py pins
the restatement against the real `model_wrapper` over the same torchvision features (golden effnet_b0.npz, and live
when /root/reference exists).  Never imported by the product package.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import torch_port as tp

MODELS = {"efficientnet_b0": 1280, "efficientnet_b3": 1536}


def features(name):
    """torchvision backbone `.features` with stochastic depth disabled."""
    import torchvision
    net = getattr(torchvision.models, name)(weights=None, stochastic_depth_prob=0.0)
    return net.features


class Wrapped(torch.nn.Module):
    """Restatement of ModelWrapper over a non-MobileNetV3 backbone (no classifier: model_builder.py:117,130 apply it
    only `if model_class is MobileNetV3`)."""

    def __init__(self, name, num_classes=9):
        super().__init__()
        width = MODELS[name]
        self.features = features(name)
        self.regressors = torch.nn.ModuleList(                       # model_builder.py:78-81,89-93
            [torch.nn.Sequential(torch.nn.Linear(width, tp.NUM_POINTS)) for _ in range(tp.MAX_CLASSES)])
        self.cls_fc = torch.nn.Sequential(torch.nn.Dropout(0.5), torch.nn.Linear(width, num_classes))   # :82-85
        self.num_classes = num_classes

    def pooled(self, x):
        f = self.features(x)                                          # extract_features
        return F.adaptive_avg_pool2d(f, 1).view(x.size(0), -1)        # _glob_feature_vector 'avg' (:95-110)

    def forward(self, x, cats, dropout_mask=None):                    # model_builder.py:126-146
        p = self.pooled(x)
        kp = torch.cat([self.regressors[int(c)](s) for c, s in zip(cats, p)])
        kp = torch.sigmoid(kp).view(x.size(0), tp.NUM_POINTS // 2, 2)
        if dropout_mask is None:
            logits = self.cls_fc(p)
        else:                                                         # injected keep-mask, scaled like nn.Dropout(0.5)
            logits = self.cls_fc[1](p * dropout_mask * 2.0)
        return kp, logits

    def forward_export(self, x):                                      # forward_to_onnx, model_builder.py:112-124
        p = self.pooled(x)
        out = [reg(p).view(1, x.size(0), tp.NUM_POINTS // 2, 2) for reg in self.regressors]
        return torch.sigmoid(torch.cat(out)), self.cls_fc(p)


def synth_state(name, seed=0, num_classes=9, bn_jitter=0.1):
    """Deterministic (numpy RNG) state_dict with the key names / shapes of `Wrapped(name)`."""
    m = Wrapped(name, num_classes)
    rng = np.random.default_rng(seed)
    bn_weights = {k[:-len("running_mean")] + "weight" for k in m.state_dict() if k.endswith("running_mean")}
    state = {}
    for k, v in m.state_dict().items():
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            t = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_mean"):
            t = torch.tensor(rng.normal(0, bn_jitter, shape).astype(np.float32))
        elif k.endswith("running_var"):
            t = torch.tensor((1.0 + bn_jitter * rng.random(shape)).astype(np.float32))
        elif len(shape) == 4:
            fan = shape[1] * shape[2] * shape[3]
            t = torch.tensor(rng.normal(0, math.sqrt(2.0 / max(fan, 1)), shape).astype(np.float32))
        elif len(shape) == 2:
            t = torch.tensor(rng.normal(0, 1.0 / math.sqrt(shape[1]), shape).astype(np.float32))
        elif k in bn_weights:                                 # BatchNorm gamma
            t = torch.tensor((1.0 + bn_jitter * rng.normal(size=shape)).astype(np.float32))
        else:                                                 # biases / BatchNorm beta
            t = torch.tensor(rng.normal(0, bn_jitter, shape).astype(np.float32))
        state[k] = t
    return state


def make(name, state, num_classes=9):
    m = Wrapped(name, num_classes)
    m.load_state_dict(state)
    return m


def forward_export(state, name, x):
    m = make(name, state).eval()
    with torch.no_grad():
        return m.forward_export(x)


def train_step(state, name, opt_state, imgs, gt_kp, cats, dropout_mask, loss_cfg=None, optim_cfg=None,
               step_optimizer=True):
    """trainer/train.py:46-55 on the wrapped torchvision backbone; same return contract as torch_port.train_step.
    Mutates `state` (weights + BN buffers) and `opt_state`."""
    m = make(name, state).train()
    kp, logits = m(imgs, cats, dropout_mask)
    total, reg, cls = tp.parse_losses(kp, gt_kp, logits, cats, loss_cfg)
    total.backward()
    grads = {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in m.named_parameters()}
    present = set(int(c) for c in torch.unique(cats))
    for k in range(tp.MAX_CLASSES):
        if k not in present:
            grads[f"regressors.{k}.0.weight"] = None
            grads[f"regressors.{k}.0.bias"] = None
    new = m.state_dict()
    for k in state:                                   # BatchNorm running statistics / counters moved
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            state[k] = new[k].detach().clone()
    if step_optimizer:
        tp.optim_step(state, grads, opt_state, optim_cfg)
    add, sadd = tp.average_distance(kp.detach(), gt_kp)
    acc = tp.accuracy(logits.detach(), cats)
    return dict(kp=kp.detach(), logits=logits.detach(), loss=float(total.detach()), grads=grads, add=add, sadd=sadd,
                acc=acc)


def layer_table(name, res):
    """[(kind, I_elems, O_elems, W_elems, MACs)] per conv / linear layer and crop (forward hooks on the torchvision
    modules) -- the roofline byte model of SURVEY.md 8d for configs 3 / 5."""
    m = Wrapped(name).eval()
    rows = []

    def hook(mod, inp, out):
        x = inp[0]
        macs = out.numel() * (mod.in_channels // mod.groups) * mod.kernel_size[0] * mod.kernel_size[1]
        kind = "dw" if mod.groups > 1 else ("se" if x.shape[-1] == 1 and x.shape[-2] == 1 else "pw")
        rows.append((kind, x.numel(), out.numel(), mod.weight.numel(), macs))

    hs = [mod.register_forward_hook(hook) for mod in m.features.modules() if isinstance(mod, torch.nn.Conv2d)]
    with torch.no_grad():
        m.pooled(torch.zeros(1, 3, res, res))
    for h in hs:
        h.remove()
    width = MODELS[name]
    rows.append(("fc", width, 18 + 9, width * (18 * 9 + 9), width * (18 + 9)))
    return rows
