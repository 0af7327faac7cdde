"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code; nothing under the product package may
import this file.  Allowed importers: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline /
--impl reference legs).

CPU restatement (plain PyTorch ops, fp32 or fp64, functional style over a flat state dict) of the
reference's second-stage 3D box regressor hot path:

  backbone      /root/reference/torchdet3d/models/mobilenetv3.py:20-221
  heads/forward /root/reference/torchdet3d/builders/model_builder.py:73-151
  losses        /root/reference/torchdet3d/losses/regression_losses.py:8-115,
                /root/reference/torchdet3d/builders/loss_builder.py:7-28
  metrics       /root/reference/torchdet3d/evaluation/metrics.py:10-68
  optimizers    /root/reference/torchdet3d/builders/optim_builder.py:5-19 (torch.optim.SGD / AdamW)
  train step    /root/reference/torchdet3d/trainer/train.py:46-55

Parity status: PINNED.  tests/test_oracle_golden.py checks this file against golden vectors in
tests/golden/ that were produced by executing the unmodified reference modules in the build
container (oracle/make_golden.py, via oracle/refshim.py), and -- when /root/reference is present --
directly against the live reference.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# architecture tables (mobilenetv3.py:20-52). columns: k, t, c, SE, HS, s
# --------------------------------------------------------------------------------------------
_CFGS = {
    "mobilenetv3_large": dict(mode="large", head=1280, rows=[
        (3, 1, 16, 0, 0, 1), (3, 4, 24, 0, 0, 2), (3, 3, 24, 0, 0, 1), (5, 3, 40, 1, 0, 2),
        (5, 3, 40, 1, 0, 1), (5, 3, 40, 1, 0, 1), (3, 6, 80, 0, 1, 2), (3, 2.5, 80, 0, 1, 1),
        (3, 2.3, 80, 0, 1, 1), (3, 2.3, 80, 0, 1, 1), (3, 6, 112, 1, 1, 1), (3, 6, 112, 1, 1, 1),
        (5, 6, 160, 1, 1, 2), (5, 6, 160, 1, 1, 1), (5, 6, 160, 1, 1, 1)]),
    "mobilenetv3_small": dict(mode="small", head=1024, rows=[
        (3, 1, 16, 1, 0, 2), (3, 4.5, 24, 0, 0, 2), (3, 3.67, 24, 0, 0, 1), (5, 4, 40, 1, 1, 2),
        (5, 6, 40, 1, 1, 1), (5, 6, 40, 1, 1, 1), (5, 3, 48, 1, 1, 1), (5, 3, 48, 1, 1, 1),
        (5, 6, 96, 1, 1, 2), (5, 6, 96, 1, 1, 1), (5, 6, 96, 1, 1, 1)]),
}
MAX_CLASSES = 9          # model_builder.py:78
NUM_POINTS = 18          # model_builder.py:73
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def make_divisible(v, divisor=8, min_value=None):
    """mobilenetv3.py:54-71."""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def block_table(name):
    """Expanded per-block description; mirrors MobileNetV3.__init__ (mobilenetv3.py:176-186)."""
    spec = _CFGS[name]
    cin = make_divisible(16)
    blocks = []
    for k, t, c, se, hs, s in spec["rows"]:
        cout = make_divisible(c)
        exp = make_divisible(cin * t)
        blocks.append(dict(k=k, cin=cin, exp=exp, cout=cout, se=bool(se), hs=bool(hs), stride=s,
                           expand=(cin != exp), residual=(s == 1 and cin == cout),
                           se_hidden=make_divisible(exp // 4) if se else 0))
        cin = cout
    return dict(stem=16, blocks=blocks, last_in=cin, last_exp=blocks[-1]["exp"], head=spec["head"])


# --------------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------------
def param_shapes(name, num_classes=9):
    """Ordered {key: shape} identical to reference `model.state_dict()` keys (incl. BN buffers)."""
    tab = block_table(name)
    out = OrderedDict()

    def bn(prefix, c):
        out[prefix + ".weight"] = (c,)
        out[prefix + ".bias"] = (c,)
        out[prefix + ".running_mean"] = (c,)
        out[prefix + ".running_var"] = (c,)
        out[prefix + ".num_batches_tracked"] = ()

    out["features.0.0.weight"] = (tab["stem"], 3, 3, 3)
    bn("features.0.1", tab["stem"])
    for i, b in enumerate(tab["blocks"], start=1):
        p = f"features.{i}.conv."
        j = 0
        if b["expand"]:
            out[p + "0.weight"] = (b["exp"], b["cin"], 1, 1)
            bn(p + "1", b["exp"])
            j = 3
        out[p + f"{j}.weight"] = (b["exp"], 1, b["k"], b["k"])
        bn(p + f"{j + 1}", b["exp"])
        # dw-first layout: dw, bn, act, se, pw, bn ; general: pw,bn,act,dw,bn,se,act,pw,bn
        se_idx = j + 3 if not b["expand"] else j + 2
        pw_idx = 4 if not b["expand"] else 7
        if b["se"]:
            out[p + f"{se_idx}.fc.0.weight"] = (b["se_hidden"], b["exp"])
            out[p + f"{se_idx}.fc.0.bias"] = (b["se_hidden"],)
            out[p + f"{se_idx}.fc.2.weight"] = (b["exp"], b["se_hidden"])
            out[p + f"{se_idx}.fc.2.bias"] = (b["exp"],)
        out[p + f"{pw_idx}.weight"] = (b["cout"], b["exp"], 1, 1)
        bn(p + f"{pw_idx + 1}", b["cout"])
    out["conv.0.weight"] = (tab["last_exp"], tab["last_in"], 1, 1)
    bn("conv.1", tab["last_exp"])
    out["classifier.0.weight"] = (tab["head"], tab["last_exp"])
    out["classifier.0.bias"] = (tab["head"],)
    bn("classifier.1", tab["head"])
    for k in range(MAX_CLASSES):
        out[f"regressors.{k}.0.weight"] = (NUM_POINTS, tab["head"])
        out[f"regressors.{k}.0.bias"] = (NUM_POINTS,)
    out["cls_fc.1.weight"] = (num_classes, tab["head"])
    out["cls_fc.1.bias"] = (num_classes,)
    return out


def synth_state(name, seed=0, num_classes=9, dtype=torch.float32, bn_jitter=0.1):
    """Deterministic synthetic weights, independent of torch's RNG stream and module construction
    order, so the reference, this oracle and the CUDA path can all be loaded with identical
    values.  Distributions follow the reference's init (mobilenetv3.py:205-218; heads keep
    torch-default Linear init, model_builder.py:76-87) with a small jitter on BN affine/running
    stats so that every BN term is exercised."""
    rng = np.random.default_rng(seed)
    state = OrderedDict()
    for key, shape in param_shapes(name, num_classes).items():
        if key.endswith("num_batches_tracked"):
            state[key] = torch.zeros((), dtype=torch.int64)
            continue
        if key.endswith("running_mean"):
            v = rng.normal(0.0, bn_jitter, shape)
        elif key.endswith("running_var"):
            v = 1.0 + rng.uniform(-bn_jitter, bn_jitter, shape)
        elif len(shape) == 4:                                   # conv
            n = shape[2] * shape[3] * shape[0]
            v = rng.normal(0.0, math.sqrt(2.0 / n), shape)
        elif len(shape) == 2:
            if key.startswith("regressors") or key.startswith("cls_fc"):
                bound = 1.0 / math.sqrt(shape[1])
                v = rng.uniform(-bound, bound, shape)
            elif ".fc." in key:                                 # SE linears: N(0, .01) is too
                v = rng.normal(0.0, 0.1, shape)                 # flat to test; widen a little
            else:
                v = rng.normal(0.0, 0.01, shape)
        else:                                                   # 1-d: BN gamma/beta or biases
            is_bn_weight = key.endswith(".weight")
            if is_bn_weight:
                v = 1.0 + rng.normal(0.0, bn_jitter, shape)
            else:
                v = rng.normal(0.0, bn_jitter, shape)
        state[key] = torch.tensor(np.asarray(v), dtype=dtype)
    return state


def synth_batch(batch, res=224, seed=1234, num_classes=9, all_classes=False, dtype=torch.float32):
    """Synthetic crops in the shape of tests/test_pipeline.py:12-15,52 (numpy RNG, travels)."""
    rng = np.random.default_rng(seed)
    imgs = torch.tensor(rng.random((batch, 3, res, res), dtype=np.float32)).to(dtype)
    gt_kp = torch.tensor(rng.random((batch, 9, 2), dtype=np.float32)).to(dtype)
    cats = rng.integers(0, num_classes, size=(batch,))
    if all_classes and batch >= num_classes:
        cats[:num_classes] = np.arange(num_classes)
    # dropout keep-mask for cls_fc, wide enough for any head width; callers slice [:, :C_f]
    mask = torch.tensor((rng.random((batch, 2048)) >= 0.5).astype(np.float32))
    return imgs, gt_kp, torch.tensor(cats, dtype=torch.int64), mask


# --------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------
def h_sigmoid(x):
    """mobilenetv3.py:74-80."""
    return F.relu6(x + 3.0) / 6.0


def h_swish(x):
    """mobilenetv3.py:83-89."""
    return x * h_sigmoid(x)


def _bn(state, prefix, x, training, new_stats):
    """nn.BatchNorm{1,2}d semantics (momentum .1, eps 1e-5; biased var to normalise, unbiased
    var into running_var; num_batches_tracked += 1)."""
    w, b = state[prefix + ".weight"], state[prefix + ".bias"]
    rm, rv = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    dims = [0] + list(range(2, x.dim()))
    shape = [1, -1] + [1] * (x.dim() - 2)
    if training:
        n = x.numel() // x.shape[1]
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        if new_stats is not None:
            with torch.no_grad():
                unb = var * (n / max(n - 1, 1))
                new_stats[prefix + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach().to(rm.dtype)
                new_stats[prefix + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * unb.detach().to(rv.dtype)
                new_stats[prefix + ".num_batches_tracked"] = state[prefix + ".num_batches_tracked"] + 1
    else:
        mean, var = rm, rv
    xhat = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS)
    return xhat * w.view(shape) + b.view(shape)


def _se(state, prefix, x):
    """SELayer, mobilenetv3.py:92-107."""
    y = x.mean(dim=(2, 3))
    y = F.relu(F.linear(y, state[prefix + ".fc.0.weight"], state[prefix + ".fc.0.bias"]))
    y = h_sigmoid(F.linear(y, state[prefix + ".fc.2.weight"], state[prefix + ".fc.2.bias"]))
    return x * y[:, :, None, None]


def extract_features(state, name, x, training=False, new_stats=None, taps=None):
    """features + final 1x1 conv (mobilenetv3.py:199-203, blocks :126-166)."""
    tab = block_table(name)
    x = F.conv2d(x, state["features.0.0.weight"], None, stride=2, padding=1)
    x = h_swish(_bn(state, "features.0.1", x, training, new_stats))
    if taps is not None:
        taps["stem"] = x
    for i, b in enumerate(tab["blocks"], start=1):
        p = f"features.{i}.conv."
        act = h_swish if b["hs"] else F.relu
        inp = x
        j = 0
        if b["expand"]:
            x = F.conv2d(x, state[p + "0.weight"])
            x = act(_bn(state, p + "1", x, training, new_stats))
            j = 3
        x = F.conv2d(x, state[p + f"{j}.weight"], None, stride=b["stride"],
                     padding=(b["k"] - 1) // 2, groups=b["exp"])
        x = _bn(state, p + f"{j + 1}", x, training, new_stats)
        if b["expand"]:
            if b["se"]:
                x = _se(state, p + f"{j + 2}", x)
            x = act(x)
            pw = 7
        else:
            x = act(x)
            if b["se"]:
                x = _se(state, p + f"{j + 3}", x)
            pw = 4
        x = F.conv2d(x, state[p + f"{pw}.weight"])
        x = _bn(state, p + f"{pw + 1}", x, training, new_stats)
        if b["residual"]:
            x = inp + x
        if taps is not None:
            taps[f"block{i}"] = x
    x = F.conv2d(x, state["conv.0.weight"])
    x = h_swish(_bn(state, "conv.1", x, training, new_stats))
    return x


def embed(state, name, x, training=False, new_stats=None, taps=None):
    """extract_features -> avg pool -> classifier (model_builder.py:128-131)."""
    f = extract_features(state, name, x, training, new_stats, taps)
    f = f.mean(dim=(2, 3))
    f = F.linear(f, state["classifier.0.weight"], state["classifier.0.bias"])
    f = h_swish(_bn(state, "classifier.1", f, training, new_stats))
    if taps is not None:
        taps["embed"] = f
    return f


def forward(state, name, x, cats, training=False, dropout_mask=None, new_stats=None, taps=None):
    """ModelWrapper.forward (model_builder.py:126-146).  `dropout_mask` [B, C_f] of {0,1} is the
    keep mask of cls_fc's Dropout(0.5) in training mode (scaled by 2 = 1/(1-p) here); eval mode
    ignores it."""
    f = embed(state, name, x, training, new_stats, taps)
    w = torch.stack([state[f"regressors.{k}.0.weight"] for k in range(MAX_CLASSES)])   # [9,18,C]
    b = torch.stack([state[f"regressors.{k}.0.bias"] for k in range(MAX_CLASSES)])     # [9,18]
    kp = torch.einsum("bjc,bc->bj", w[cats], f) + b[cats]        # head picked per sample (:137)
    kp = torch.sigmoid(kp).view(x.shape[0], NUM_POINTS // 2, 2)
    g = f
    if training:
        assert dropout_mask is not None, "training-mode forward needs the dropout keep mask"
        g = f * dropout_mask[:, : f.shape[1]].to(f.dtype) * 2.0
    logits = F.linear(g, state["cls_fc.1.weight"], state["cls_fc.1.bias"])
    return kp, logits


def forward_export(state, name, x):
    """ModelWrapper.forward_to_onnx (model_builder.py:112-124): all 9 heads, eval mode."""
    f = embed(state, name, x, False)
    outs = []
    for k in range(MAX_CLASSES):
        o = F.linear(f, state[f"regressors.{k}.0.weight"], state[f"regressors.{k}.0.bias"])
        outs.append(o.view(1, x.shape[0], NUM_POINTS // 2, 2))
    kp_all = torch.sigmoid(torch.cat(outs))
    logits = F.linear(f, state["cls_fc.1.weight"], state["cls_fc.1.bias"])
    return kp_all, logits


def select_by_argmax(kp_all, logits):
    """Deployment consumer (utils/ie_wrappers.py:138-142): label = argmax, kp = kp_all[label]."""
    label = torch.argmax(logits, dim=1)
    return kp_all[label, torch.arange(logits.shape[0])], label


# --------------------------------------------------------------------------------------------
# losses (regression_losses.py, loss_builder.py)
# --------------------------------------------------------------------------------------------
def _diag(kp):
    """compute_diag, regression_losses.py:51-58."""
    x0, x1 = kp[:, :, 0].min(dim=1).values, kp[:, :, 0].max(dim=1).values
    y0, y1 = kp[:, :, 1].min(dim=1).values, kp[:, :, 1].max(dim=1).values
    return torch.sqrt((x1 - x0) ** 2 + (y1 - y0) ** 2)


def loss_term(name, pred, gt, cfg):
    if name == "l1":
        return (pred - gt).abs().mean()
    if name == "mse":
        return ((pred - gt) ** 2).mean()
    if name == "smoothl1":
        beta = cfg["smoothl1_beta"]
        d = (pred - gt).abs()
        return torch.where(d < beta, 0.5 * d * d / beta, d - 0.5 * beta).mean()
    if name == "add_loss":                                       # regression_losses.py:22-26
        return torch.linalg.norm(pred - gt, dim=2).sum(dim=1).mean()
    if name == "diag_loss":                                      # regression_losses.py:8-20
        d = (_diag(pred) - _diag(gt)).abs()
        return torch.where(d < 0.4, 0.5 * d * d / 0.4, d - 0.2).mean()
    if name == "wing":                                           # regression_losses.py:28-49
        w, eps = cfg["w"], cfg["eps"]
        const = w - w * math.log(1.0 + w / eps)
        d = (pred - gt).abs()
        small = d < w
        core = w * torch.log(1.0 + d / eps)
        # the reference applies the two masked updates sequentially on the same buffer (:37-38):
        # values rewritten by the first (core) that land >= w get `const` subtracted as well.
        first = torch.where(small, core, d)
        return torch.where(first >= w, first - const, first).mean()
    if name == "cross_entropy":
        return F.cross_entropy(pred, gt)
    raise AssertionError(name)


DEFAULT_LOSS = dict(names=["l1", "add_loss", "cross_entropy"], coeffs=([1.0, 0.1], [0.2]),
                    smoothl1_beta=0.2, w=5.18, eps=1.0)          # configs/default_config.py:22-23


def parse_losses(pred_kp, gt_kp, logits, cats, loss_cfg=None):
    """LossManager.parse_losses without ALWA (regression_losses.py:79-92).
    Returns (total, reg_terms, cls_terms) with the coefficient already applied to each term."""
    cfg = DEFAULT_LOSS if loss_cfg is None else loss_cfg
    reg_names = [n for n in cfg["names"] if n != "cross_entropy"]
    cls_names = [n for n in cfg["names"] if n == "cross_entropy"]
    reg = [loss_term(n, pred_kp, gt_kp, cfg) * k for n, k in zip(reg_names, cfg["coeffs"][0])]
    cls = [loss_term(n, logits, cats, cfg) * k for n, k in zip(cls_names, cfg["coeffs"][1])]
    total = sum(reg)
    if cls:
        total = total + sum(cls)
    return total, reg, cls


# --------------------------------------------------------------------------------------------
# metrics (evaluation/metrics.py:10-68)
# --------------------------------------------------------------------------------------------
@torch.no_grad()
def average_distance(pred_kp, gt_kp, reduce_mean=True):
    d_same = torch.linalg.norm(pred_kp - gt_kp, dim=2)                            # [B,9]
    d_all = torch.linalg.norm(pred_kp[:, :, None, :] - gt_kp[:, None, :, :], dim=3)  # [B,9,9]
    nearest = torch.minimum(d_same, d_all.min(dim=2).values)                      # :15-21
    if reduce_mean:
        return d_same.mean().item(), (nearest.sum(dim=1).mean() / 9).item()
    return (d_same.sum() / 9).item(), (nearest.sum() / 9).item()


@torch.no_grad()
def accuracy(logits, cats, reduce_mean=True):
    hit = (torch.argmax(logits, dim=1) == cats).float()
    return hit.mean().item() if reduce_mean else hit.sum().item()


@torch.no_grad()
def metrics_per_cls(pred_kp, gt_kp, logits, cats):
    """compute_metrics_per_cls with compute_iou=False (IoU branch is out of scope)."""
    rows, tot = [], [0.0, 0.0, 0.0]
    for cl in torch.unique(cats):
        m = cats == cl
        add, sadd = average_distance(pred_kp[m], gt_kp[m], reduce_mean=False)
        acc = accuracy(logits[m], cats[m], reduce_mean=False)
        n = int(m.sum())
        rows.append((int(cl), add / n, sadd / n, 0.0, acc / n))
        tot = [tot[0] + add, tot[1] + sadd, tot[2] + acc]
    bs = pred_kp.shape[0]
    return rows, tot[0] / bs, tot[1] / bs, 0.0, tot[2] / bs


# --------------------------------------------------------------------------------------------
# optimizers (optim_builder.py:5-19 -> torch.optim semantics, restated functionally)
# --------------------------------------------------------------------------------------------
DEFAULT_OPTIM = dict(name="adam", lr=1e-3, momentum=0.9, wd=1e-4, betas=(0.9, 0.999),
                     rho=0.9, alpha=0.99, nesterov=True)         # configs/default_config.py:18


def trainable_keys(state):
    return [k for k in state if not (k.endswith("running_mean") or k.endswith("running_var")
                                     or k.endswith("num_batches_tracked"))]


@torch.no_grad()
def optim_step(state, grads, opt_state, cfg=None):
    """One optimizer step in place.  `grads[k] is None` <=> the reference leaves `.grad = None`
    (regressor heads whose class is absent from the batch, model_builder.py:137) and torch
    optimizers then skip the tensor entirely (no decay, no moment update, no step count)."""
    cfg = DEFAULT_OPTIM if cfg is None else cfg
    lr, wd = cfg["lr"], cfg["wd"]
    for k in trainable_keys(state):
        g = grads.get(k)
        if g is None:
            continue
        p = state[k]
        st = opt_state.setdefault(k, {})
        if cfg["name"] == "adam":                                # torch.optim.AdamW
            b1, b2 = cfg["betas"]
            if not st:
                st.update(step=0, m=torch.zeros_like(p), v=torch.zeros_like(p))
            st["step"] += 1
            p.mul_(1 - lr * wd)
            st["m"].lerp_(g, 1 - b1)
            st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
            bc1 = 1 - b1 ** st["step"]
            bc2 = 1 - b2 ** st["step"]
            denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(1e-8)
            p.addcdiv_(st["m"], denom, value=-lr / bc1)
        elif cfg["name"] == "sgd":                               # torch.optim.SGD
            g = g + wd * p
            mom = cfg["momentum"]
            if mom != 0:
                if "buf" not in st:
                    st["buf"] = g.clone()
                else:
                    st["buf"].mul_(mom).add_(g)
                g = g + mom * st["buf"] if cfg["nesterov"] else st["buf"]
            p.add_(g, alpha=-lr)
        elif cfg["name"] == "rmsprop":                           # torch.optim.RMSprop
            g = g + wd * p
            if not st:
                st.update(sq=torch.zeros_like(p))
            st["sq"].mul_(cfg["alpha"]).addcmul_(g, g, value=1 - cfg["alpha"])
            p.addcdiv_(g, st["sq"].sqrt().add_(1e-8), value=-lr)
        elif cfg["name"] == "adadelta":                          # torch.optim.Adadelta
            g = g + wd * p
            rho = cfg["rho"]
            if not st:
                st.update(sq=torch.zeros_like(p), acc=torch.zeros_like(p))
            st["sq"].mul_(rho).addcmul_(g, g, value=1 - rho)
            delta = (st["acc"] + 1e-6).sqrt() / (st["sq"] + 1e-6).sqrt() * g
            st["acc"].mul_(rho).addcmul_(delta, delta, value=1 - rho)
            p.add_(delta, alpha=-lr)
        else:
            raise AssertionError(cfg["name"])


# --------------------------------------------------------------------------------------------
# one training step (trainer/train.py:46-55)
# --------------------------------------------------------------------------------------------
def train_step(state, name, opt_state, imgs, gt_kp, cats, dropout_mask, loss_cfg=None,
               optim_cfg=None, step_optimizer=True):
    """fwd -> loss -> bwd -> optimizer -> metrics.  Mutates `state` (weights + BN buffers) and
    `opt_state`.  Returns dict(kp, logits, loss, reg_terms, cls_terms, grads, add, sadd, acc)."""
    keys = trainable_keys(state)
    leaves = {k: state[k].detach().clone().requires_grad_(True) for k in keys}
    work = dict(state)
    work.update(leaves)
    new_stats = {}
    kp, logits = forward(work, name, imgs, cats, training=True, dropout_mask=dropout_mask,
                         new_stats=new_stats)
    total, reg, cls = parse_losses(kp, gt_kp, logits, cats, loss_cfg)
    gl = torch.autograd.grad(total, [leaves[k] for k in keys], allow_unused=True)
    grads = dict(zip(keys, gl))
    present = set(int(c) for c in torch.unique(cats))
    for k in range(MAX_CLASSES):                                 # absent head -> grad None
        if k not in present:
            grads[f"regressors.{k}.0.weight"] = None
            grads[f"regressors.{k}.0.bias"] = None
    state.update(new_stats)
    if step_optimizer:
        optim_step(state, grads, opt_state, optim_cfg)
    add, sadd = average_distance(kp.detach(), gt_kp)
    acc = accuracy(logits.detach(), cats)
    return dict(kp=kp.detach(), logits=logits.detach(), loss=float(total.detach()),
                reg_terms=[float(t.detach()) for t in reg], cls_terms=[float(t.detach()) for t in cls],
                grads=grads, add=add, sadd=sadd, acc=acc)


# --------------------------------------------------------------------------------------------
# per-layer shape/MAC table (used for the roofline in bench.py)
# --------------------------------------------------------------------------------------------
def layer_table(name, res=224):
    """[(kind, I_elems, O_elems, W_elems, MACs)] per conv/linear layer, per crop."""
    tab = block_table(name)
    rows = []
    h = res // 2
    rows.append(("stem", 3 * res * res, tab["stem"] * h * h, 27 * tab["stem"], 27 * tab["stem"] * h * h))
    for b in tab["blocks"]:
        if b["expand"]:
            rows.append(("pw", b["cin"] * h * h, b["exp"] * h * h, b["cin"] * b["exp"],
                         b["cin"] * b["exp"] * h * h))
        ho = (h + b["stride"] - 1) // b["stride"]
        rows.append(("dw", b["exp"] * h * h, b["exp"] * ho * ho, b["k"] ** 2 * b["exp"],
                     b["k"] ** 2 * b["exp"] * ho * ho))
        if b["se"]:
            rows.append(("se", 2 * b["exp"], 2 * b["exp"], 2 * b["exp"] * b["se_hidden"],
                         2 * b["exp"] * b["se_hidden"]))
        h = ho
        rows.append(("pw", b["exp"] * h * h, b["cout"] * h * h, b["exp"] * b["cout"],
                     b["exp"] * b["cout"] * h * h))
    rows.append(("pw", tab["last_in"] * h * h, tab["last_exp"] * h * h, tab["last_in"] * tab["last_exp"],
                 tab["last_in"] * tab["last_exp"] * h * h))
    rows.append(("fc", tab["last_exp"], tab["head"], tab["last_exp"] * tab["head"], tab["last_exp"] * tab["head"]))
    rows.append(("fc", tab["head"], 18 + 9, tab["head"] * (18 * 9 + 9), tab["head"] * (18 + 9)))
    return rows
