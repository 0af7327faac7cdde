"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by executing the UNMODIFIED reference
(`/root/reference`, imported through oracle/refshim.py) on seeded synthetic inputs.

Run in the build container (the reference tree does not exist on the GPU box):

    python oracle/make_golden.py

Every fixture stores only what cannot be regenerated without the reference: the reference's
outputs.  Inputs/weights are regenerated from seeds by oracle.torch_port.synth_state/synth_batch
(numpy RNG, stable across machines).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim, torch_port as tp  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# tensors whose full gradient / post-step value is stored (small ones, covering every layer type)
FULL_KEYS_COMMON = ["features.0.0.weight", "features.0.1.weight", "features.0.1.bias",
                    "features.1.conv.0.weight", "conv.1.weight", "classifier.1.bias",
                    "cls_fc.1.bias", "cls_fc.1.weight"]

LOSS_ALL = dict(names=["smoothl1", "l1", "mse", "wing", "add_loss", "diag_loss", "cross_entropy"],
                coeffs=([1.0, 0.5, 2.0, 0.3, 0.1, 0.7], [0.2]), smoothl1_beta=0.2, w=0.3, eps=0.5)

CASES = [
    # tag, model, batch, res, optimizer cfg override, loss cfg, all_classes, steps
    ("small_adamw", "mobilenetv3_small", 6, 64, dict(name="adam"), None, False, 2),
    ("small_sgd_allloss", "mobilenetv3_small", 12, 64, dict(name="sgd", lr=0.05), LOSS_ALL, True, 2),
    ("large_adamw", "mobilenetv3_large", 4, 64, dict(name="adam"), None, False, 2),
    ("small_224_adamw", "mobilenetv3_small", 3, 224, dict(name="adam"), None, False, 1),
]


def _ref_cfg(model, optim_over, loss_cfg):
    cfg = refshim.reference_config(model)
    for k, v in optim_over.items():
        cfg.optim[k] = v
    if loss_cfg is not None:
        cfg.loss.names = list(loss_cfg["names"])
        cfg.loss.coeffs = (list(loss_cfg["coeffs"][0]), list(loss_cfg["coeffs"][1]))
        cfg.loss.smoothl1_beta = loss_cfg["smoothl1_beta"]
        cfg.loss.w = loss_cfg["w"]
        cfg.loss.eps = loss_cfg["eps"]
    return cfg


def dropout_mask_for(seed, batch, width):
    """Keep-mask nn.Dropout(0.5) draws as first RNG consumer after torch.manual_seed(seed)."""
    torch.manual_seed(seed)
    return (F.dropout(torch.ones(batch, width), 0.5, True) > 0).float()


def run_case(tag, model, batch, res, optim_over, loss_cfg, all_classes, steps):
    from torchdet3d.builders import build_model, build_loss, build_optimizer
    from torchdet3d.losses import LossManager
    from torchdet3d.evaluation import (compute_average_distance, compute_accuracy,
                                       compute_metrics_per_cls)
    cfg = _ref_cfg(model, optim_over, loss_cfg)
    net = build_model(cfg)
    net.load_state_dict(tp.synth_state(model, seed=0))
    width = net.cls_fc[1].in_features
    out = {}

    # ---- eval-mode forward + export-mode forward + per-class metrics ----
    imgs, gt_kp, cats, _ = tp.synth_batch(batch, res=res, seed=77, all_classes=all_classes)
    net.eval()
    with torch.no_grad():
        kp, logits = net(imgs, cats)
    out["eval_kp"], out["eval_logits"] = kp.numpy(), logits.numpy()
    rows, add, sadd, _, acc = compute_metrics_per_cls(kp, gt_kp, logits, cats, compute_iou=False)
    out["eval_percls"] = np.array([[float(r[0]), r[1], r[2], r[4]] for r in rows], dtype=np.float64)
    out["eval_metrics"] = np.array([add, sadd, acc], dtype=np.float64)
    enet = build_model(cfg, export_mode=True)
    enet.load_state_dict(tp.synth_state(model, seed=0))
    enet.eval()
    with torch.no_grad():
        kp_all, elog = enet(imgs)
    out["export_kp_all"], out["export_logits"] = kp_all.numpy(), elog.numpy()

    # ---- training steps (trainer/train.py:46-55) ----
    criterions = build_loss(cfg)
    lm = LossManager(criterions, cfg.loss.coeffs, cfg.loss.alwa)
    opt = build_optimizer(cfg, net)
    net.train()
    names = [n for n, _ in net.named_parameters()]
    for step in range(steps):
        imgs, gt_kp, cats, _ = tp.synth_batch(batch, res=res, seed=1000 + step, all_classes=all_classes)
        seed = 500 + step
        mask = dropout_mask_for(seed, batch, width)
        torch.manual_seed(seed)
        kp, logits = net(imgs, cats)
        loss = lm.parse_losses(kp, gt_kp, logits, cats, step)
        opt.zero_grad()
        loss.backward()
        grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in net.named_parameters()}
        opt.step()
        add, sadd = compute_average_distance(kp, gt_kp)
        acc = compute_accuracy(logits, cats)
        s = f"s{step}_"
        out[s + "mask"] = np.packbits(mask.numpy().astype(np.uint8), axis=1)
        out[s + "kp"], out[s + "logits"] = kp.detach().numpy(), logits.detach().numpy()
        out[s + "loss"] = np.array([loss.item()], dtype=np.float64)
        out[s + "metrics"] = np.array([add, sadd, acc], dtype=np.float64)
        out[s + "grad_none"] = np.array([grads[n] is None for n in names])
        out[s + "grad_l2"] = np.array([0.0 if grads[n] is None else grads[n].double().norm().item()
                                       for n in names])
        out[s + "grad_sum"] = np.array([0.0 if grads[n] is None else grads[n].double().sum().item()
                                        for n in names])
        sd = net.state_dict()
        out[s + "param_l2"] = np.array([sd[n].double().norm().item() for n in names])
        out[s + "param_sum"] = np.array([sd[n].double().sum().item() for n in names])
        full = list(FULL_KEYS_COMMON) + [f"regressors.{k}.0.bias" for k in range(9)]
        for n in full:
            if grads[n] is not None:
                out[s + "grad/" + n] = grads[n].numpy()
            out[s + "param/" + n] = sd[n].detach().numpy().copy()
        for n in ["features.0.1.running_mean", "features.0.1.running_var", "conv.1.running_mean",
                  "conv.1.running_var", "classifier.1.running_mean", "classifier.1.running_var",
                  "features.2.conv.4.running_var", "classifier.1.num_batches_tracked"]:
            out[s + "buf/" + n] = sd[n].detach().numpy().copy()
    out["param_names"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print(tag, "->", sum(v.nbytes for v in out.values()) // 1024, "KiB (uncompressed)")


def run_loss_metric_golden():
    """Loss values + gradients and metrics on the reference tests' own shapes
    (tests/test_pipeline.py:12-15,24-30)."""
    from torchdet3d.losses import WingLoss, ADD_loss, DiagLoss
    from torchdet3d.evaluation import compute_average_distance, compute_accuracy, compute_metrics_per_cls
    rng = np.random.default_rng(42)
    out = {}
    for B in (128, 512, 1):
        pred = torch.tensor(rng.random((B, 9, 2), dtype=np.float32), requires_grad=True)
        gt = torch.tensor(rng.random((B, 9, 2), dtype=np.float32))
        logits = torch.tensor(rng.standard_normal((B, 9)).astype(np.float32), requires_grad=True)
        cats = torch.tensor(rng.integers(0, 9, size=(B,)), dtype=torch.int64)
        out[f"B{B}_pred"], out[f"B{B}_gt"] = pred.detach().numpy(), gt.numpy()
        out[f"B{B}_logits"], out[f"B{B}_cats"] = logits.detach().numpy(), cats.numpy()
        crits = dict(l1=torch.nn.L1Loss(), mse=torch.nn.MSELoss(),
                     smoothl1=torch.nn.SmoothL1Loss(beta=0.2), add_loss=ADD_loss(),
                     diag_loss=DiagLoss(), wing=WingLoss(w=0.3, eps=0.5),
                     wing_default=WingLoss(), wing_cfg=WingLoss(w=5.18, eps=1.0))
        for n, c in crits.items():
            pred.grad = None
            v = c(pred, gt)
            v.backward()
            out[f"B{B}_{n}"] = np.array([v.item()], dtype=np.float64)
            out[f"B{B}_{n}_grad"] = pred.grad.numpy().copy()
        v = torch.nn.CrossEntropyLoss()(logits, cats)
        v.backward()
        out[f"B{B}_cross_entropy"] = np.array([v.item()], dtype=np.float64)
        out[f"B{B}_cross_entropy_grad"] = logits.grad.numpy().copy()
        add, sadd = compute_average_distance(pred.detach(), gt)
        acc = compute_accuracy(logits.detach(), cats)
        out[f"B{B}_metrics"] = np.array([add, sadd, acc], dtype=np.float64)
        add, sadd = compute_average_distance(pred.detach(), gt, reduce_mean=False)
        acc = compute_accuracy(logits.detach(), cats, reduce_mean=False)
        out[f"B{B}_metrics_sum"] = np.array([add, sadd, acc], dtype=np.float64)
        rows, add, sadd, _, acc = compute_metrics_per_cls(pred.detach(), gt, logits.detach(), cats,
                                                          compute_iou=False)
        out[f"B{B}_percls"] = np.array([[float(r[0]), r[1], r[2], r[4]] for r in rows], dtype=np.float64)
        out[f"B{B}_percls_tot"] = np.array([add, sadd, acc], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "loss_metrics.npz"), **out)
    print("loss_metrics ->", sum(v.nbytes for v in out.values()) // 1024, "KiB")


def run_alwa_golden():
    """The reference LossManager with ALWA re-weighting on (regression_losses.py:96-115): returned loss, its gradients
    and lam_cls for 9 iterations with C=3 (updates at iterations 3 and 6), both variants (with / without std)."""
    from torchdet3d.builders import build_loss
    from torchdet3d.losses import LossManager
    out = {}
    for ver, compute_std in (("v1", True), ("v2", False)):
        cfg = refshim.reference_config("mobilenetv3_small")
        cfg.loss.names = ["l1", "add_loss", "cross_entropy"]
        cfg.loss.coeffs = ([1.0, 0.1], [1.0])
        cfg.loss.alwa = dict(use=True, lam_cls=1., lam_reg=1., C=3, compute_std=compute_std)
        lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
        rng = np.random.default_rng(7)
        B = 16
        for it in range(9):
            pred = torch.tensor(rng.random((B, 9, 2), dtype=np.float32), requires_grad=True)
            gt = torch.tensor(rng.random((B, 9, 2), dtype=np.float32))
            logits = torch.tensor((rng.standard_normal((B, 9)) * (3.0 if it < 5 else 0.3)).astype(np.float32), requires_grad=True)
            cats = torch.tensor(rng.integers(0, 9, size=(B,)), dtype=torch.int64)
            loss = lm.parse_losses(pred, gt, logits, cats, it)
            loss.backward()
            k = f"{ver}_i{it}_"
            out[k + "pred"], out[k + "gt"] = pred.detach().numpy(), gt.numpy()
            out[k + "logits"], out[k + "cats"] = logits.detach().numpy(), cats.numpy()
            out[k + "loss"] = np.array([loss.item()], dtype=np.float64)
            out[k + "lam_cls"] = np.array([float(lm.lam_cls)], dtype=np.float64)
            out[k + "g_pred"], out[k + "g_logits"] = pred.grad.numpy().copy(), logits.grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "alwa.npz"), **out)
    print("alwa ->", sum(v.nbytes for v in out.values()) // 1024, "KiB; lam_cls trajectory",
          [round(float(out[f"v1_i{i}_lam_cls"][0]), 4) for i in range(9)])


def run_effnet_golden():
    """torchvision efficientnet_b0 `.features` through the REFERENCE's own model_wrapper: pins
    oracle/effnet_port.py (which restates the wrapper because /root/reference does not travel to the GPU box)."""
    from torchdet3d.builders.model_builder import model_wrapper
    from torchdet3d.builders import build_loss
    from torchdet3d.losses import LossManager
    from oracle import effnet_port as ep
    name, B, res = "efficientnet_b0", 5, 64

    class TVFeat(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.features = ep.features(name)

        def extract_features(self, x):
            return self.features(x)

    net = model_wrapper(model_class=TVFeat, output_channels=1280, num_classes=9)
    state = ep.synth_state(name, seed=0)
    net.load_state_dict(state)
    out = {}
    imgs, gt_kp, cats, _ = tp.synth_batch(B, res=res, seed=77)
    net.eval()
    with torch.no_grad():
        kp, logits = net(imgs, cats)
        kp_all, elog = net.forward_to_onnx(imgs)
    out["eval_kp"], out["eval_logits"] = kp.numpy(), logits.numpy()
    out["export_kp_all"], out["export_logits"] = kp_all.numpy(), elog.numpy()
    cfg = refshim.reference_config("mobilenetv3_small")
    lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
    net.train()
    imgs, gt_kp, cats, _ = tp.synth_batch(B, res=res, seed=1000)
    mask = dropout_mask_for(500, B, 1280)
    torch.manual_seed(500)
    kp, logits = net(imgs, cats)
    loss = lm.parse_losses(kp, gt_kp, logits, cats, 0)
    loss.backward()
    out["s0_mask"] = np.packbits(mask.numpy().astype(np.uint8), axis=1)
    out["s0_kp"], out["s0_logits"] = kp.detach().numpy(), logits.detach().numpy()
    out["s0_loss"] = np.array([loss.item()], dtype=np.float64)
    names = [n for n, _ in net.named_parameters()]
    out["param_names"] = np.array(names)
    out["s0_grad_none"] = np.array([p.grad is None for _, p in net.named_parameters()])
    out["s0_grad_l2"] = np.array([0.0 if p.grad is None else p.grad.double().norm().item() for _, p in net.named_parameters()])
    sd = net.state_dict()
    for n in ["features.0.1.running_mean", "features.1.0.block.0.1.running_var", "features.8.1.running_var"]:
        out["s0_buf/" + n] = sd[n].detach().numpy().copy()
    for n in ["features.0.0.weight", "features.1.0.block.1.fc1.weight", "features.2.0.block.2.fc2.bias", "features.8.1.weight",
              "cls_fc.1.weight"]:
        out["s0_grad/" + n] = dict(net.named_parameters())[n].grad.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "effnet_b0.npz"), **out)
    print("effnet_b0 ->", sum(v.nbytes for v in out.values()) // 1024, "KiB")


def run_iou_golden():
    """the reference's own lift_2d (utils/geometry.py:51-108) and compute_2d_based_iou (evaluation/metrics.py:70-89,
    Objectron box / iou vendored under 3rdparty/) on jittered copies of the known-answer keypoints of tests/test_geometry.py:13-21,
    plus the edge cases: identical sets, disjoint boxes, a degenerate (all-equal) set."""
    sys.path.insert(0, os.path.join(refshim.REFERENCE_ROOT, "3rdparty", "Objectron"))
    from torchdet3d.utils import lift_2d
    from objectron.dataset import box as obox, iou as oiou
    base = np.array([[0.47714591, 0.47491544], [0.73884577, 0.39749265], [0.18508956, 0.40002537], [0.74114597, 0.48664019],
                     [0.18273196, 0.48833901], [0.64639187, 0.46719882], [0.32766378, 0.46827659], [0.64726073, 0.51853681],
                     [0.32699507, 0.51933688]])
    rng = np.random.default_rng(1234)
    pred, gt = [], []
    for i in range(96):
        sig = [0.002, 0.01, 0.03, 0.08][i % 4]
        gt.append(np.clip(base + rng.normal(0, 0.02, base.shape), 0, 1))
        pred.append(np.clip(gt[-1] + rng.normal(0, sig, base.shape), 0, 1))
    pred.append(base.copy()); gt.append(base.copy())                                  # identical -> IoU 1
    pred.append(np.clip(base * 0.3, 0, 1)); gt.append(np.clip(base * 0.3 + 0.6, 0, 1))  # far apart
    pred.append(np.full_like(base, 0.5)); gt.append(base.copy())                        # degenerate prediction
    pred.append(rng.random(base.shape)); gt.append(rng.random(base.shape))              # arbitrary (non-box) keypoints
    pred, gt = np.array(pred, dtype=np.float32), np.array(gt, dtype=np.float32)
    lifted_p, lifted_g, ious = [], [], []
    for p_, g_ in zip(pred, gt):
        l = lift_2d([p_, g_], portrait=True)
        lifted_p.append(l[0]); lifted_g.append(l[1])
        try:
            ious.append(oiou.IoU(obox.Box(vertices=l[0]), obox.Box(vertices=l[1])).iou())
        except Exception:                      # the reference swallows QhullError / LinAlgError and adds 0 (metrics.py:83-86)
            ious.append(0.0)
    out = dict(pred=pred, gt=gt, lifted_pred=np.array(lifted_p), lifted_gt=np.array(lifted_g), iou=np.array(ious),
               lifted_landscape=np.array(lift_2d([base], portrait=False)[0]))
    np.savez_compressed(os.path.join(OUT, "iou.npz"), **out)
    print("iou ->", len(ious), "pairs, mean IoU", float(np.mean(ious)), "min", float(np.min(ious)), "max", float(np.max(ious)))


def main():
    refshim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    only = sys.argv[1] if len(sys.argv) > 1 else None
    if only in (None, "alwa"):
        run_alwa_golden()
    if only in (None, "loss"):
        run_loss_metric_golden()
    if only in (None, "effnet"):
        run_effnet_golden()
    if only in (None, "iou"):
        run_iou_golden()
    for case in CASES:
        if only in (None, case[0]):
            run_case(*case)


if __name__ == "__main__":
    main()
