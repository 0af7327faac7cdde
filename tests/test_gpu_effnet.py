"""EfficientNet-B0 / B3 regressors (BASELINE configs 3 and 5) against the torchvision-through-the-wrapper oracle
(oracle/effnet_port.py): SiLU, sigmoid-gated SE after the activation with input-channel squeeze widths,
32 / 40-channel stem, no classifier, heads on the pooled 1280 / 1536 features."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_port as tp, effnet_port as ep                # noqa: E402
from test_gpu_model import make_cfg, t2n, rel, DEV                    # noqa: E402
from torchdet3d_b200 import InferSession                              # noqa: E402
from torchdet3d_b200.builders import build_model, build_loss, build_optimizer   # noqa: E402
from torchdet3d_b200.evaluation import compute_average_distance, compute_accuracy   # noqa: E402
from torchdet3d_b200.losses import LossManager                       # noqa: E402


def _model(name, dtype="fp32", gemm="auto", optim=None):
    case = dict(model=name, optim=optim or dict(name="sgd", lr=0.01), loss=None)
    cfg = make_cfg(case, dtype, gemm)
    model = build_model(cfg)
    state = ep.synth_state(name, seed=0)
    model.load_state_dict(state)
    return cfg, model.to(DEV), state


@pytest.mark.parametrize("name,B,res", [("efficientnet_b0", 6, 64), ("efficientnet_b3", 4, 96), ("efficientnet_b0", 3, 224)])
def test_train_step_fp32_vs_oracle(name, B, res):
    cfg, model, state = _model(name)
    lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
    opt = build_optimizer(cfg, model)
    imgs, gt_kp, cats, keep = tp.synth_batch(B, res=res, seed=21, all_classes=False)
    keep = keep[:, :ep.MODELS[name]].contiguous()
    ocfg = dict(tp.DEFAULT_OPTIM, name="sgd", lr=0.01)
    r = ep.train_step(state, name, {}, imgs, gt_kp, cats, keep, optim_cfg=ocfg)
    model.train()
    kp, logits = model(imgs.to(DEV), cats.to(DEV), dropout_keep=keep.to(DEV))
    loss = lm.parse_losses(kp, gt_kp.to(DEV), logits, cats.to(DEV), 0)
    opt.zero_grad()
    loss.backward()
    grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in model.named_parameters()}
    opt.step()
    assert rel(t2n(kp), r["kp"].numpy()) < 1e-3 and rel(t2n(logits), r["logits"].numpy()) < 2e-3
    assert abs(loss.item() - r["loss"]) < 1e-3 * abs(r["loss"])
    assert np.array_equal(t2n(logits).argmax(1), r["logits"].numpy().argmax(1))
    add, sadd = compute_average_distance(kp, gt_kp.to(DEV))
    assert add == pytest.approx(r["add"], rel=1e-4) and sadd == pytest.approx(r["sadd"], rel=1e-4)
    assert compute_accuracy(logits, cats.to(DEV)) == pytest.approx(r["acc"], abs=1e-6)
    assert [n for n in grads if grads[n] is None] == [n for n in r["grads"] if r["grads"][n] is None]
    num = den = 0.0
    worst = []
    for n, g in grads.items():
        gr = r["grads"][n]
        if gr is None:
            continue
        num += float((g.cpu().double() - gr.double()).pow(2).sum())
        den += float(gr.double().pow(2).sum())
        if gr.norm() / gr.numel() ** 0.5 > 1e-6:
            worst.append((rel(t2n(g), gr.numpy()), n))
    worst.sort(reverse=True)
    assert (num / den) ** 0.5 < 2e-3, ((num / den) ** 0.5, worst[:6])
    assert worst[0][0] < 5e-2, worst[:6]
    sd = model.state_dict()
    for k, v in state.items():                       # post-step weights and BatchNorm buffers
        if k.endswith("num_batches_tracked"):
            assert int(sd[k]) == int(v), k
        elif k.endswith("running_mean") or k.endswith("running_var"):
            np.testing.assert_allclose(t2n(sd[k]), v.numpy(), rtol=5e-3, atol=1e-5, err_msg=k)
        else:
            assert rel(t2n(sd[k]), v.numpy()) < 2e-3, k


@pytest.mark.parametrize("name,B,res", [("efficientnet_b0", 10, 96), ("efficientnet_b3", 5, 128)])
def test_eval_export_and_infer_session_fp32(name, B, res):
    _, model, state = _model(name)
    x = torch.rand(B, 3, res, res, generator=torch.Generator().manual_seed(9))
    kp_all_ref, logits_ref = ep.forward_export(state, name, x)
    ref_sel, ref_lab = tp.select_by_argmax(kp_all_ref, logits_ref)
    model.eval()
    kp_all, logits = model.forward_to_onnx(x.to(DEV))
    assert rel(t2n(kp_all), kp_all_ref.numpy()) < 1e-3 and rel(t2n(logits), logits_ref.numpy()) < 1e-3
    sess = InferSession(model, B, res, res, chunk=4)
    for _ in range(3):
        kp, labels, _ = sess(x)
    assert np.array_equal(labels.cpu().numpy(), ref_lab.numpy()) and rel(t2n(kp), ref_sel.numpy()) < 1e-3
    cats = torch.arange(B) % 9
    with torch.no_grad():
        kp_e, lg_e = model(x.to(DEV), cats.to(DEV))
    ref = ep.make(name, state).eval()
    with torch.no_grad():
        kp_r, lg_r = ref(x, cats)
    assert rel(t2n(kp_e), kp_r.numpy()) < 1e-3 and rel(t2n(lg_e), lg_r.numpy()) < 1e-3


def test_efficientnet_b0_bf16_tcgen05_vs_oracle():
    """bf16 storage + tcgen05 GEMMs; stated bf16 bounds: kp <= 5e-2 (train-mode BN), loss <= 2e-2 rel."""
    name, B, res = "efficientnet_b0", 16, 128
    cfg, model, state = _model(name, "bf16", "auto")
    lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
    imgs, gt_kp, cats, keep = tp.synth_batch(B, res=res, seed=22, all_classes=True)
    keep = keep[:, :ep.MODELS[name]].contiguous()
    r = ep.train_step(state, name, {}, imgs, gt_kp, cats, keep, step_optimizer=False)
    model.train()
    kp, logits = model(imgs.to(DEV), cats.to(DEV), dropout_keep=keep.to(DEV))
    loss = lm.parse_losses(kp, gt_kp.to(DEV), logits, cats.to(DEV), 0)
    loss.backward()
    assert rel(t2n(kp), r["kp"].numpy()) < 5e-2
    assert abs(loss.item() - r["loss"]) < 2e-2 * abs(r["loss"])
    num = sum(float((p.grad.cpu().double() - r["grads"][n].double()).pow(2).sum()) for n, p in model.named_parameters() if p.grad is not None)
    den = sum(float(r["grads"][n].double().pow(2).sum()) for n, p in model.named_parameters() if p.grad is not None)
    assert (num / den) ** 0.5 < 0.2, (num / den) ** 0.5
