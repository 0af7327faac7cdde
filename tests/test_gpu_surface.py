"""GPU tests of the drop-in surface around the hot path: every optimizer kind against the oracle, ALWA against a
golden produced by the reference LossManager, checkpoint formats (torch.optim layout, reference-style snapshots),
eval-mode BatchNorm fold freshness after CUDA-graph training, epoch meters with a partial last batch, and
Evaluator.val(compute_iou=True) / visual_test() as an unmodified scripts/main.py calls them."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import torch_port as tp                                   # noqa: E402
from _cases import CASES, GOLDEN, train_batch, eval_batch             # noqa: E402
from test_gpu_model import make_cfg, make_model, t2n, rel, DEV        # noqa: E402
from torchdet3d_b200 import _lib as L                                 # noqa: E402
from torchdet3d_b200.builders import build_loss, build_optimizer, build_scheduler   # noqa: E402
from torchdet3d_b200.evaluation import Evaluator, set_iou_backend    # noqa: E402
from torchdet3d_b200.losses import LossManager                       # noqa: E402
from torchdet3d_b200.trainer import Trainer, FusedTrainStep          # noqa: E402
from torchdet3d_b200.utils import Dict, resume_from, save_snap       # noqa: E402


@pytest.mark.parametrize("name,over", [("sgd", dict(lr=0.05)), ("sgd", dict(lr=0.05, nesterov=False, momentum=0.0)),
                                        ("adam", {}), ("rmsprop", dict(lr=0.01)), ("adadelta", dict(lr=1.0))])
def test_optimizer_kernels_vs_oracle(name, over):
    """td3d_optim_step (one launch over the flat arena) against oracle/torch_port.optim_step (= torch.optim semantics,
    optim_builder.py:5-19) on well-conditioned random gradients, 3 steps, with two heads absent in step 1."""
    case = dict(CASES["small_adamw"], optim=dict(name=name, **over))
    cfg, model = make_model(case)
    opt = build_optimizer(cfg, model)
    ocfg = dict(tp.DEFAULT_OPTIM, name=name, **over)
    imgs, _, cats, _ = train_batch(case, 0)
    model.train()
    with torch.no_grad():
        model(imgs.to(DEV), cats.to(DEV))           # creates / binds the plan the optimizer steps through
    state = {k: v.clone() for k, v in tp.synth_state(case["model"], seed=0).items()}
    ostate = {}
    gen = torch.Generator().manual_seed(3)
    for step in range(3):
        present = [1] * 9
        if step == 1:
            present[2] = present[7] = 0
        grads = {}
        flat = torch.zeros_like(model._gflat, device="cpu")
        for pname, off, numel, shape in model._param_table:
            g = torch.randn(shape, generator=gen) * 0.1
            flat[off:off + numel] = g.reshape(-1)
            grads[pname] = g
            if pname.startswith("regressors.") and not present[int(pname.split(".")[1])]:
                grads[pname] = None
        model._gflat.copy_(flat)
        model.present.copy_(torch.tensor(present, dtype=torch.int32))
        opt.step()
        tp.optim_step(state, grads, ostate, ocfg)
        sd = model.state_dict()
        worst = max((float((sd[n].cpu() - state[n]).abs().max() / state[n].abs().max().clamp_min(1e-6)), n)
                    for n in tp.trainable_keys(state))
        assert worst[0] < 2e-5, (name, step, worst)
    assert opt.steps.tolist() == [3, 3, 3, 2, 3, 3, 3, 3, 2, 3]


def test_alwa_vs_reference_golden():
    """LossManager with ALWA on against the reference's own LossManager (golden from oracle/make_golden.py alwa):
    returned loss, gradients and lam_cls per iteration, including the two iterations where lam_cls changes and the
    reference already applies the new value to the returned loss (regression_losses.py:96-115)."""
    g = np.load(GOLDEN + "/alwa.npz")
    for ver, compute_std in (("v1", True), ("v2", False)):
        cfg = Dict(loss=dict(names=["l1", "add_loss", "cross_entropy"], coeffs=([1.0, 0.1], [1.0]), smoothl1_beta=0.2,
                             w=0.3, eps=0.5, alwa=dict(use=True, lam_cls=1., lam_reg=1., C=3, compute_std=compute_std)))
        lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
        changed = 0
        for it in range(9):
            k = f"{ver}_i{it}_"
            pred = torch.tensor(g[k + "pred"], device=DEV, requires_grad=True)
            logits = torch.tensor(g[k + "logits"], device=DEV, requires_grad=True)
            before = lm.lam_cls
            loss = lm.parse_losses(pred, torch.tensor(g[k + "gt"], device=DEV), logits, torch.tensor(g[k + "cats"], device=DEV), it)
            loss.backward()
            changed += int(lm.lam_cls != before)
            assert abs(lm.lam_cls - g[k + "lam_cls"][0]) < 2e-5 * max(1.0, abs(g[k + "lam_cls"][0])), (ver, it)
            assert abs(loss.item() - g[k + "loss"][0]) < 1e-5 * abs(g[k + "loss"][0]), (ver, it, loss.item(), g[k + "loss"][0])
            assert rel(t2n(pred.grad), g[k + "g_pred"]) < 1e-4, (ver, it)
            assert rel(t2n(logits.grad), g[k + "g_logits"]) < 1e-4, (ver, it)
        assert changed >= 1


def _ref_style_snapshot(path, name, steps=2):
    """A snapshot in the reference's format (utils/utils.py:56-64): nn.Module.state_dict() keys + torch.optim.AdamW
    state_dict + scheduler + epoch, written with torch.optim itself; returns the post-step oracle state too."""
    state = tp.synth_state(name, seed=0)
    keys = tp.trainable_keys(state)
    params = [torch.nn.Parameter(state[k].clone()) for k in keys]
    o = tp.DEFAULT_OPTIM
    opt = torch.optim.AdamW(params, lr=o["lr"], betas=o["betas"], weight_decay=o["wd"])
    gen = torch.Generator().manual_seed(11)
    for _ in range(steps):
        for p in params:
            p.grad = torch.randn(p.shape, generator=gen) * 0.05
        opt.step()
    sd = dict(state)
    sd.update({k: p.detach().clone() for k, p in zip(keys, params)})
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[60, 90, 120], gamma=0.6)
    torch.save({'state_dict': {"module." + k: v for k, v in sd.items()}, 'optimizer': opt.state_dict(),
                'scheduler': sched.state_dict(), 'epoch': 4}, path)
    return sd, params, opt, keys


def test_reference_snapshot_resumes_and_roundtrips(tmp_path):
    case = CASES["small_adamw"]
    path = str(tmp_path / "ref_snap.pth")
    sd, params, topt, keys = _ref_style_snapshot(path, case["model"])
    cfg, model = make_model(case)
    opt = build_optimizer(cfg, model)
    sched = build_scheduler(cfg, opt)
    assert resume_from(model, path, optimizer=opt, scheduler=sched) == 5          # 'module.' prefix stripped, epoch + 1
    msd = model.state_dict()
    for k in sd:
        assert torch.equal(msd[k].cpu(), sd[k]), k
    # one more step with identical gradients: the flat-arena AdamW must continue exactly where torch.optim.AdamW was
    imgs, _, cats, _ = train_batch(case, 0)
    model.train()
    with torch.no_grad():
        model(imgs.to(DEV), cats.to(DEV))
    model.load_state_dict({k: v for k, v in sd.items()})                          # undo the running-stat update of that forward
    gen = torch.Generator().manual_seed(12)
    flat = torch.zeros_like(model._gflat, device="cpu")
    table = {n: (off, numel) for n, off, numel, _ in model._param_table}
    for k, p in zip(keys, params):
        p.grad = torch.randn(p.shape, generator=gen) * 0.05
        off, numel = table[k]
        flat[off:off + numel] = p.grad.reshape(-1)
    model._gflat.copy_(flat)
    model.present.fill_(1)
    opt.step()
    topt.step()
    msd = model.state_dict()
    worst = max((float((msd[k].cpu() - p.detach()).abs().max()), k) for k, p in zip(keys, params))
    assert worst[0] < 2e-6, worst
    # our snapshot -> torch.optim.AdamW accepts it (same per-parameter layout) and our own loader round-trips it
    snap = save_snap(model, opt, sched, 7, str(tmp_path))
    ck = torch.load(snap, map_location="cpu", weights_only=False)
    t2 = torch.optim.AdamW([torch.nn.Parameter(torch.zeros_like(p)) for p in params], lr=1e-3)
    t2.load_state_dict({k: v for k, v in ck['optimizer'].items() if k != 'td3d_steps'})
    st = t2.state_dict()['state']
    assert int(st[0]['step']) == 3 and torch.allclose(st[0]['exp_avg'], topt.state_dict()['state'][0]['exp_avg'], atol=1e-7)
    cfg2, model2 = make_model(case)
    opt2 = build_optimizer(cfg2, model2)
    assert resume_from(model2, snap, optimizer=opt2) == 8
    assert torch.equal(opt2.state0, opt.state0) and torch.equal(opt2.state1, opt.state1) and torch.equal(opt2.steps, opt.steps)
    with pytest.raises(RuntimeError):
        opt2.load_state_dict({'state': {10 ** 6: {}}, 'param_groups': ck['optimizer']['param_groups']})


def test_eval_fold_is_fresh_after_graph_replays():
    """graph replays move BN parameters and running statistics without running any Python; the
    eval-mode fold must still be rebuilt.  Checked against a FRESH model loaded from state_dict() (always folds)."""
    case = CASES["small_adamw"]
    cfg, model = make_model(case)
    lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
    opt = build_optimizer(cfg, model)
    step = FusedTrainStep(model, lm, opt, case["batch"], case["res"], case["res"], use_graph=True)
    ximgs, _, xcats, _ = eval_batch(case)
    prev = None
    for rnd in range(3):
        for it in range(4):
            imgs, gt_kp, cats, _ = train_batch(case, it % 2)
            step(imgs, gt_kp, cats)
        assert rnd == 0 or step._graph is not None
        model.eval()
        with torch.no_grad():
            kp, logits = model(ximgs.to(DEV), xcats.to(DEV))
        _, fresh = make_model(case)
        fresh.load_state_dict(model.state_dict())
        fresh.eval()
        with torch.no_grad():
            kp2, logits2 = fresh(ximgs.to(DEV), xcats.to(DEV))
        torch.testing.assert_close(kp, kp2, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(logits, logits2, rtol=1e-5, atol=1e-5)
        assert prev is None or not torch.allclose(prev, kp, atol=1e-7)      # training did move the model between evals
        prev = kp.clone()
        model.train()


def test_trainer_epoch_meters_with_partial_last_batch(tmp_path):
    case = CASES["small_adamw"]
    cfg, model = make_model(case)
    lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
    opt = build_optimizer(cfg, model)
    full = [tuple(t for t in train_batch(case, s)[:3]) for s in range(3)]
    part = tuple(t[:3] for t in train_batch(case, 3)[:3])
    trainer = Trainer(model=model, train_loader=full + [part], optimizer=opt, scheduler=None, loss_manager=lm, writer=None,
                      max_epoch=1, log_path=str(tmp_path), device=DEV, save_chkpt=False, print_freq=100)
    trainer.train(0, True)
    n = 3 * case["batch"] + 3
    assert trainer.meters["loss"].count == n and trainer.meters["ACC"].count == n
    assert np.isfinite(trainer.meters["loss"].avg) and 0 <= trainer.meters["ADD"].avg <= 2


def test_evaluator_runs_as_reference_main_calls_it(tmp_path):
    """scripts/main.py:105-106: evaluator.val(epoch, is_last_epoch) with compute_iou=True on the last epoch, then
    evaluator.visual_test().  Neither may raise without the reference's CPU IoU / drawing code."""
    case = CASES["small_adamw"]
    cfg, model = make_model(case)
    batches = [tuple(t for t in train_batch(case, s)[:3]) for s in range(2)]
    ev = Evaluator(model=model, val_loader=batches, test_loader=None, cfg=cfg, writer=None, max_epoch=2, device=DEV)
    set_iou_backend(None)                      # the built-in CUDA kernel (td3d_iou_2d_based)
    with warnings.catch_warnings(record=True):
        warnings.simplefilter("always")
        r = ev.val(1, True)
        assert ev.visual_test() is None        # drawing is out of scope: delegates to the reference or warns, never raises
    assert 0 <= r["ADD"] <= 2 and len(r["per_class"]) == 9
    # the IOU column against the oracle (oracle/iou_port.py restates lift_2d + Objectron IoU) on the same predictions
    from oracle import iou_port
    model.eval()
    tot, n = 0.0, 0
    with torch.no_grad():
        for imgs, gt_kp, gt_cats in batches:
            imgs, gt_kp, gt_cats = imgs.to(DEV), gt_kp.to(DEV), gt_cats.to(DEV)
            pred_kp, _ = model(imgs, gt_cats)
            tot += iou_port.compute_2d_based_iou(pred_kp.float().cpu().numpy(), gt_kp.float().cpu().numpy(), reduce_mean=False)
            n += imgs.shape[0]
    assert abs(r["IOU"] - tot / n) < 1e-4, (r["IOU"], tot / n)
    calls = []

    def fake_iou(pred_kp, gt_kp, reduce_mean=True):
        calls.append(pred_kp.shape[0])
        return 0.5 * pred_kp.shape[0]

    set_iou_backend(fake_iou)
    try:
        r = ev.val(1, compute_iou=True)
    finally:
        set_iou_backend(None)
    assert abs(r["IOU"] - 0.5) < 1e-9 and sum(calls) == 2 * case["batch"]


def test_single_class_model_leaves_cls_fc_untouched():
    """num_classes == 1 (model_builder.py:140-144, regression_losses.py:84-88): forward returns the categories themselves,
    LossManager runs without class criteria, and cls_fc -- never executed by the reference -- keeps grad=None, so the
    optimizer must not decay or move it."""
    case = CASES["small_adamw"]
    cfg = make_cfg(case)
    cfg.model.num_classes = 1
    cfg.loss.names = ["l1", "add_loss"]
    cfg.loss.coeffs = ([1.0, 0.1], [])
    from torchdet3d_b200.builders import build_model
    model = build_model(cfg)
    state = tp.synth_state(case["model"], seed=0, num_classes=1)
    model.load_state_dict(state)
    model = model.to(DEV).train()
    lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
    opt = build_optimizer(cfg, model)
    imgs, gt_kp, cats, _ = train_batch(case, 0)
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    kp, targets = model(imgs.to(DEV), cats.to(DEV))
    assert targets.shape == (case["batch"], 1) and torch.equal(targets.cpu().view(-1), cats)
    loss = lm.parse_losses(kp, gt_kp.to(DEV), targets, cats.to(DEV), 0)
    opt.zero_grad()
    loss.backward()
    opt.step()
    for n, p in model.named_parameters():
        if n.startswith("cls_fc."):
            assert p.grad is None and torch.equal(p.detach(), before[n]), n
    assert not torch.equal(model.state_dict()["features.0.0.weight"], before["features.0.0.weight"])
    r = tp.train_step(state, case["model"], {}, imgs, gt_kp, cats, torch.ones(case["batch"], 1024),
                      loss_cfg=dict(tp.DEFAULT_LOSS, names=["l1", "add_loss"], coeffs=([1.0, 0.1], [])), step_optimizer=False)
    assert abs(loss.item() - r["loss"]) < 1e-3 * abs(r["loss"])
    assert r["grads"]["cls_fc.1.weight"] is None and r["grads"]["cls_fc.1.bias"] is None
