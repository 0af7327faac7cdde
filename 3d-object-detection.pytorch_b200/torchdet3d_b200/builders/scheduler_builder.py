"""`build_scheduler(cfg, optimizer)` (torchdet3d/builders/scheduler_builder.py:5-25). Schedulers
only move a host-side scalar (`param_groups[0]['lr']`) once per epoch, so the torch schedulers are
kept; the fused optimizer reads the scalar at every step."""
import torch

AVAILABLE_SCHEDS = ['cosine', 'exp', 'stepLR', 'multistepLR']


def build_scheduler(cfg, optimizer):
    name = cfg.scheduler.name
    if name is None:
        return None
    assert name in AVAILABLE_SCHEDS
    sched = torch.optim.lr_scheduler
    if name == 'cosine':
        return sched.CosineAnnealingLR(optimizer, T_max=cfg.data.max_epochs, eta_min=5e-6)
    if name == 'exp':
        return sched.ExponentialLR(optimizer, gamma=cfg.scheduler.exp_gamma)
    if name == 'stepLR':
        return sched.StepLR(optimizer, step_size=cfg.scheduler.steps[0], gamma=cfg.scheduler.gamma)
    return sched.MultiStepLR(optimizer, milestones=list(cfg.scheduler.steps), gamma=cfg.scheduler.gamma)
