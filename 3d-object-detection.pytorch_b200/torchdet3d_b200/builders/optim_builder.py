"""`build_optimizer(cfg, net)` (torchdet3d/builders/optim_builder.py:5-19): 'sgd' (momentum +
nesterov), 'adam' (which the reference builds as AdamW), 'rmsprop', 'adadelta', one global weight
decay over all parameters.

For the B200 Regressor this returns FusedOptimizer: a torch.optim.Optimizer (so LR schedulers,
`param_groups[0]['lr']`, `state_dict()` keep working) whose `step()` is ONE multi-tensor kernel
over the flat parameter arena.  Heads whose class was absent from the batch (grad=None in the
reference, model_builder.py:137) are skipped via the device-side `present` mask and keep their
own step counters."""
import ctypes as C

import torch

from .. import _lib as L
from ..models.regressor import Regressor, MAX_CLASSES

AVAILABLE_OPTIMS = ['sgd', 'rmsprop', 'adam', 'adadelta']
_KIND = {'sgd': L.OPT_SGD, 'adam': L.OPT_ADAMW, 'rmsprop': L.OPT_RMSPROP, 'adadelta': L.OPT_ADADELTA}


class FusedOptimizer(torch.optim.Optimizer):
    def __init__(self, model, name, lr, weight_decay=0.0, momentum=0.0, nesterov=False, betas=(0.9, 0.999),
                 alpha=0.99, rho=0.9):
        assert isinstance(model, Regressor)
        self.model = model
        self.kind = _KIND[name]
        eps = {'adam': 1e-8, 'rmsprop': 1e-8, 'adadelta': 1e-6, 'sgd': 0.0}[name]
        defaults = dict(lr=lr, weight_decay=weight_decay, momentum=momentum, nesterov=nesterov, betas=tuple(betas),
                        alpha=alpha, rho=rho, eps=eps)
        super().__init__(list(model.parameters()), defaults)
        self.grad_scale = 1.0
        self._alloc()

    def _alloc(self):
        flat = self.model._flat
        self.state0 = torch.zeros_like(flat)
        self.state1 = torch.zeros_like(flat) if self.kind in (L.OPT_ADAMW, L.OPT_ADADELTA) else None
        self.steps = torch.zeros(1 + MAX_CLASSES, dtype=torch.int32, device=flat.device)

    def desc(self):
        g = self.param_groups[0]
        return L.OptimDesc(self.kind, float(g['lr']), float(g['weight_decay']), float(g['momentum']),
                           int(bool(g['nesterov'])), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']),
                           float(g['alpha']), float(g['rho']), float(self.grad_scale))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        m = self.model
        if self.state0.device != m._flat.device:
            self._alloc()
        plan = m._last_plan
        if plan is None:
            raise L.Td3dError("FusedOptimizer.step() before any forward/backward")
        d = self.desc()
        L.check(L.lib().td3d_optim_step(plan.handle, C.byref(d), L.ptr(self.state0), L.ptr(self.state1), L.ptr(self.steps),
                                        L.ptr(m.present), L.stream()))
        m.mark_packed()
        return loss

    def zero_grad(self, set_to_none=True):
        """td3d_backward overwrites the gradient arena, so this only drops the `.grad` views."""
        for p in self.model._params:
            p.grad = None

    def state_dict(self):
        return {'state': {'state0': self.state0, 'state1': self.state1, 'steps': self.steps, 'kind': self.kind},
                'param_groups': [{k: v for k, v in g.items() if k != 'params'} for g in self.param_groups]}

    def load_state_dict(self, sd):
        st = sd['state']
        assert st['kind'] == self.kind, "optimizer kind mismatch"
        self.state0.copy_(st['state0'])
        if self.state1 is not None and st['state1'] is not None:
            self.state1.copy_(st['state1'])
        self.steps.copy_(st['steps'])
        for g, s in zip(self.param_groups, sd['param_groups']):
            g.update(s)


def build_optimizer(cfg, net):
    assert cfg.optim.name in AVAILABLE_OPTIMS
    o = cfg.optim
    target = net.module if hasattr(net, "module") else net
    if isinstance(target, Regressor):
        return FusedOptimizer(target, o.name, lr=o.lr, weight_decay=o.wd, momentum=o.momentum if o.name == 'sgd' else 0.0,
                              nesterov=bool(o.nesterov) if o.name == 'sgd' else False,
                              betas=o.betas if o.name == 'adam' else (0.9, 0.999), alpha=o.alpha, rho=o.rho)
    # any other nn.Module: plain torch optimizers, exactly the reference mapping
    if o.name == 'adadelta':
        return torch.optim.Adadelta(net.parameters(), lr=o.lr, rho=o.rho, weight_decay=o.wd)
    if o.name == 'adam':
        return torch.optim.AdamW(net.parameters(), lr=o.lr, betas=o.betas, weight_decay=o.wd)
    if o.name == 'rmsprop':
        return torch.optim.RMSprop(net.parameters(), lr=o.lr, weight_decay=o.wd, alpha=o.alpha)
    return torch.optim.SGD(net.parameters(), lr=o.lr, weight_decay=o.wd, momentum=o.momentum, nesterov=o.nesterov)
