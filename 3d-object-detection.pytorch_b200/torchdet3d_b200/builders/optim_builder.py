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
        # plans hold a raw device pointer to `steps` (dropout counter): drop it before the tensor is replaced
        if getattr(self, "steps", None) is not None:
            for plan in self.model._plans.values():
                L.check(L.lib().td3d_plan_set_dropout_counter(plan.handle, None))
            self.model._dropout_counter_ref = None
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

    # ---- checkpoint format: the torch.optim layout the reference's snapshots use (utils/utils.py:56-64) --------------
    _STATE_KEYS = {L.OPT_ADAMW: ("exp_avg", "exp_avg_sq"), L.OPT_SGD: ("momentum_buffer", None),
                   L.OPT_RMSPROP: ("square_avg", None), L.OPT_ADADELTA: ("square_avg", "acc_delta")}

    def _head_of(self, pname):
        return int(pname.split(".")[1]) + 1 if pname.startswith("regressors.") else 0

    def state_dict(self):
        """Per-parameter `state` keyed by parameter index + `param_groups`, exactly what torch.optim.AdamW / SGD /
        RMSprop / Adadelta write, so a snapshot saved here resumes in the reference and vice versa.  A tensor that
        never received a gradient (absent-class head) has no entry, like torch's lazily created state."""
        steps = self.steps.tolist()
        k0, k1 = self._STATE_KEYS[self.kind]
        state = {}
        uses_buf = self.kind != L.OPT_SGD or self.param_groups[0]['momentum'] != 0
        for i, (pname, off, numel, shape) in enumerate(self.model._param_table):
            n = steps[self._head_of(pname)]
            if n == 0:
                continue
            ent = {}
            if self.kind != L.OPT_SGD:
                ent['step'] = torch.tensor(float(n))
            if uses_buf:
                ent[k0] = self.state0[off:off + numel].view(shape).clone()
            if k1 is not None:
                ent[k1] = self.state1[off:off + numel].view(shape).clone()
            if self.kind == L.OPT_SGD and not uses_buf:
                ent['momentum_buffer'] = None
            state[i] = ent
        groups = []
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != 'params'}
            d['params'] = list(range(len(self.model._param_table)))
            groups.append(d)
        return {'state': state, 'param_groups': groups, 'td3d_steps': steps}

    def load_state_dict(self, sd):
        st = sd['state']
        if 'state0' in st:                     # flat layout written by round-1 snapshots of this package
            assert st['kind'] == self.kind, "optimizer kind mismatch"
            self.state0.copy_(st['state0'])
            if self.state1 is not None and st['state1'] is not None:
                self.state1.copy_(st['state1'])
            self.steps.copy_(st['steps'])
        else:
            k0, k1 = self._STATE_KEYS[self.kind]
            table = self.model._param_table
            if any((not isinstance(i, int)) or i < 0 or i >= len(table) for i in st):
                raise RuntimeError("optimizer state_dict does not index this model's parameters (expected torch.optim "
                                   f"layout with indices 0..{len(table) - 1})")
            self.state0.zero_()
            if self.state1 is not None:
                self.state1.zero_()
            steps = [0] * (1 + MAX_CLASSES)
            for i, ent in st.items():
                pname, off, numel, shape = table[i]
                if tuple(ent[k0].shape if ent.get(k0) is not None else shape) != tuple(shape):
                    raise RuntimeError(f"optimizer state for '{pname}' has shape {tuple(ent[k0].shape)}, expected {shape}")
                if ent.get(k0) is not None:
                    self.state0[off:off + numel].copy_(ent[k0].reshape(-1))
                if k1 is not None:
                    if k1 not in ent:
                        raise RuntimeError(f"optimizer state_dict was written by a different optimizer (no '{k1}')")
                    self.state1[off:off + numel].copy_(ent[k1].reshape(-1))
                # SGD keeps no step count in torch: any saved buffer means "past the first step"
                n = int(ent['step']) if 'step' in ent else 1
                h = self._head_of(pname)
                steps[h] = max(steps[h], n)
            if 'td3d_steps' in sd:
                steps = list(sd['td3d_steps'])
            self.steps.copy_(torch.tensor(steps, dtype=torch.int32))
        for g, s in zip(self.param_groups, sd['param_groups']):
            g.update({k: v for k, v in s.items() if k != 'params'})


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
