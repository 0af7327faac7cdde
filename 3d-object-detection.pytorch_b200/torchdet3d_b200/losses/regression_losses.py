"""Keypoint / classification losses on the fused sm_100a loss kernel.

Drop-in surface of torchdet3d/losses/regression_losses.py (DiagLoss :8-20, ADD_loss :22-26,
WingLoss :28-49, LossManager :60-115) and of the torch criteria the reference's build_loss picks
(loss_builder.py:13-26).  Every criterion object is callable like the reference's
(`criterion(pred, target) -> scalar tensor`, differentiable), but `LossManager.parse_losses`
evaluates the WHOLE weighted sum, forward and backward, in ONE kernel launch
(td3d_loss_fwd_bwd) instead of ~15 micro-kernels per term.
"""
import ctypes as C

import torch

from .. import _lib as L

_FIELDS = {"l1": "w_l1", "smoothl1": "w_smoothl1", "mse": "w_mse", "add_loss": "w_add",
           "diag_loss": "w_diag", "wing": "w_wing", "cross_entropy": "w_ce"}
_TERM_INDEX = {"l1": 1, "smoothl1": 2, "mse": 3, "add_loss": 4, "diag_loss": 5, "wing": 6, "cross_entropy": 7}


def loss_desc_from(weights, smoothl1_beta=1.0, wing_w=0.05, wing_eps=2.0):
    """weights: {term name: coefficient}."""
    d = L.LossDesc()
    for k, v in weights.items():
        setattr(d, _FIELDS[k], float(v) + getattr(d, _FIELDS[k]))
    d.smoothl1_beta, d.wing_w, d.wing_eps = float(smoothl1_beta), float(wing_w), float(wing_eps)
    return d


class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, desc, pred_kp, gt_kp, logits, cats):
        if not pred_kp.is_cuda:
            raise L.Td3dError("td3d losses run on CUDA (sm_100a) only; there is no CPU fallback")
        B = pred_kp.shape[0]
        kp = pred_kp.detach().contiguous().float().view(B, -1)
        gt = gt_kp.detach().contiguous().float().view(B, -1)
        assert kp.shape[1] == 18, "the fused loss expects 9 keypoints x (x, y)"
        lg = logits.detach().contiguous().float() if logits is not None else None
        ct = cats.contiguous().to(torch.int64) if cats is not None else None
        nc = lg.shape[1] if lg is not None else 1
        out = torch.empty(8, device=kp.device)
        d_kp = torch.empty_like(kp)
        d_lg = torch.empty_like(lg) if lg is not None else None
        L.check(L.lib().td3d_loss_fwd_bwd(C.byref(desc), L.ptr(kp), L.ptr(gt), L.ptr(lg), L.ptr(ct), B, nc, L.ptr(out),
                                          L.ptr(d_kp), L.ptr(d_lg), L.stream()))
        ctx.save_for_backward(d_kp, d_lg if d_lg is not None else torch.empty(0, device=kp.device))
        ctx.kp_shape = pred_kp.shape
        ctx.has_logits = lg is not None
        total = out[0].clone()
        ctx.mark_non_differentiable(out)
        return total, out

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        d_kp, d_lg = ctx.saved_tensors
        gk = (d_kp * g_total).view(ctx.kp_shape)
        gl = d_lg * g_total if ctx.has_logits else None
        return None, gk, None, gl, None


def fused_loss(desc, pred_kp, gt_kp, logits=None, cats=None):
    """-> (total, terms[8]) ; terms = total, l1, smoothl1, mse, add, diag, wing, ce (weighted)."""
    total, terms = _FusedLossFn.apply(desc, pred_kp, gt_kp, logits, cats)
    return total, terms


class _Criterion:
    """One loss term; callable standalone exactly like the reference/torch criterion objects."""
    term = None
    is_class_loss = False

    def params(self):
        return {}

    def __call__(self, input_, target):
        desc = loss_desc_from({self.term: 1.0}, **self.params())
        if self.is_class_loss:
            B = input_.shape[0]
            dummy = torch.zeros(B, 18, device=input_.device)
            total, _ = _FusedLossFn.apply(desc, dummy, dummy, input_, target)
        else:
            total, _ = _FusedLossFn.apply(desc, input_, target, None, None)
        return total

    forward = __call__


class L1Loss(_Criterion):
    term = "l1"


class MSELoss(_Criterion):
    term = "mse"


class SmoothL1Loss(_Criterion):
    term = "smoothl1"

    def __init__(self, reduction="mean", beta=1.0):
        assert reduction == "mean"
        self.beta = beta

    def params(self):
        return dict(smoothl1_beta=self.beta)


class ADD_loss(_Criterion):
    """mean over instances of the summed per-keypoint L2 distance (regression_losses.py:22-26)."""
    term = "add_loss"


class DiagLoss(_Criterion):
    """SmoothL1(beta=.4) between predicted and target bbox diagonals (regression_losses.py:8-20,51-58)."""
    term = "diag_loss"


class WingLoss(_Criterion):
    """Wing loss with the reference's sequential masked updates (regression_losses.py:28-49)."""
    term = "wing"

    def __init__(self, w=0.05, eps=2):
        self.w, self.eps = w, eps

    def params(self):
        return dict(wing_w=self.w, wing_eps=self.eps)


class CrossEntropyLoss(_Criterion):
    term = "cross_entropy"
    is_class_loss = True


class LossManager:
    """Same constructor and `parse_losses` contract as the reference (regression_losses.py:60-115),
    including the optional ALWA re-weighting; the weighted sum runs as one fused kernel."""

    def __init__(self, criterions, coefficients, alwa):
        self.reg_criterions, self.class_criterions = criterions
        self.reg_coeffs, self.class_coeffs = coefficients
        assert len(self.reg_coeffs) == len(self.reg_criterions)
        assert len(self.class_coeffs) == len(self.class_criterions)
        assert self.reg_criterions
        self.use_alwa = bool(alwa.use) if hasattr(alwa, "use") else bool(alwa["use"])
        get = (lambda k: getattr(alwa, k)) if hasattr(alwa, "use") else (lambda k: alwa[k])
        if self.use_alwa:
            assert self.class_criterions
            assert self.reg_coeffs[0] == self.class_coeffs[0] == 1.
        self.lam_cls, self.lam_reg = get("lam_cls"), get("lam_reg")
        self.s_cls, self.s_reg = [], []
        self.C = get("C")
        self.alwa_version = 'ver_1' if get("compute_std") else 'ver_2'
        self.last_terms = None

    def _desc(self, lam_reg, lam_cls):
        weights, extra = {}, {}
        for k, cr in zip(self.reg_coeffs, self.reg_criterions):
            weights[cr.term] = weights.get(cr.term, 0.0) + float(k) * lam_reg
            extra.update(cr.params())
        for k, cr in zip(self.class_coeffs, self.class_criterions):
            weights[cr.term] = weights.get(cr.term, 0.0) + float(k) * lam_cls
        return loss_desc_from(weights, **extra)

    def loss_desc(self):
        lam_reg, lam_cls = (self.lam_reg, self.lam_cls) if self.use_alwa else (1.0, 1.0)
        return self._desc(lam_reg, lam_cls)

    def parse_losses(self, pred_kp, gt_kp, pred_cats, gt_cats, iter_):
        has_cls = bool(self.class_criterions)
        total, terms = fused_loss(self.loss_desc(), pred_kp, gt_kp, pred_cats if has_cls else None,
                                  gt_cats if has_cls else None)
        self.last_terms = terms
        if not self.use_alwa:
            return total
        # ALWA (regression_losses.py:96-113): running lists of lambda-weighted losses, re-weighted every C iters
        cls_w = terms[7].detach()
        reg_w = (terms[0] - terms[7]).detach()
        self.s_cls.append(cls_w)
        self.s_reg.append(reg_w)
        if iter_ % self.C == 0 and iter_ != 0:
            s_cls, s_reg = torch.stack(self.s_cls), torch.stack(self.s_reg)
            cls_mean, reg_mean = s_cls.mean(), s_reg.mean()
            if self.alwa_version == 'ver_1':
                cls, reg = cls_mean + s_cls.std(), reg_mean + s_reg.std()
            else:
                cls, reg = cls_mean, reg_mean
            self.s_cls.clear()
            self.s_reg.clear()
            if cls > reg:
                self.lam_cls = (1 - (cls - reg) / cls).item()
                print(f"classification coefficient changed : {self.lam_cls}")
                # the reference returns lam_reg*reg + lam_cls*cls with the NEW lam_cls (regression_losses.py:115):
                # re-evaluate the fused sum (forward + gradient) with the updated weight; one extra launch every C iterations
                total, terms = fused_loss(self.loss_desc(), pred_kp, gt_kp, pred_cats, gt_cats)
                self.last_terms = terms
        return total
