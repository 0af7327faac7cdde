"""CPU: the C-ABI shared library loads and exports every symbol include/td3d.h declares; the
plan (host-only code) reproduces the reference parameter table; compute entry points fail loudly
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from oracle import torch_port as tp
from torchdet3d_b200 import _lib as L
from torchdet3d_b200.builders import build_model, build_loss, build_optimizer, build_scheduler, AVAILABLE_LOSS, AVAILABLE_OPTIMS, AVAILABLE_SCHEDS
from torchdet3d_b200.losses import LossManager
from torchdet3d_b200.models import Regressor
from torchdet3d_b200.utils import Dict, read_py_config, AverageMeter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "td3d.h")).read()
    declared = set(re.findall(r"\b(td3d_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"td3d_plan"}
    lib = L.lib()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert set(L.SYMBOLS) == declared, (set(L.SYMBOLS) ^ declared)
    assert lib.td3d_abi_version() == 2


@pytest.mark.parametrize("name", ["mobilenetv3_small", "mobilenetv3_large"])
def test_plan_param_table_matches_reference_state_dict(name):
    m = Regressor(name, num_classes=9)
    ref = tp.param_shapes(name, 9)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(tuple(sd[k].shape) == tuple(ref[k]) for k in ref)
    n_ref = {"mobilenetv3_small": 1695179, "mobilenetv3_large": 4423643}[name]
    assert sum(p.numel() for p in m.parameters()) == n_ref
    # load/save round trip through the flat arena
    st = tp.synth_state(name, seed=1)
    m.load_state_dict(st)
    assert all(torch.equal(m.state_dict()[k], st[k]) for k in st)
    off = m._param_table[3][1]
    assert torch.equal(m._flat[off:off + m._param_table[3][2]].view(m._param_table[3][3]), st[m._param_table[3][0]])


@pytest.mark.parametrize("name,n_params", [("efficientnet_b0", 4226599), ("efficientnet_b3", 10959059)])
def test_plan_param_table_matches_torchvision_efficientnet_through_the_wrapper(name, n_params):
    """BASELINE configs 3 / 5: keys, order and shapes of the torchvision `.features` state_dict + the reference wrapper's
    heads (oracle/effnet_port.py)."""
    from oracle import effnet_port as ep
    m = Regressor(name, num_classes=9)
    ref = ep.synth_state(name, seed=1)
    sd = m.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in ref)
    assert sum(p.numel() for p in m.parameters()) == n_params
    m.load_state_dict(ref)
    assert all(torch.equal(m.state_dict()[k], ref[k]) for k in ref)


def test_initialize_weights_distributions_match_reference():
    """`_initialize_weights` (mobilenetv3.py:205-218): conv N(0, sqrt(2 / (k*k*Cout))), BatchNorm 1 / 0, Linear N(0, 0.01)
    with zero bias -- run BEFORE the heads exist, so `regressors` / `cls_fc` keep torch's default Linear init
    (uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)), model_builder.py:76-87); BatchNorm buffers 0 / 1 / 0."""
    import math
    torch.manual_seed(0)
    m = Regressor("mobilenetv3_large", num_classes=9)
    sd = m.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean"):
            assert float(v.abs().max()) == 0.0, k
        elif k.endswith("running_var"):
            assert torch.all(v == 1), k
        elif k.endswith("num_batches_tracked"):
            assert int(v) == 0, k
        elif k.startswith("regressors") or k.startswith("cls_fc"):
            bound = 1.0 / math.sqrt(1280)
            assert float(v.abs().max()) <= bound + 1e-7, k
            if v.numel() > 1000:
                assert abs(float(v.std()) - bound / math.sqrt(3)) < 0.1 * bound, k        # uniform, not normal
        elif v.dim() == 4:
            std = math.sqrt(2.0 / (v.shape[2] * v.shape[3] * v.shape[0]))
            if v.numel() >= 2000:
                assert abs(float(v.std()) - std) < 0.12 * std and abs(float(v.mean())) < 0.1 * std, (k, float(v.std()), std)
        elif v.dim() == 2:
            assert abs(float(v.std()) - 0.01) < 0.002, k                                  # SE / classifier Linear
        elif k.endswith(".weight"):
            assert torch.all(v == 1), k                                                  # BatchNorm gamma
        else:
            assert float(v.abs().max()) == 0.0, k                                        # biases, BatchNorm beta


def test_no_cpu_fallback():
    m = Regressor("mobilenetv3_small")
    with pytest.raises(L.Td3dError):
        m(torch.rand(2, 3, 64, 64), torch.zeros(2, dtype=torch.int64))
    from torchdet3d_b200.losses import ADD_loss
    with pytest.raises(L.Td3dError):
        ADD_loss()(torch.rand(4, 9, 2), torch.rand(4, 9, 2))
    from torchdet3d_b200.evaluation import compute_average_distance
    with pytest.raises(L.Td3dError):
        compute_average_distance(torch.rand(4, 9, 2), torch.rand(4, 9, 2))


def test_plan_rejects_bad_descriptors():
    lib = L.lib()
    blocks = (L.BlockDesc * 1)(L.BlockDesc(3, 1, 16, 20, 16, 0, 0, 0, 0, 0))     # exp_ch not a multiple of 8
    net = L.NetDesc(0, 16, 1, blocks, 96, 128, 9, 9, 18)
    h = C.c_void_p()
    assert lib.td3d_plan_create(C.byref(net), 4, 64, 64, 0, 0, C.byref(h)) != 0
    assert b"multiples of 8" in lib.td3d_last_error()
    assert lib.td3d_plan_create(C.byref(net), 0, 64, 64, 0, 0, C.byref(h)) != 0


def test_builders_surface_like_reference_tests():
    """Mirror of the reference's tests/test_pipeline.py::test_builders (:32-48)."""
    cfg = read_py_config(os.path.join(ROOT, "tests", "configs", "default_config.py"))
    for loss_ in AVAILABLE_LOSS:
        if loss_ != 'cross_entropy':
            cfg['loss']['names'] = [loss_, 'cross_entropy']
            cfg.loss.coeffs = ([1.], [1.])
            reg, cls = build_loss(cfg)
            assert len(reg) == 1 and len(cls) == 1
            LossManager((reg, cls), cfg.loss.coeffs, cfg.loss.alwa)
    model = build_model(cfg)
    assert model is not None
    for optim_ in AVAILABLE_OPTIMS:
        cfg['optim']['name'] = optim_
        optimizer = build_optimizer(cfg, model)
        assert optimizer is not None
        for schd in AVAILABLE_SCHEDS:
            cfg['scheduler']['name'] = schd
            assert build_scheduler(cfg, optimizer) is not None
    assert not cfg.model.load_weights and not cfg.model.resume           # missing keys read as falsy
    with pytest.raises(AssertionError):
        cfg.model.name = "resnet50"
        build_model(cfg)
    with pytest.raises(NotImplementedError):
        cfg.model.name = "mobilenetv3_large_21k"
        build_model(cfg)


def test_average_meter_semantics():
    m = AverageMeter()
    m.update(2.0, 4); m.update(4.0, 12)
    assert m.val == 4.0 and m.count == 16 and abs(m.avg - 3.5) < 1e-12
