"""`build_model(config, export_mode=False, weights_path='')` -- same config surface as the
reference factory (torchdet3d/builders/model_builder.py:25-71), returning the B200-native
Regressor instead of a torch.nn layer stack.

B200-specific knobs live under an optional `config.b200` key so reference configs run unchanged:
    b200 = dict(dtype='fp32'|'bf16', gemm='auto'|'simt'|'tcgen05')
"""
from ..models import Regressor, AVAILABLE_MODELS
from ..models.arch import REFERENCE_ONLY_MODELS
from ..utils import load_pretrained_weights
from .. import _lib as L

__AVAI_MODELS__ = set(AVAILABLE_MODELS) | set(REFERENCE_ONLY_MODELS)
_GEMM = {"auto": L.GEMM_AUTO, "simt": L.GEMM_SIMT, "tcgen05": L.GEMM_TCGEN05}


def build_model(config, export_mode=False, weights_path=''):
    name = config.model.name
    assert name in __AVAI_MODELS__, f"Wrong model name parameter. Expected one of {__AVAI_MODELS__}"
    if name in REFERENCE_ONLY_MODELS:
        raise NotImplementedError(
            f"'{name}' is defined by an un-vendored third-party package in the reference (timm / "
            "efficientnet_lite_pytorch, model_builder.py:4-8,62-69); the B200 path provides the in-repo "
            f"backbones {sorted(AVAILABLE_MODELS)}")
    b200 = config.get("b200") if hasattr(config, "get") else None
    b200 = b200 or {}
    model = Regressor(name, num_classes=config.model.num_classes, export_mode=export_mode,
                      compute_dtype=b200.get("dtype", "fp32"), gemm_impl=_GEMM[b200.get("gemm", "auto")])
    load = config.model.load_weights if "load_weights" in config.model else ''
    if load:
        load_pretrained_weights(model, load)
    elif weights_path:
        load_pretrained_weights(model, weights_path)
    elif config.model.pretrained and not export_mode:
        raise RuntimeError("pretrained ImageNet weights are downloaded by the reference (gdown, mobilenetv3.py:234-271); "
                           "no network here -- pass config.model.load_weights or weights_path instead")
    return model
