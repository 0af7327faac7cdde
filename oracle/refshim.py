"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference package.

The reference (`/root/reference`, torchdet3d) is pure Python but imports ten third-party
modules that are not installed in this image (addict, albumentations, timm, efficientnet_lite*,
glog, icecream, prettytable ...).  None of them contributes arithmetic to the hot path
(SURVEY.md section 8c), so we register empty stand-ins in `sys.modules` and import the real
reference files from where they lie.  Nothing is copied and `/root/reference` is never written.

Used only by `oracle/make_golden.py` (fixture generation, this container) and by CPU tests that
are skipped when `/root/reference` is absent (it does not exist on the GPU box).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TD3D_REFERENCE_ROOT", "/root/reference")


class _AttrDict(dict):
    """Minimal stand-in for addict.Dict: attribute access, missing key -> empty (falsy) dict."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, cls):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return self[k]

    def __setattr__(self, k, v):
        self[k] = v

    def __missing__(self, k):
        v = type(self)()
        super().__setitem__(k, v)
        return v


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Anything:
    """Class usable as base class / callable placeholder for import-time symbol lookups."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "torchdet3d"))


def install():
    """Make `import torchdet3d` work against the reference tree. Idempotent."""
    if "torchdet3d" in sys.modules:
        return sys.modules["torchdet3d"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    _stub("addict", Dict=_AttrDict)
    alb_names = ["Compose", "Resize", "HorizontalFlip", "HueSaturationValue", "RGBShift",
                 "RandomBrightnessContrast", "ColorJitter", "Blur", "OneOf", "KeypointParams",
                 "Crop", "Normalize"]
    alb = _stub("albumentations", **{n: type(n, (_Anything,), {}) for n in alb_names})
    core = _stub("albumentations.core")
    ti = _stub("albumentations.core.transforms_interface",
               BasicTransform=type("BasicTransform", (_Anything,), {}),
               ImageOnlyTransform=type("ImageOnlyTransform", (_Anything,), {}),
               DualTransform=type("DualTransform", (_Anything,), {}),
               to_tuple=lambda *a, **k: tuple(a))
    aug = _stub("albumentations.augmentations")
    tr = _stub("albumentations.augmentations.transforms",
               Normalize=type("Normalize", (_Anything,), {}))
    alb.core, core.transforms_interface, alb.augmentations, aug.transforms = core, ti, aug, tr

    timm = _stub("timm")
    timm_models = _stub("timm.models")
    timm_mnv3 = _stub("timm.models.mobilenetv3", mobilenetv3_large_100=_Anything())
    timm.models, timm_models.mobilenetv3 = timm_models, timm_mnv3

    eff = _stub("efficientnet_lite_pytorch", EfficientNet=type("EfficientNet", (_Anything,), {}))
    eff.utils = _stub("efficientnet_lite_pytorch.utils", get_model_params=lambda *a, **k: (None, None))
    for i in range(3):
        cls = type(f"EfficientnetLite{i}ModelFile", (),
                   {"get_model_file_path": staticmethod(lambda: "")})
        _stub(f"efficientnet_lite{i}_pytorch_model", **{cls.__name__: cls})

    _stub("glog", info=print, warning=print, error=print)
    _stub("icecream", ic=lambda *a, **k: None)
    _stub("prettytable", PrettyTable=type("PrettyTable", (_Anything,), {
        "add_row": lambda self, *a, **k: None, "__str__": lambda self: "<table>"}))

    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "3rdparty", "Objectron")):
        if p not in sys.path:
            sys.path.append(p)
    import torchdet3d  # noqa: F401  (the real, unmodified reference package)
    return torchdet3d


def reference_config(model_name="mobilenetv3_small", num_classes=9):
    """Reference default config (configs/default_config.py) with an in-repo backbone selected."""
    install()
    from torchdet3d.utils import read_py_config
    cfg = read_py_config(os.path.join(REFERENCE_ROOT, "configs", "default_config.py"))
    cfg.model.name = model_name
    cfg.model.pretrained = False
    cfg.model.num_classes = num_classes
    return cfg
