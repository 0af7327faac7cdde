"""torchdet3d_b200 -- B200-native (sm_100a) drop-in for the hot path of torchdet3d
(sovrasov/3d-object-detection.pytorch): the second-stage 3D box regressor and its training step.

Same sub-package names as the reference (`builders`, `models`, `losses`, `trainer`,
`evaluation`, `utils`); everything outside the hot path (data loading, augmentation, geometry,
tracking, OpenVINO export/demo) is intentionally absent -- see DESIGN.md.
"""
from . import _lib, utils, models, losses, evaluation, builders, trainer, inference, preprocess  # noqa: F401
from .inference import InferSession  # noqa: F401

__version__ = "0.1.0"
