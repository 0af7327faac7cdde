from .arch import MODEL_TABLES, AVAILABLE_MODELS, blocks_for, round_channels
from .regressor import Regressor, MAX_CLASSES, NUM_POINTS

# name the reference exports for its parameter tables (torchdet3d/models/mobilenetv3.py:20)
model_params = {k: dict(cfgs=[list(r) for r in v["rows"]], mode=k.rsplit("_", 1)[1]) for k, v in MODEL_TABLES.items()}

__all__ = ["Regressor", "MODEL_TABLES", "AVAILABLE_MODELS", "blocks_for", "round_channels", "model_params",
           "MAX_CLASSES", "NUM_POINTS"]
