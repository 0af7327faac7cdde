from .train import Trainer
from .step import FusedTrainStep

__all__ = ["Trainer", "FusedTrainStep"]
