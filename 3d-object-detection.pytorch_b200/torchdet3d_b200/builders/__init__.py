from .loss_builder import build_loss, AVAILABLE_LOSS
from .optim_builder import build_optimizer, AVAILABLE_OPTIMS, FusedOptimizer
from .scheduler_builder import build_scheduler, AVAILABLE_SCHEDS
from .model_builder import build_model

__all__ = ["build_loss", "AVAILABLE_LOSS", "build_optimizer", "AVAILABLE_OPTIMS", "FusedOptimizer",
           "build_scheduler", "AVAILABLE_SCHEDS", "build_model"]
