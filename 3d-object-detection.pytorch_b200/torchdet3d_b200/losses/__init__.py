from .regression_losses import (DiagLoss, ADD_loss, WingLoss, L1Loss, SmoothL1Loss, MSELoss, CrossEntropyLoss,
                                LossManager, fused_loss, loss_desc_from)

__all__ = ["DiagLoss", "ADD_loss", "WingLoss", "L1Loss", "SmoothL1Loss", "MSELoss", "CrossEntropyLoss",
           "LossManager", "fused_loss", "loss_desc_from"]
