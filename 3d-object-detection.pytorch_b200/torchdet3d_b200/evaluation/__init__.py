from .metrics import (compute_average_distance, compute_accuracy, compute_metrics_per_cls, compute_2d_based_iou,
                      MetricAccumulator, set_iou_backend)
from .evaluate import Evaluator

__all__ = ["compute_average_distance", "compute_accuracy", "compute_metrics_per_cls", "compute_2d_based_iou", "MetricAccumulator",
           "set_iou_backend", "Evaluator"]
