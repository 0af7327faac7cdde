from .geometry import (get_default_camera_matrix, convert_camera_matrix_2_ndc, convert_2d_to_ndc, project_3d_points, lift_2d,
                       lift_2d_batch)
from .utils import (Dict, OBJECTRON_CLASSES, AverageMeter, put_on_device, read_py_config, save_snap,
                    load_checkpoint, load_pretrained_weights, resume_from, set_random_seed)

__all__ = ["Dict", "OBJECTRON_CLASSES", "AverageMeter", "put_on_device", "read_py_config", "save_snap",
           "load_checkpoint", "load_pretrained_weights", "resume_from", "set_random_seed", "get_default_camera_matrix",
           "convert_camera_matrix_2_ndc", "convert_2d_to_ndc", "project_3d_points", "lift_2d", "lift_2d_batch"]
