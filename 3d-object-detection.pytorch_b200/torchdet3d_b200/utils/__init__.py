from .utils import (Dict, OBJECTRON_CLASSES, AverageMeter, put_on_device, read_py_config, save_snap,
                    load_checkpoint, load_pretrained_weights, resume_from, set_random_seed)

__all__ = ["Dict", "OBJECTRON_CLASSES", "AverageMeter", "put_on_device", "read_py_config", "save_snap",
           "load_checkpoint", "load_pretrained_weights", "resume_from", "set_random_seed"]
