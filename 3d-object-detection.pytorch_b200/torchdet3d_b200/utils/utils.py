"""Host-side helpers the hot path needs (subset of torchdet3d/utils/utils.py): config reading,
meters, device placement, checkpoint save / shape-matched load / resume.  Drawing, logging tee,
geometry, tracking and OpenVINO wrappers of the reference are outside the hot path."""
import os
import os.path as osp
import random
import sys
from collections import OrderedDict
from importlib import import_module

import numpy as np
import torch

OBJECTRON_CLASSES = ('bike', 'book', 'bottle', 'cereal_box', 'camera', 'chair', 'cup', 'laptop', 'shoe')  # utils.py:22


class Dict(dict):
    """Attribute dictionary with the behaviour the reference relies on from `addict.Dict`
    (utils.py:66-84): attribute + item access, nested dicts wrapped, a missing key reads as an
    empty (falsy) Dict -- e.g. `config.model.load_weights`, `cfg.model.resume`."""

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

    __setattr__ = __setitem__

    def __missing__(self, k):
        v = type(self)()
        super().__setitem__(k, v)
        return v


def set_random_seed(seed, deterministic=False):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


def read_py_config(filename):
    """Python-file config -> Dict (utils.py:66-84)."""
    filename = osp.abspath(osp.expanduser(filename))
    if not osp.isfile(filename):
        raise RuntimeError("config not found")
    assert filename.endswith('.py')
    module_name = osp.basename(filename)[:-3]
    if '.' in module_name:
        raise ValueError('Dots are not allowed in config file path.')
    sys.path.insert(0, osp.dirname(filename))
    try:
        sys.modules.pop(module_name, None)
        mod = import_module(module_name)
    finally:
        sys.path.pop(0)
    return Dict({k: v for k, v in mod.__dict__.items() if not k.startswith('__')})


def put_on_device(items, device):
    """utils.py:242-245 (non_blocking so pinned host batches overlap with compute)."""
    for i, item in enumerate(items):
        items[i] = item.to(device, non_blocking=True)
    return items


class AverageMeter:
    """utils.py:272-287."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def save_snap(model, optimizer, scheduler, epoch, log_path):
    """Checkpoint layout of the reference (utils.py:56-64): state_dict / optimizer / scheduler / epoch."""
    checkpoint = {'state_dict': model.state_dict(), 'optimizer': optimizer.state_dict(),
                  'scheduler': scheduler.state_dict() if scheduler is not None else None, 'epoch': epoch}
    snap_name = f'{log_path}/snap_{epoch}.pth'
    print(f'==> saving checkpoint to {snap_name}')
    torch.save(checkpoint, snap_name)
    return snap_name


def load_checkpoint(fpath):
    if fpath is None:
        raise ValueError('File path is None')
    if not osp.exists(fpath):
        raise FileNotFoundError('File is not found at "{}"'.format(fpath))
    return torch.load(fpath, map_location='cpu', weights_only=False)


def load_pretrained_weights(model, file_path='', pretrained_dict=None, extra_prefix=''):
    """Shape-matched partial load that strips 'module.' (utils.py:127-183): reference snapshots load
    into the B200 model and vice versa because the state_dict keys/shapes are identical."""
    checkpoint = load_checkpoint(file_path) if not pretrained_dict else pretrained_dict
    state_dict = checkpoint['state_dict'] if 'state_dict' in checkpoint else checkpoint
    model_dict = model.state_dict()
    new_state, matched, discarded = OrderedDict(), [], []
    for k, v in state_dict.items():
        if k.startswith('module.'):
            k = k[len('module.'):]
        k = extra_prefix + k
        if k in model_dict and model_dict[k].size() == v.size():
            new_state[k] = v
            matched.append(k)
        else:
            discarded.append(k)
    if not matched:
        raise RuntimeError(f'The pretrained weights {file_path or "pretrained dict"} cannot be loaded')
    model_dict.update(new_state)
    model.load_state_dict(model_dict)
    if discarded:
        print('** The following layers are discarded due to unmatched keys or layer size: {}'.format(discarded))
    return matched, discarded


def resume_from(model, chkpt_path, optimizer=None, scheduler=None):
    """utils.py:185-208."""
    checkpoint = load_checkpoint(chkpt_path)
    load_pretrained_weights(model, pretrained_dict=checkpoint['state_dict'] if 'state_dict' in checkpoint else checkpoint)
    if optimizer is not None and checkpoint.get('optimizer') is not None:
        optimizer.load_state_dict(checkpoint['optimizer'])
    if scheduler is not None and checkpoint.get('scheduler') is not None:
        scheduler.load_state_dict(checkpoint['scheduler'])
    return checkpoint['epoch'] + 1 if 'epoch' in checkpoint else 0
