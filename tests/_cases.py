"""Shared helpers: golden-case definitions and oracle replay (test infrastructure)."""
import os

import numpy as np
import torch

from oracle import torch_port as tp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

LOSS_ALL = dict(names=["smoothl1", "l1", "mse", "wing", "add_loss", "diag_loss", "cross_entropy"],
                coeffs=([1.0, 0.5, 2.0, 0.3, 0.1, 0.7], [0.2]), smoothl1_beta=0.2, w=0.3, eps=0.5)

# must mirror oracle/make_golden.py::CASES
CASES = {
    "small_adamw": dict(model="mobilenetv3_small", batch=6, res=64, optim=dict(name="adam"),
                        loss=None, all_classes=False, steps=2),
    "small_sgd_allloss": dict(model="mobilenetv3_small", batch=12, res=64,
                              optim=dict(name="sgd", lr=0.05), loss=LOSS_ALL, all_classes=True, steps=2),
    "large_adamw": dict(model="mobilenetv3_large", batch=4, res=64, optim=dict(name="adam"),
                        loss=None, all_classes=False, steps=2),
    "small_224_adamw": dict(model="mobilenetv3_small", batch=3, res=224, optim=dict(name="adam"),
                            loss=None, all_classes=False, steps=1),
}


def load_golden(tag):
    return np.load(os.path.join(GOLDEN, f"{tag}.npz"), allow_pickle=False)


def optim_cfg(case):
    cfg = dict(tp.DEFAULT_OPTIM)
    cfg.update(case["optim"])
    return cfg


def loss_cfg(case):
    return tp.DEFAULT_LOSS if case["loss"] is None else case["loss"]


def unpack_mask(g, step, width):
    bits = np.unpackbits(g[f"s{step}_mask"], axis=1)[:, :width]
    return torch.tensor(bits.astype(np.float32))


def head_width(model):
    return tp.block_table(model)["head"]


def eval_batch(case):
    return tp.synth_batch(case["batch"], res=case["res"], seed=77, all_classes=case["all_classes"])


def train_batch(case, step):
    return tp.synth_batch(case["batch"], res=case["res"], seed=1000 + step, all_classes=case["all_classes"])
