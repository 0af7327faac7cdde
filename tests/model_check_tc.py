"""bf16 + tcgen05 end-to-end check, run in a subprocess by test_gpu_tc_model.py."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-object-detection.pytorch_b200"), os.path.dirname(os.path.abspath(__file__))):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import torch_port as tp  # noqa: E402
from torchdet3d_b200.builders import build_loss, build_optimizer  # noqa: E402
from torchdet3d_b200.losses import LossManager  # noqa: E402
import test_gpu_model as T  # noqa: E402


def main():
    out = {}
    for name, B, res in (("mobilenetv3_small", 16, 96), ("mobilenetv3_large", 8, 128)):
        case = dict(model=name, optim=dict(name="sgd", lr=0.01), loss=None)
        imgs, gt_kp, cats, keep = tp.synth_batch(B, res=res, seed=99, all_classes=True)
        keep = keep[:, :tp.block_table(name)["head"]].contiguous()
        r = {}
        for gemm in ("simt", "tcgen05"):
            cfg, model = T.make_model(case, "bf16", gemm)
            lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
            model.train()
            kp, logits = model(imgs.cuda(), cats.cuda(), dropout_keep=keep.cuda())
            loss = lm.parse_losses(kp, gt_kp.cuda(), logits, cats.cuda(), 0)
            loss.backward()
            torch.cuda.synchronize()
            r[gemm] = (kp.detach().cpu().numpy(), loss.item(), model._gflat.clone().cpu().numpy())
        state = tp.synth_state(name, seed=0)
        o = tp.train_step(state, name, {}, imgs, gt_kp, cats, keep, step_optimizer=False)
        out[name] = dict(kp_tc_vs_simt=T.rel(r["tcgen05"][0], r["simt"][0]),
                         kp_tc_vs_oracle=T.rel(r["tcgen05"][0], o["kp"].numpy()),
                         loss_tc=r["tcgen05"][1], loss_simt=r["simt"][1], loss_oracle=o["loss"],
                         grad_tc_vs_simt=float(np.linalg.norm(r["tcgen05"][2] - r["simt"][2]) / np.linalg.norm(r["simt"][2])))
    print("MODEL_TC_RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
