"""Export / inference forward at a large batch (BASELINE config 4: 4096 crops).  usage: export_bigb.py B [dtype] [gemm]
Compares a sub-sample with the CPU oracle (forward_export + select_by_argmax) and prints 'BIGB {json}'."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-object-detection.pytorch_b200")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from oracle import torch_port as tp  # noqa: E402
from torchdet3d_b200 import _lib as L  # noqa: E402
from torchdet3d_b200.builders import build_model  # noqa: E402
from torchdet3d_b200.utils import Dict  # noqa: E402


def main():
    B = int(sys.argv[1])
    dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    gemm = sys.argv[3] if len(sys.argv) > 3 else "auto"
    name = "mobilenetv3_large"
    L.require_b200()
    cfg = Dict(model=dict(name=name, pretrained=False, num_classes=9), b200=dict(dtype=dtype, gemm=gemm))
    m = build_model(cfg)
    state = tp.synth_state(name, seed=0)
    m.load_state_dict(state)
    m = m.cuda().eval()
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 3, 224, 224, generator=g)
    kp_sel, labels, logits = m.forward_to_onnx(x.cuda(), select=True)
    torch.cuda.synchronize()
    idx = torch.linspace(0, B - 1, 24).long()
    kp_all_ref, logits_ref = tp.forward_export(state, name, x[idx])
    ref_sel, ref_lab = tp.select_by_argmax(kp_all_ref, logits_ref)
    err_kp = (kp_sel[idx.cuda()].cpu() - ref_sel).abs().max().item()
    err_lg = ((logits[idx.cuda()].cpu() - logits_ref).abs().max() / logits_ref.abs().max()).item()
    agree = (labels[idx.cuda()].cpu() == ref_lab).float().mean().item()
    print("BIGB " + json.dumps(dict(B=B, dtype=dtype, gemm=gemm, kp_abs_err=err_kp, logits_rel_err=err_lg, label_agree=agree)))


if __name__ == "__main__":
    main()
