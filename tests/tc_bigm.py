"""Large-M reproducer for the tcgen05 NT GEMM (BASELINE config 4 shapes: batch 1024..4096 crops, 112x112 pixels).
usage: tc_bigm.py M N K slots   -> prints 'BIGM {json}' (run in a subprocess: a device trap poisons the context)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-object-detection.pytorch_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from torchdet3d_b200 import _lib as L  # noqa: E402
import _k as K  # noqa: E402


def main():
    M, N, Kd, slots = (int(v) for v in sys.argv[1:5])
    torch.manual_seed(0)
    L.require_b200()
    dev = "cuda"
    a = torch.randn(M, Kd, device=dev).bfloat16()
    w = (torch.randn(N, Kd, device=dev) / Kd ** 0.5).bfloat16()
    y, st = K.gemm_nt(a, w, L.BF16, L.GEMM_TCGEN05, slots=slots)
    torch.cuda.synchronize()
    # sub-sampled check: first / last / strided row blocks against torch fp32 matmul of the same bf16 inputs
    err = 0.0
    for r0 in (0, M // 3, M - 4096):
        r0 = max(0, min(r0, M - 4096))
        ref = a[r0:r0 + 4096].float() @ w.float().t()
        err = max(err, ((y[r0:r0 + 4096].float() - ref).abs().max() / ref.abs().max()).item())
    es = None
    if slots:
        s1 = torch.zeros(N, device=dev, dtype=torch.float64)
        for r0 in range(0, M, 1 << 20):
            s1 += y[r0:r0 + (1 << 20)].double().sum(0)
        es = ((st.sum(0)[0].double() - s1).abs().max() / s1.abs().max().clamp_min(1e-9)).item()
    print("BIGM " + json.dumps(dict(M=M, N=N, K=Kd, slots=slots, err=err, stats_err=es)))


if __name__ == "__main__":
    main()
