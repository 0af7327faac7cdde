"""Standalone tcgen05 GEMM check (run in a subprocess by test_gpu_tc.py so that a device-side
trap cannot poison the pytest process).  Prints one JSON object."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-object-detection.pytorch_b200"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch  # noqa: E402

from torchdet3d_b200 import _lib as L  # noqa: E402
import _k as K  # noqa: E402


def rel_err(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


NT = [(128, 64, 64), (300, 64, 16), (1000, 24, 72), (257, 88, 24), (129, 960, 160), (64, 1280, 960), (50000, 16, 64),
      (130, 200, 80), (128, 184, 80), (77, 40, 120), (20000, 72, 24), (4096, 240, 40), (3000, 576, 96), (100, 16, 16),
      (100, 32, 32), (640, 120, 40),
      # long-M layers with N <= 64 run 256-row tiles (two MMA row blocks per tile): ragged last tile whose second row block is
      # partly / entirely beyond M, one and several k blocks
      (120001, 16, 16), (115000, 64, 16), (130050, 24, 64), (114177, 40, 120), (113600, 64, 200)]
TN = [(256, 64, 64), (1000, 64, 16), (5000, 24, 72), (257, 88, 24), (4097, 960, 160), (64, 1280, 960), (50000, 16, 64),
      (130, 200, 80), (20000, 72, 24), (3000, 96, 576), (640, 40, 120)]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.manual_seed(0)
    L.require_b200()
    out = {"nt": {}, "tn": {}}
    dev = "cuda"
    if which in ("all", "nt"):
        for M, N, Kd in NT:
            a = torch.randn(M, Kd, device=dev).bfloat16()
            w = (torch.randn(N, Kd, device=dev) / Kd ** 0.5).bfloat16()
            add = torch.randn(M, N, device=dev).bfloat16()
            bias = torch.randn(N, device=dev)
            ys = torch.randn(M, N, device=dev).bfloat16()
            ref = a.float() @ w.float().t()
            y, st = K.gemm_nt(a, w, L.BF16, L.GEMM_TCGEN05, slots=5)
            torch.cuda.synchronize()
            e1 = rel_err(y.float(), ref)
            yf = y.float()
            es = max(rel_err(st.sum(0)[0], yf.sum(0)) if yf.sum(0).abs().max() > 1e-3 else 0.0,
                     rel_err(st.sum(0)[1], (yf * yf).sum(0)))
            y2, _ = K.gemm_nt(a, w, L.BF16, L.GEMM_TCGEN05, addend=add, bias=bias, slots=0, out_f32=True)
            torch.cuda.synchronize()
            e2 = rel_err(y2, ref + add.float() + bias)
            # data-gradient flavour: bf16 output + residual addend, sums of g and g * saved-y of the stored (rounded) values
            y3, st3 = K.gemm_nt(a, w, L.BF16, L.GEMM_TCGEN05, addend=add, ysaved=ys, slots=3)
            torch.cuda.synchronize()
            y3f = y3.float()
            es2 = max(rel_err(st3.sum(0)[1], (y3f * ys.float()).sum(0)), rel_err(st3.sum(0)[0], y3f.sum(0)) if y3f.sum(0).abs().max() > 1e-3 else 0.0,
                      rel_err(y3f, ref + add.float()) / 3.0)
            out["nt"][f"{M}x{N}x{Kd}"] = [e1, es, e2, es2]
    if which in ("all", "tn"):
        for M, N1, N2 in TN:
            a = torch.randn(M, N1, device=dev).bfloat16()
            b = torch.randn(M, N2, device=dev).bfloat16()
            ref = a.float().t() @ b.float()
            c = K.gemm_tn(a, b, L.BF16, L.GEMM_TCGEN05)
            torch.cuda.synchronize()
            out["tn"][f"{M}x{N1}x{N2}"] = [rel_err(c, ref)]
    print("TC_RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
