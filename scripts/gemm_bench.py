"""Micro-benchmark of the tcgen05 NT GEMM on the hot streaming shapes (run on the B200 box), with and without the statistics
epilogue, plus the pipeline timeline of CTA 0 (TD3D_TC_DBG=32)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "3d-object-detection.pytorch_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torchdet3d_b200 import _lib as L
import _k as K
os.environ["TD3D_TC_LIVE_ENV"] = "1"
L.require_b200()
dev = "cuda"
SHAPES = [(256 * 112 * 112, 64, 16), (256 * 112 * 112, 16, 16), (256 * 56 * 56, 24, 64), (256 * 56 * 56, 72, 24), (256 * 14 * 14, 480, 80), (256 * 14 * 14, 112, 672)]
def run(M, N, Kd, slots, dbg):
    os.environ["TD3D_TC_DBG"] = str(dbg)
    a = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    st = torch.zeros(slots, 2, N, device=dev) if slots else None
    def call():
        L.check(L.lib().td3d_k_gemm_nt(L.ptr(a), L.ptr(w), L.ptr(y), None, None, None, L.ptr(st), slots, M, N, Kd, L.BF16, 0, L.GEMM_TCGEN05, L.stream()))
    for _ in range(3): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    gb = (M * Kd + M * N) * 2 / 1e9
    return us, gb / (us * 1e-6) / 1e3
for M, N, Kd in SHAPES:
    for slots, dbg in [(32, 0), (0, 0)]:
        us, tbs = run(M, N, Kd, slots, dbg)
        print(f"M={M} N={N} K={Kd} slots={slots} dbg={dbg:2d}: {us:8.1f} us  {tbs:6.2f} TB/s", flush=True)

# ---- pipeline timeline of CTA 0 (ns since first event) ----
import ctypes as C
import numpy as np
for M, N, Kd in [(256 * 112 * 112, 64, 16), (256 * 28 * 28, 240, 40), (256 * 14 * 14, 672, 112)]:
    for slots in (0, 32):
        os.environ["TD3D_TC_DBG"] = "32"
        a = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16()
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        st = torch.zeros(slots, 2, N, device=dev) if slots else None
        L.check(L.lib().td3d_k_gemm_nt(L.ptr(a), L.ptr(w), L.ptr(y), None, None, None, L.ptr(st), slots, M, N, Kd, L.BF16, 0, L.GEMM_TCGEN05, L.stream()))
        torch.cuda.synchronize()
        buf = (C.c_uint64 * 512)()
        L.check(L.lib().td3d_debug_tc_timeline(buf, 512))
        t = np.array(buf, dtype=np.int64).reshape(8, 64)
        t0 = t[t > 0].min()
        names = ["prod:empty_ok", "prod:tma_issued", "mma:tempty_ok", "mma:full_ok", "mma:issued", "epi:tfull_ok", "epi:chunks_done", "epi:flush_done"]
        print(f"--- timeline N={N} K={Kd} slots={slots} (ns, tiles 0..23 of CTA 0)")
        for e in range(8):
            print(f"{names[e]:16s}", " ".join(f"{int(v - t0):6d}" for v in t[e, :24]))
os.environ["TD3D_TC_DBG"] = "0"
