"""CPU emulation of the one-pass depthwise backward: tests/host/dwc_emul.cu compiles the __host__ side of the very
per-thread body the GPU runs (csrc/dwc_core.cuh) and checks it against a naive double-precision conv2d backward for
every template instance the launcher uses (k 3/5, stride 1/2, fp32/bf16, odd plane sizes, partial bands)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_dwc_backward_host_emulation(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "dwc_emul")
    subprocess.run([nvcc, "-O1", "-std=c++17", "-x", "cu", "--expt-relaxed-constexpr", "-gencode",
                    "arch=compute_100a,code=sm_100a", "-diag-suppress", "20011", "-o", exe,
                    os.path.join(HERE, "host", "dwc_emul.cu")], check=True, capture_output=True, timeout=900)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "DWC_EMUL OK" in p.stdout, p.stdout[-3000:]
