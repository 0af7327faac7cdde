"""2-rank NCCL data-parallel parity (tests/dp_check.py under torchrun); skipped with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_data_parallel_two_ranks_nccl_vs_oracle():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", "--tee", "3", os.path.join(HERE, "dp_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    err = "\n".join(l for l in p.stderr.splitlines() if "[default" in l and "frame #" not in l)
    assert p.returncode == 0, p.stdout[-2000:] + "\n" + err[-6000:]
    assert "use_graph=False: OK" in p.stdout and "use_graph=True: OK" in p.stdout
