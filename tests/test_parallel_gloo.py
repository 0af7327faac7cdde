"""CPU, world_size=2, gloo: host-side logic of the data-parallel path (bucket planning, in-place
chunked all-reduce of the flat gradient arena, OR of the head-present mask, 1/world folding)."""
import os
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torchdet3d_b200.parallel import plan_buckets, reduce_ready_chunks


def test_plan_buckets_cover_all_stages_once():
    for n in (1, 2, 13, 17):
        for k in (1, 2, 3, 4, 40):
            b = plan_buckets(n, k)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
            assert len(b) <= max(1, min(k, n))


def _worker(rank, world, path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    dist.init_process_group("gloo", init_method=f"file://{path}", rank=rank, world_size=world)
    n = 1000
    g = torch.arange(n, dtype=torch.float32) * (rank + 1)
    present = torch.tensor([1, 0, 0, 1] if rank == 0 else [0, 0, 1, 1], dtype=torch.int32)
    # chunks arrive in reverse parameter order, like the backward pass produces them
    chunks = [(700, 1000), (250, 700), (0, 250)]
    reduce_ready_chunks(g, chunks)
    dist.all_reduce(present, op=dist.ReduceOp.MAX)
    expect = torch.arange(n, dtype=torch.float32) * sum(r + 1 for r in range(world))
    assert torch.equal(g, expect)
    assert present.tolist() == [1, 0, 1, 1]
    # folding 1/world into the optimizer == averaging
    assert torch.allclose(g * (1.0 / world), torch.arange(n, dtype=torch.float32) * 1.5)
    dist.destroy_process_group()


def test_chunked_allreduce_world2_gloo():
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, os.path.join(d, "rdzv")), nprocs=2, join=True)
