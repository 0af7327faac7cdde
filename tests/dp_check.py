"""Data-parallel parity on real GPUs (launched by tests/test_gpu_dp.py under torchrun, NCCL, world >= 2).

Each rank trains on its own shard with LOCAL BatchNorm statistics (the reference's DataParallel does not sync
BN, scripts/main.py:60-61); gradients are summed over NCCL inside the step and 1/world is folded into the
optimizer.  Checked per step against the CPU oracle (oracle/torch_port.py) run shard by shard:
  * the all-reduced gradient arena == sum over ranks of the oracle's per-shard gradients,
  * post-step weights == oracle optimizer step on the rank-averaged gradient,
  * a regressor head is skipped only if its class is absent on EVERY rank (classes 6 and 7 are present on rank 1 only,
    class 8 nowhere),
  * replicas stay bit-identical,
for the eager launch sequence and for the CUDA-graph capture (NCCL all-reduce captured on a side stream).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "3d-object-detection.pytorch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import torch_port as tp                                   # noqa: E402
from torchdet3d_b200 import _lib as L                                 # noqa: E402
from torchdet3d_b200.builders import build_model, build_loss, build_optimizer   # noqa: E402
from torchdet3d_b200.losses import LossManager                       # noqa: E402
from torchdet3d_b200.parallel import GradAllReduce                   # noqa: E402
from torchdet3d_b200.trainer import FusedTrainStep                   # noqa: E402
from torchdet3d_b200.utils import Dict                                # noqa: E402

MODEL, BS, RES, STEPS = "mobilenetv3_small", 6, 64, 5
OPT = dict(tp.DEFAULT_OPTIM, name="sgd", lr=0.01)
# Tolerances against the oracle: step 0 is a pure fwd+bwd+all-reduce+step comparison (tight).  From step 1 on the
# inputs of the comparison already differ by the fp32 rounding of step 0, which batch-statistic BatchNorm over
# 6 x (2x2) values amplifies (same factors as tests/test_gpu_model.py::_train_steps: x25 on gradients).
# Measured on 2 x B200: 2.7e-2 at step 1, 1.1e-1 at step 3 (chaotic growth of rounding differences; both ranks see the
# same number because they hold the same all-reduced arena).  Steps >= 2 are therefore checked against the EAGER run of
# the same kernels (below) and for bit-identical replicas, not against the oracle.
TOL_GRAD = [2e-3, 5e-2] + [None] * (STEPS - 2)
TOL_PARAM = [2e-3, 2e-2] + [None] * (STEPS - 2)


def shard(rank, step):
    imgs, gt_kp, cats, keep = tp.synth_batch(BS, res=RES, seed=500 + 10 * step + rank)
    # classes 4, 5 only on even ranks, 6, 7 only on odd ranks, class 8 nowhere
    cats = torch.tensor([0, 1, 2, 3, 4, 5] if rank % 2 == 0 else [7, 6, 0, 1, 2, 3])
    return imgs, gt_kp, cats, keep[:, :tp.block_table(MODEL)["head"]].contiguous()


def oracle_run(world):
    """Per step: summed gradients, post-step parameters, per-rank BN buffers, per-rank loss."""
    state = tp.synth_state(MODEL, seed=0)
    train_keys = tp.trainable_keys(state)
    bn_keys = [k for k in state if k not in train_keys]
    rank_bn = [{k: state[k].clone() for k in bn_keys} for _ in range(world)]
    opt_state, out = {}, []
    for step in range(STEPS):
        per = []
        for r in range(world):
            st = {k: v.clone() for k, v in state.items()}
            st.update({k: v.clone() for k, v in rank_bn[r].items()})
            res = tp.train_step(st, MODEL, {}, *shard(r, step), step_optimizer=False)
            rank_bn[r] = {k: st[k].clone() for k in bn_keys}
            per.append(res)
        gsum = {}
        for k in train_keys:
            gs = [p["grads"][k] for p in per if p["grads"][k] is not None]
            gsum[k] = None if not gs else sum(gs)
        avg = {k: (None if g is None else g / world) for k, g in gsum.items()}
        tp.optim_step(state, avg, opt_state, OPT)
        out.append(dict(gsum=gsum, params={k: state[k].clone() for k in train_keys},
                        bn=[dict(b) for b in rank_bn], loss=[p["loss"] for p in per]))
    return out


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    L.require_b200()
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    ref = oracle_run(world)
    traces = {}
    for use_graph in (False, True):
        cfg = Dict(model=dict(name=MODEL, pretrained=False, num_classes=9), optim=dict(OPT),
                   loss=dict(tp.DEFAULT_LOSS, alwa=dict(use=False, lam_cls=1., lam_reg=1., C=100, compute_std=True)),
                   b200=dict(dtype="fp32", gemm="auto"))
        cfg.loss.coeffs = (list(tp.DEFAULT_LOSS["coeffs"][0]), list(tp.DEFAULT_LOSS["coeffs"][1]))
        model = build_model(cfg)
        state0 = tp.synth_state(MODEL, seed=0)
        if rank != 0:                      # GradAllReduce must broadcast rank 0's replica
            state0 = {k: (v + 1 if v.dtype.is_floating_point else v) for k, v in state0.items()}
        model.load_state_dict(state0)
        model = model.to(dev).train()
        lm = LossManager(build_loss(cfg), cfg.loss.coeffs, cfg.loss.alwa)
        opt = build_optimizer(cfg, model)
        ar = GradAllReduce(model, opt)
        step = FusedTrainStep(model, lm, opt, BS, RES, RES, use_graph=use_graph, allreduce=ar)
        names = [n for n, _ in model.named_parameters()]
        table = {n: (off, numel) for n, off, numel, _ in model._param_table}
        trace = []
        for it in range(STEPS):
            imgs, gt_kp, cats, keep = shard(rank, it)
            step(imgs.to(dev), gt_kp.to(dev), cats.to(dev), keep.to(dev))
            torch.cuda.synchronize(dev)
            r = ref[it]
            assert abs(step.loss_terms[0].item() - r["loss"][rank]) < (2e-3 if it == 0 else 5e-2) * abs(r["loss"][rank]), (it, "loss")
            present = model.present.tolist()
            assert present == [1] * 8 + [0], (it, present)   # 4..7 live on one rank only -> still stepped; class 8: nowhere
            g = model._gflat.cpu().numpy()
            num = den = 0.0
            for n in names:
                off, numel = table[n]
                gr = r["gsum"][n]
                if gr is None:
                    continue
                d = g[off:off + numel] - gr.numpy().reshape(-1)
                num += float((d.astype(np.float64) ** 2).sum())
                den += float((gr.double() ** 2).sum())
            assert TOL_GRAD[it] is None or (num / den) ** 0.5 < TOL_GRAD[it], (it, "grad arena", (num / den) ** 0.5)
            trace.append((model._gflat.clone(), model._flat.clone()))
            sd = model.state_dict()
            worst = max((rel(sd[n].cpu().numpy(), r["params"][n].numpy()), n) for n in names)
            assert TOL_PARAM[it] is None or worst[0] < TOL_PARAM[it], (it, "params", worst)
            for k, v in r["bn"][rank].items():
                if k.endswith("num_batches_tracked"):
                    assert int(sd[k]) == int(v)
                elif it < 2:
                    np.testing.assert_allclose(sd[k].cpu().numpy(), v.numpy(), rtol=5e-3 if it == 0 else 5e-2, atol=5e-4 if it == 0 else 5e-3)
            flats = [torch.empty_like(model._flat) for _ in range(world)]
            dist.all_gather(flats, model._flat)
            for f in flats[1:]:
                assert torch.equal(f, flats[0]), (it, "replicas diverged")
        if use_graph:
            assert step._graph is not None, "graph was never captured"
            # graph replay (NCCL all-reduce captured on the side stream) against the eager launch sequence, step by step:
            # same kernels, same inputs -> only float-atomic ordering differs
            for it, ((g_e, p_e), (g_g, p_g)) in enumerate(zip(traces[False], trace)):
                dg = ((g_e - g_g).double().norm() / g_e.double().norm()).item()
                dp = ((p_e - p_g).double().norm() / p_e.double().norm()).item()
                # it = 2 is the first graph replay: the state it starts from differs from the eager run only by the
                # atomics noise of two steps; later steps amplify that noise chaotically (tiny batch-stat BatchNorms)
                tol_g = 0.1 if it <= 2 else 0.5        # measured 3.8e-2 at it = 2 on 8 x B200
                assert dg < tol_g and dp < 1e-2, (it, "graph vs eager", dg, dp)
        traces[use_graph] = trace
        if rank == 0:
            print(f"dp_check use_graph={use_graph}: OK ({STEPS} steps, world {world})", flush=True)
        # a CUDA graph that captured NCCL kernels must be gone before the communicator is torn down
        step._graph = None
        del step, ar
        torch.cuda.synchronize(dev)
    dist.barrier()
    torch.cuda.synchronize(dev)
    sys.stdout.flush()
    os._exit(0)       # skip communicator teardown (it can block after captured collectives); the process ends here


if __name__ == "__main__":
    main()
