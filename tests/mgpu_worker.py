"""Worker of tests/test_multigpu_nccl.py: run under `python -m torch.distributed.run --nproc-per-node N` (one rank per
GPU, NCCL).  Checks, on real hardware, the N>1 path of the AL-MPC solver:

 1. a rex-quadrotor batch sharded through `sharding.run_sharded` (real `MPC.forward` + implicit backward on every
    rank) equals, bit for bit, the same shards solved one after the other on a single GPU (per-shard stop rule);
 2. with `MPC.global_stop = sharding.GlobalStop(...)` (all-reduce hook inside the solve) a sharded run realises the
    Newton-step counts of the single-GPU run on the CONCATENATED batch and its results equal that run's bit for bit,
    on a batch built so that the per-shard rule would stop the shards at different steps;
 3. `sharding.allreduce_gradients` over NCCL (the policy-gradient all-reduce of DEQ-MPC training, north_star).
Prints "MGPU_OK <json>" on rank 0 when everything holds.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from deq_mpc_corl_b200 import sharding  # noqa: E402
from deq_mpc_corl_b200.al_mpc import MPC  # noqa: E402
from deq_mpc_corl_b200.al_utils import QuadCost  # noqa: E402
from deq_mpc_corl_b200.dynamics import make as make_dyn  # noqa: E402
from deq_mpc_corl_b200.problems import MODEL_SPECS, make_inputs, tracking_cost  # noqa: E402


def solve(dev, dyn, spec, T, x0, x_ref, u_ref, Qd, q, gx, gu, u_lo, u_hi, global_stop=None, calls=2):
    """`calls` chained MPC.forward + backward on one device; returns the last call's tensors and all step counts."""
    B = x0.shape[0]
    ctrl = MPC(spec["nx"], spec["nu"], T, u_lower=u_lo, u_upper=u_hi, n_batch=B)
    ctrl.global_stop = global_stop
    ctrl.reinitialize(x_ref.to(dev))
    ctrl.x_init, ctrl.u_init = x_ref.to(dev), u_ref.to(dev)
    steps = []
    for _ in range(calls):
        ql = q.to(dev).requires_grad_(True)
        x, u, _ = ctrl(x0.to(dev), QuadCost(Qd.to(dev), ql, None), dyn, dyn)
        qg, = torch.autograd.grad((x, u), ql, (gx.to(dev), gu.to(dev)))
        steps.append(ctrl.newton_steps())
    return (x.detach(), u.detach(), qg, ctrl.lamda_prev, ctrl.last_xu), steps


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    model, T, B = "rex_quadrotor", 5, 1000 + 3  # ragged split
    spec = MODEL_SPECS[model]
    dyn = make_dyn(model, dt=spec["dt"])
    inp = make_inputs(model, T, B, 4711)
    # skew the batch: the first shard tracks a trajectory that is already consistent (its residual statistic drops
    # below the stop threshold early), the rest is the generic recipe -> per-shard stop counts differ between shards
    lo0, hi0 = sharding.shard_bounds(B, world, 0)
    x_ref = inp["x_ref"].clone()
    x_ref[lo0:hi0] = inp["x0"][lo0:hi0, None, :] * torch.tensor([1.0, 0.999, 0.998, 0.997, 0.996])[None, :, None]
    Qd, q, _ = tracking_cost(inp["Qd"], x_ref, inp["u_ref"])
    g = torch.Generator().manual_seed(1)
    gx, gu = torch.randn(B, T, spec["nx"], generator=g), torch.randn(B, T, spec["nu"], generator=g)
    batch = dict(x0=inp["x0"], x_ref=x_ref, u_ref=inp["u_ref"], Qd=Qd, q=q, gx=gx, gu=gu)
    report = {"world": world}

    # ---- 1. per-shard stop rule: sharded run == the shards solved one by one on a single GPU
    steps_local = {}

    def per_gpu(**sh):
        out, st = solve(dev, dyn, spec, T, u_lo=inp["u_lower"], u_hi=inp["u_upper"], **sh)
        steps_local["steps"] = st
        return out

    got = sharding.run_sharded(per_gpu, batch)
    ok1 = True
    ref_steps = []
    for r in range(world):
        lo, hi = sharding.shard_bounds(B, world, r)
        ref, st = solve(dev, dyn, spec, T, u_lo=inp["u_lower"], u_hi=inp["u_upper"], **{k: v[lo:hi] for k, v in batch.items()})
        ref_steps.append(st)
        ok1 = ok1 and all(torch.equal(a[lo:hi], b) for a, b in zip(got, ref))
        if r == rank:
            ok1 = ok1 and st == steps_local["steps"]
    report["per_shard_steps"] = ref_steps

    # ---- 2. batch-global stop rule across the shards == single GPU on the whole batch
    full, full_steps = solve(dev, dyn, spec, T, u_lo=inp["u_lower"], u_hi=inp["u_upper"], **batch)
    gs = sharding.GlobalStop(dev)

    def per_gpu_global(**sh):
        out, st = solve(dev, dyn, spec, T, u_lo=inp["u_lower"], u_hi=inp["u_upper"], global_stop=gs, **sh)
        steps_local["global"] = st
        return out

    got2 = sharding.run_sharded(per_gpu_global, batch)
    ok2 = steps_local["global"] == full_steps and gs.error is None and gs.calls == 2 * 2 * 4
    ok2 = ok2 and all(torch.equal(a, b) for a, b in zip(got2, full))
    report["global_steps"], report["single_gpu_steps"] = steps_local["global"], full_steps
    report["global_differs_from_some_shard"] = any(st != full_steps for st in ref_steps)

    # ---- 3. policy-gradient all-reduce over NCCL
    net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.Linear(64, 16)).to(dev)
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    n_coll = sharding.allreduce_gradients(net.parameters(), bucket_bytes=2048)
    mean_rank = sum(range(1, world + 1)) / world
    ok3 = n_coll >= 2 and all(torch.allclose(p.grad, torch.full_like(p, mean_rank * (i + 1)))
                              for i, p in enumerate(net.parameters()))

    flags = torch.tensor([float(ok1), float(ok2), float(ok3)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    report.update(per_shard_equal=bool(flags[0]), global_stop_equal=bool(flags[1]), grad_allreduce=bool(flags[2]))
    if rank == 0:
        print(("MGPU_OK " if bool(flags.min()) else "MGPU_FAIL ") + json.dumps(report), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if bool(flags.min()) else 1)


if __name__ == "__main__":
    main()
