"""CPU, world_size 2, gloo: the N>1 host logic (batch partition, gather, gradient all-reduce)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deq_mpc_corl_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        B = 11  # ragged: 6 + 5
        x = torch.randn(B, 5, 3, dtype=torch.float64)
        y = torch.randn(B, 4, dtype=torch.float64)

        def per_gpu_solve(x, y):  # stand-in for the per-shard MPC solve (independent problems)
            return (x * 2 + y[:, :3, None].transpose(1, 2).sum(-1, keepdim=True), y.cumsum(-1))

        full = per_gpu_solve(x, y)
        got = sharding.run_sharded(per_gpu_solve, {"x": x, "y": y})
        ok_gather = all(torch.equal(a, b) for a, b in zip(full, got))
        lo, hi = sharding.shard_bounds(B, world, rank)
        local_only = sharding.run_sharded(per_gpu_solve, {"x": x, "y": y}, gather=False)
        ok_local = torch.equal(local_only[0], full[0][lo:hi])

        # policy-gradient all-reduce: mixed dtypes, several buckets
        net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
        net[1].double()
        for i, p in enumerate(net.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        n_coll = sharding.allreduce_gradients(net.parameters(), bucket_bytes=64)
        exp = [(1 + 2) / 2 * (i + 1) for i in range(4)]
        ok_grad = all(torch.allclose(p.grad, torch.full_like(p, e)) for p, e in zip(net.parameters(), exp))
        stat = sharding.allreduce_stop_statistic(torch.tensor(float(rank + 1) ** 2, dtype=torch.float64))
        ok_stat = abs(stat.item() - (1 + 4) ** 0.5) < 1e-12
        q.put((rank, ok_gather, ok_local, ok_grad, n_coll, ok_stat))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_gather, ok_local, ok_grad, n_coll, ok_stat in res:
        assert ok_gather and ok_local and ok_grad and ok_stat, (rank, ok_gather, ok_local, ok_grad, ok_stat)
        assert n_coll >= 2


def test_shard_bounds_partition():
    for n in (0, 1, 7, 64, 65536 + 3):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)
