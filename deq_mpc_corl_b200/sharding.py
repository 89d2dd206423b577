"""Multi-GPU plumbing for the AL-MPC path: one process per GPU, the problem batch split contiguously.

The solve is embarrassingly parallel over problems, so the data path has NO
collective: every rank runs the fused solver on its own shard.  Collectives appear in exactly two
places, both here:

* ``allreduce_gradients`` — the policy-gradient all-reduce of DEQ-MPC training (the only trainable
  tensors are the network parameters, deqmpc/train.py:265); NCCL over NVLink on GPUs, gloo on CPU.
* ``gather_batch`` — optional re-assembly of per-shard outputs (evaluation / tests).

Stop-rule semantics under sharding: the reference's Newton stop test is a statistic of the WHOLE
batch (qpth/al_utils.py:486,552,560-564), and so is the rho_max exit of the streaming solve (qpth/AL_mpc.py:412).
Default: each rank applies the same rule to ITS shard ("per-shard stopping"); realised Newton-step counts are
reported per rank, and a sharded run is bit-identical to the reference run on each shard.
``GlobalStop`` (pass it as ``MPC.global_stop``) restores the reference's whole-batch decisions: after every line search
the library hands this rank's three statistics to a hook that all-reduces them over the group on the solver's stream
(two 8-byte SUMs and a MAX over NVLink: latency only, no host synchronisation), and a one-thread kernel takes the
decision identically on every rank — a sharded run then realises the Newton-step counts of the single-GPU run on the
concatenated batch.
"""
import ctypes

from typing import Callable, Dict, Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous [lo, hi) of rank's shard; the first n % world ranks get one extra problem."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], world, rank)
    return t[lo:hi]


def _world(group=None):
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_batch(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather shards of unequal length back into the [n_total, ...] batch (every rank gets it)."""
    world, rank = _world(group)
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    max_len = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((max_len,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)


def run_sharded(fn: Callable[..., Sequence[torch.Tensor]], batch_inputs: Dict[str, torch.Tensor], group=None,
                gather: bool = True):
    """Slice every [B, ...] input to this rank's shard, call ``fn(**shard)`` and (optionally) gather
    every output back to the full batch.  ``fn`` is the per-GPU solve (e.g. a closure over ``MPC``)."""
    world, rank = _world(group)
    n = next(iter(batch_inputs.values())).shape[0]
    local = {k: shard(v, world, rank) for k, v in batch_inputs.items()}
    outs = fn(**local)
    if not gather:
        return outs
    return tuple(gather_batch(o, n, group) for o in outs)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, average: bool = True,
                        bucket_bytes: int = 32 << 20) -> int:
    """Sum (or average) ``p.grad`` over the ranks with flat buckets sized for launch latency, not link
    count (NVSwitch: every peer at full bandwidth).  Returns the number of collectives issued."""
    world, _ = _world(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return 0
    n_coll = 0
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size, n_coll
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
        off = 0
        for g in bucket:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0
        n_coll += 1

    for g in grads:
        if bucket and (g.dtype != bucket[0].dtype or size + g.numel() * g.element_size() > bucket_bytes):
            flush()
        bucket.append(g)
        size += g.numel() * g.element_size()
    flush()
    return n_coll


class GlobalStop:
    """Batch-global stop rule across shards (include/almpc.h: ``almpc_desc.stop_stats`` / ``stop_reduce``).

    ``stats`` is the 4-double device buffer the solver's last line-search CTA fills with (sum ||c+||^2 at the entry
    point, at the exit point, max rho) of this rank's shard; ``hook`` is the C callback the library invokes at
    enqueue time, once per Newton-step slot.  The all-reduces are issued on the current stream (the one the
    solver launches on), so everything stays stream-ordered.  Every rank must call the solver with the same
    ``al_iter`` / ``max_newton`` (same number of collectives)."""

    def __init__(self, device, group=None):
        from . import _lib

        self.group = group
        self.stats = torch.zeros(4, dtype=torch.float64, device=device)
        self.calls = 0
        self.error = None

        def _reduce(stats_ptr, stream, user):
            try:
                if stats_ptr != self.stats.data_ptr():
                    raise RuntimeError("stop_reduce called with a foreign buffer")
                if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
                    dist.all_reduce(self.stats[:2], op=dist.ReduceOp.SUM, group=self.group)
                    dist.all_reduce(self.stats[2:3], op=dist.ReduceOp.MAX, group=self.group)
                self.calls += 1
                return 0
            except Exception as exc:  # never let an exception cross the C boundary
                self.error = exc
                return 1

        self.hook = _lib.STOP_REDUCE_FN(_reduce)  # keep a reference: the library stores the raw pointer only


def allreduce_stop_statistic(sum_sq: torch.Tensor, group=None) -> torch.Tensor:
    """Helper for callers that want the reference's batch-GLOBAL stop statistic across shards:
    sqrt(sum over ranks of per-shard sum of squares).  8 bytes, latency-only over NVLink."""
    world, _ = _world(group)
    t = sum_sq.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.sqrt()
