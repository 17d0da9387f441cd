"""Data parallelism over the batch axis: one process per GPU, torch.distributed for plumbing.

Models (or parameter draws) are independent, so the batch is cut into contiguous shards
with no data-path exchange; the only collective is one all-reduce (NCCL over NVLink on
GPUs, gloo in the CPU tests) of the pooled -2 log L that an optimiser consumes
(SURVEY.md section 8e).  The reference has no counterpart: it evaluates one model per call
(metran/solver.py:62).
"""
import torch


def shard_bounds(total, rank, world):
    """Contiguous block [lo, hi) of ceil(total / world) models owned by `rank`."""
    per = -(-int(total) // int(world))
    lo = min(int(total), rank * per)
    hi = min(int(total), lo + per)
    return lo, hi


def pooled_objective(engine, mle_local, group=None, out=None):
    """sum over ALL ranks of the per-model -2 log L: deterministic on-device reduction of the
    local shard (mkf_reduce_sum) followed by one all-reduce of a single double."""
    pooled = engine.reduce_sum(mle_local, out=out)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(pooled, op=torch.distributed.ReduceOp.SUM, group=group)
    return pooled


def gather_per_model(mle_local, total, group=None):
    """Optional all-gather of the per-model vector (equal shard sizes are padded)."""
    dist = torch.distributed
    if not (dist.is_available() and dist.is_initialized()):
        return mle_local
    world = dist.get_world_size(group)
    per = -(-int(total) // world)
    buf = torch.full((per,), float("nan"), dtype=mle_local.dtype, device=mle_local.device)
    buf[: mle_local.numel()] = mle_local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat(parts)[:total]
