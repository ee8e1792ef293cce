"""Data-parallel host logic (device-agnostic, so it is testable with gloo on CPU).

Sharding follows the reference: rank p owns envs [p*B/P, (p+1)*B/P) with B/P a multiple of group_size
(rlinf/config.py:1109-1117); per-rank shuffle seed = actor.seed + rank
(workers/actor/embodied_fsdp_actor_worker.py:511-513); per-rank batch = global_batch_size // world (:523);
gradients are summed with ONE all-reduce on the flat buffer and averaged (FSDP/DDP semantics) by folding
1/world into the clip+AdamW kernel; metrics: AVG over ranks, explained-variance statistics SUM (:573-589).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_envs(total_envs: int, world_size: int, rank: int, group_size: int = 1):
    if total_envs % world_size != 0:
        raise ValueError(f"total_num_envs={total_envs} is not divisible by world_size={world_size}")
    per = total_envs // world_size
    if per % max(group_size, 1) != 0:
        raise ValueError(f"envs per rank {per} must be a multiple of group_size {group_size}")
    return rank * per, per


def shuffle_seed(actor_seed: int, rank: int) -> int:
    return int(actor_seed) + int(rank)


def per_rank_batch(global_batch_size: int, world_size: int, micro_batch_size: int, rollout_size: int):
    per = global_batch_size // world_size
    if rollout_size % per != 0:
        raise AssertionError(f"{rollout_size} is not divisible by {per}")
    if per % micro_batch_size != 0:
        raise AssertionError(f"train_global_batch_size={per}, {micro_batch_size}")
    return per, per // micro_batch_size, rollout_size // per


def allreduce_flat_grads(flat_grads: torch.Tensor, world_size: int, group=None) -> float:
    """SUM all-reduce in place; returns the grad_scale (1/world) the optimiser kernel must apply."""
    if world_size > 1:
        dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world_size


def reduce_metric_pack(mean_vec: torch.Tensor, ev_sum: torch.Tensor, grad_norm: torch.Tensor, world_size: int,
                       group=None) -> torch.Tensor:
    """[mean metrics | EV sufficient statistics | mean grad norm] -> AVG / SUM / AVG over ranks (one or two
    collectives on a few dozen floats)."""
    packed = torch.cat([mean_vec.float(), ev_sum.float(), grad_norm.reshape(1).float()])
    if world_size <= 1:
        return packed
    summed = packed.clone()
    dist.all_reduce(summed, op=dist.ReduceOp.SUM, group=group)
    n = mean_vec.numel()
    out = summed / world_size
    out[n: n + ev_sum.numel()] = summed[n: n + ev_sum.numel()]  # explained-variance statistics are SUMs
    return out


def broadcast_params(flat_params: torch.Tensor, world_size: int, src: int = 0, group=None):
    if world_size > 1:
        dist.broadcast(flat_params, src=src, group=group)
