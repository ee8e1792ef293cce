"""Loss-scale stages of the dynamic (multi-turn, multi-agent) GRPO batches - mirror of rlinf/algorithms/loss_scales.py.

`group_level` (:21-51), `agent_level` (:54-109), `turn_level` (:112-182) and the registry of registry.py:127-141.
The grouping of turns by trajectory / agent is HOST logic on Python lists in the reference as well (a few dozen
turns per batch); the per-turn factors are computed here as plain Python floats (`*_factors`, testable without a
device) and applied to the device tensors with one multiply each.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable

import torch

LOSS_SCALE_REGISTRY: dict[str, Callable] = {}


def register_loss_scale(name: str):
    def bind(fn):
        LOSS_SCALE_REGISTRY[name.lower()] = fn
        return fn

    return bind


def get_loss_scales(names: list[str]) -> list[Callable]:
    out = []
    for name in names:
        if name not in LOSS_SCALE_REGISTRY:
            raise ValueError(f"Loss scale process {name} not registered")
        out.append(LOSS_SCALE_REGISTRY[name])
    return out


def _agents_of_each_trajectory(idx_to_traj, idx_to_sub_traj):
    """{trajectory: {agent: [turn indices]}} in first-appearance order."""
    tree: "OrderedDict[int, OrderedDict[int, list[int]]]" = OrderedDict()
    for turn, (traj, agent) in enumerate(zip(idx_to_traj, idx_to_sub_traj)):
        tree.setdefault(traj, OrderedDict()).setdefault(agent, []).append(turn)
    return tree


def group_factor(num_sequence: int, dp_world_size: int, actor_global_batch_size: int) -> float:
    return num_sequence * dp_world_size / actor_global_batch_size


def agent_factors(idx_to_traj, idx_to_sub_traj) -> list[float]:
    """1 / A_i / T_{i,a} for every turn (A_i agents in trajectory i, T_{i,a} turns of agent a)."""
    out = [1.0] * len(idx_to_traj)
    for agents in _agents_of_each_trajectory(idx_to_traj, idx_to_sub_traj).values():
        for turns in agents.values():
            for t in turns:
                out[t] = 1 / len(agents) / len(turns)
    return out


def turn_factors(idx_to_traj, idx_to_sub_traj, response_token_counts) -> list[float]:
    """T_{i,a} * |o_t| / sum_t |o_t|: turns the uniform per-turn weight of `agent_level` into a token-proportional one."""
    out = [1.0] * len(idx_to_traj)
    for agents in _agents_of_each_trajectory(idx_to_traj, idx_to_sub_traj).values():
        for turns in agents.values():
            total = sum(response_token_counts[t] for t in turns)
            for t in turns:
                out[t] = 1 * len(turns) * response_token_counts[t] / total
    return out


def _times(t: torch.Tensor, factors) -> None:
    t.mul_(torch.tensor(factors, dtype=t.dtype).to(t.device).reshape(-1, *([1] * (t.dim() - 1))))


@register_loss_scale("group_level")
def group_scale(context: dict, batch: dict) -> dict:
    done = context["folding_scale"]
    assert "group_level" not in done, (
        "`group_level` loss scaling can only be applied once. Apply the group-level factor before any `agent_level` "
        "or `turn_level` factor.")
    done.append("group_level")
    batch["advantages"] *= group_factor(len(batch["idx_to_traj"]), context.get("data_parallel_world_size", 1),
                                        context["actor_global_batch_size"])
    return batch


@register_loss_scale("agent_level")
def agent_scale(context: dict, batch: dict) -> dict:
    done = context["folding_scale"]
    assert "group_level" in done and "agent_level" not in done, (
        "`agent_level` loss scaling requires `group_level` to be applied first, and it can only be applied once.")
    done.append("agent_level")
    _times(batch["loss_scales"], agent_factors(batch["idx_to_traj"], batch["extra:idx_to_sub_traj"].tolist()))
    return batch


@register_loss_scale("turn_level")
def turn_scale(context: dict, batch: dict) -> dict:
    done = context["folding_scale"]
    assert "group_level" in done and "agent_level" in done and "turn_level" not in done, (
        "`turn_level` loss scaling requires both `group_level` and `agent_level` to be applied first, and it can only "
        "be applied once.")
    done.append("turn_level")
    counts = batch["response_mask"].sum(dim=tuple(range(1, batch["response_mask"].dim()))).tolist()  # one D2H
    _times(batch["loss_scales"], turn_factors(batch["idx_to_traj"], batch["extra:idx_to_sub_traj"].tolist(), counts))
    return batch
