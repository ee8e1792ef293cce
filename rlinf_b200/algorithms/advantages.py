"""Registered advantage functions (CUDA-backed), signatures as rlinf/algorithms/advantages.py."""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from .registry import register_advantage


@register_advantage("gae")
def compute_gae_advantages_and_returns(rewards: torch.Tensor, gamma: float = 1.0, gae_lambda: float = 1.0,
                                       values: Optional[torch.Tensor] = None, normalize_advantages: bool = True,
                                       normalize_returns: bool = False, loss_mask: Optional[torch.Tensor] = None,
                                       dones: Optional[torch.Tensor] = None, **kwargs):
    """GAE over step-major [T,B] tensors (advantages.py:24-86): one scan kernel that also produces the
    normalisation statistics, then an in-place normalise pass for each normalised output.
    `normalize_advantages` defaults to True exactly as in the reference."""
    if rewards.dim() != 2:
        raise ValueError(f"rewards must be [seq_len, bsz], got {tuple(rewards.shape)}")
    T, B = rewards.shape
    if not rewards.is_contiguous() or (values is not None and not values.is_contiguous()):
        rewards = rewards.contiguous()
        values = values.contiguous() if values is not None else None
    dones = dones.contiguous()
    loss_mask = loss_mask.contiguous() if loss_mask is not None else None
    need_stats = bool(normalize_advantages or normalize_returns)
    adv, ret, stats = ops.gae(rewards, values, dones, gamma, gae_lambda, loss_mask, want_stats=need_stats)
    if normalize_advantages:
        ops.normalize_(adv, stats[0:3], 1e-5)
    if normalize_returns:
        ops.normalize_(ret, stats[3:6], 1e-5)
    return adv, ret


@register_advantage("grpo")
def compute_grpo_advantages(rewards: torch.Tensor, loss_mask: torch.Tensor, group_size: int, **kwargs):
    """Group-normalised scores broadcast over T through the loss mask (advantages.py:89-121).
    rewards: [num_groups, group_size] scores; loss_mask: [T, B]. Returns (advantages [T,B], None)."""
    scores = rewards.reshape(-1)
    T = loss_mask.shape[0]
    if loss_mask.shape[1] != scores.numel():
        raise RuntimeError(f"loss_mask {tuple(loss_mask.shape)} does not match {scores.numel()} scores")
    adv = ops.grpo_advantages(scores.contiguous(), loss_mask.contiguous(), T, group_size, 1e-6)
    return adv, None
