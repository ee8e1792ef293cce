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


@register_advantage("grpo_video")
def compute_grpo_video_advantages(rewards: torch.Tensor, loss_mask: torch.Tensor, group_size: int, **kwargs):
    """Frame- or video-level group normalisation of step rewards [num_steps, B] (advantages.py:124-164)."""
    mode = kwargs.get("advantage_mode")
    if mode not in ("frame", "video"):
        raise ValueError(f"Unsupported grpo_video advantage_mode: {mode}")
    return ops.grpo_video_advantages(rewards, loss_mask, group_size, mode), None


@register_advantage("grpo_dynamic")
def compute_grpo_dynamic_advantages(rewards: torch.Tensor, loss_mask: torch.Tensor, group_size: int,
                                    idx_to_traj: list, advantage_mode: str = "turn", **kwargs):
    """Multi-turn GRPO per question (advantages.py:167-299). rewards [num_sequence, 1], loss_mask [seq_len, num_sequence]."""
    num_sequence = len(idx_to_traj)
    if rewards.numel() != num_sequence:
        raise AssertionError(f"Rewards size mismatch: {rewards.numel()} != {num_sequence}")
    num_trajectories = max(idx_to_traj) + 1
    if num_trajectories % group_size != 0:
        raise AssertionError(f"num_trajectories {num_trajectories} not divisible by group_size {group_size}")
    if advantage_mode not in ("trajectory", "turn"):
        raise ValueError(f"Invalid advantage_mode: {advantage_mode}. Must be 'trajectory' or 'turn'")
    turn_adv = ops.grpo_dynamic_turn_advantages(rewards, idx_to_traj, group_size, advantage_mode)
    adv, _ = ops.raw_advantages(turn_adv, loss_mask)  # broadcast over the sequence through the mask
    return adv, None


@register_advantage("reinpp")
def compute_reinpp_advantages(rewards: torch.Tensor, loss_mask: torch.Tensor, group_size: int,
                              use_reinpp_baseline: bool = False, kl_beta: float = 0.0, logprob=None, ref_logprob=None,
                              kl_penalty_type: str = "", **kwargs):
    """REINFORCE++ (advantages.py:302-364): reward at the (reference-defined) eos token, optional per-token KL penalty,
    reverse cumulative return, masked whitening with the biased variance. loss_mask [L, B]."""
    if use_reinpp_baseline:
        # the reference flattens the baselined rewards to 1-D and scatters them with a 2-D index (advantages.py:331-345):
        # torch raises, so the option is unusable in RLinf v0.4.0 as well - same error type here
        raise IndexError("Dimension out of range (expected to be in range of [-1, 0], but got 1)")
    ret, stats = ops.reinpp_returns(rewards, loss_mask, kl_beta, logprob, ref_logprob, kl_penalty_type or "k1")
    return ops.masked_normalize(ret, stats, mode=2, eps=1e-8), None


@register_advantage("opd")
def compute_opd_advantages(prev_logprobs: torch.Tensor, teacher_logprobs: torch.Tensor,
                           loss_mask: Optional[torch.Tensor] = None, normalize_advantages: bool = False, **kwargs):
    """Dense reverse-KL rewards teacher_logp - student_logp (advantages.py:367-407)."""
    assert teacher_logprobs is not None, "OPD advantage computation requires post-rollout teacher_logprobs."
    assert prev_logprobs is not None, "OPD advantage computation requires prev_logprobs from student rollout."
    assert teacher_logprobs.shape == prev_logprobs.shape, (
        f"teacher_logprobs shape {teacher_logprobs.shape} must match prev_logprobs shape {prev_logprobs.shape}.")
    assert not normalize_advantages, "VLA-OPD uses raw reverse-KL rewards; set normalize_advantages to False."
    num_action_chunks = kwargs.get("num_action_chunks", None)
    assert num_action_chunks is not None, "OPD advantage computation requires num_action_chunks."
    adv = ops.sub(teacher_logprobs, prev_logprobs)
    assert adv.shape[-1] % num_action_chunks == 0, (
        f"OPD token count {adv.shape[-1]} must be divisible by num_action_chunks {num_action_chunks}.")
    adv = adv.reshape(*adv.shape[:-1], num_action_chunks, -1)
    if loss_mask is not None:
        target_steps = loss_mask.shape[0]
        assert adv.shape[0] in {target_steps, target_steps + 1}, (
            f"OPD advantages time dimension {adv.shape[0]} must match loss_mask time dimension {target_steps} or "
            f"include one bootstrap step.")
        adv = adv[:target_steps]
    return adv, None


@register_advantage("raw")
def compute_raw_advantages(rewards: torch.Tensor, loss_mask: torch.Tensor, normalize_advantages: bool = False, **kwargs):
    """Scores broadcast over the sequence, optionally normalised over the valid entries (advantages.py:410-438)."""
    adv, stats = ops.raw_advantages(rewards, loss_mask, want_stats=bool(normalize_advantages))
    if normalize_advantages:
        ops.normalize_(adv, stats, 1e-5)  # unchanged when no entry is valid (the kernel checks the count)
    return adv, None
