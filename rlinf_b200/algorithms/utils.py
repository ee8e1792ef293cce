"""Pre/post-processing around the advantage and loss kernels.

Mirror of rlinf/algorithms/utils.py: `preprocess_embodied_advantages_inputs` (:67-131),
`calculate_scores` (:134-152), `postprocess_embodied_advantages_outputs` (:155-174),
`preprocess_reasoning_advantages_inputs` (:177-260), `postprocess_reasoning_advantages_outputs`
(:263-277), `safe_normalize` (:397-404).  Layout changes are views (plus one contiguous copy where the
reference also copies); all arithmetic is done by librlinf_b200.so kernels.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib as L
from .. import ops


def preprocess_embodied_advantages_inputs(rewards, dones, values=None, loss_mask=None, loss_mask_sum=None,
                                          **kwargs) -> dict:
    """[nc,B,C] chunk-major -> [T=nc*C, B] step-major (utils.py:67-131). Tensors are moved to the
    device here (the reference's embodied batch lives on the host)."""
    dev = L.default_device() if not rewards.is_cuda else rewards.device
    rewards = L.to_device(rewards, dev, torch.float32)
    dones = L.to_device(dones, dev)
    values = L.to_device(values, dev, torch.float32)
    loss_mask = L.to_device(loss_mask, dev)
    loss_mask_sum = L.to_device(loss_mask_sum, dev) if (loss_mask_sum is not None and not _is_expanded(loss_mask_sum)) else (
        loss_mask_sum.to(dev) if loss_mask_sum is not None else None)
    if kwargs["reward_type"] == "chunk_level":
        rewards = rewards.sum(dim=-1, keepdim=True)
        dones = dones.max(dim=-1, keepdim=True)[0]
        if loss_mask is not None:
            loss_mask = loss_mask.max(dim=-1, keepdim=True)[0]
        if loss_mask_sum is not None:
            loss_mask_sum = loss_mask_sum.max(dim=-1, keepdim=True)[0]
    num_chunk, bsz, chunk_size = rewards.shape
    n_steps = num_chunk * chunk_size
    kwargs.update({"num_chunk": num_chunk, "batch_size": bsz, "chunk_size": chunk_size, "n_steps": n_steps})
    rewards = rewards.transpose(1, 2).reshape(n_steps, bsz)
    if loss_mask is not None:
        loss_mask = loss_mask.transpose(1, 2).reshape(n_steps, bsz)
    dones = dones.transpose(1, 2).reshape((num_chunk + 1) * chunk_size, bsz)[-(n_steps + 1):]
    if kwargs["adv_type"] == "gae":
        values = values.transpose(1, 2).reshape((num_chunk + 1) * chunk_size, bsz)[: n_steps + 1]
    kwargs.update({"rewards": rewards, "dones": dones, "values": values, "loss_mask": loss_mask,
                   "loss_mask_sum": loss_mask_sum})
    return kwargs


def _is_expanded(t: torch.Tensor) -> bool:
    return any(s == 0 for s in t.stride()) and t.numel() > 1


def calculate_scores(rewards, dones, **kwargs) -> dict:
    """Per-env return of the first episode, grouped (utils.py:134-152). The reference allocates the
    accumulator with torch.zeros(batch_size) (CPU only); here it is a kernel and runs on the device."""
    scores = ops.grpo_scores(rewards, dones).reshape(-1, kwargs["group_size"])
    kwargs.update({"rewards": scores, "dones": dones})
    return kwargs


def postprocess_embodied_advantages_outputs(advantages, num_chunk, chunk_size, returns=None, **kwargs) -> dict:
    res = {"advantages": advantages.reshape(num_chunk, chunk_size, -1).transpose(1, 2)}
    if returns is not None:
        res["returns"] = returns.reshape(num_chunk, chunk_size, -1).transpose(1, 2)
    return res


def preprocess_reasoning_advantages_inputs(rewards, loss_mask, values=None, logprob=None, ref_logprob=None,
                                           **kwargs) -> dict:
    """utils.py:177-260 for gae / grpo (rewards [bsz], loss_mask/values [bsz, L])."""
    dev = L.default_device() if not loss_mask.is_cuda else loss_mask.device
    rewards = L.to_device(rewards, dev, torch.float32)
    loss_mask = L.to_device(loss_mask, dev)
    bsz, seq_len = loss_mask.shape
    loss_mask = loss_mask.transpose(0, 1)
    assert rewards.ndim == 1, f"Unsupported reward shape {rewards.shape}"
    if kwargs["adv_type"] == "gae":
        expanded = torch.zeros((seq_len, bsz), dtype=rewards.dtype, device=dev)
        expanded[-1] = rewards
        kwargs.update({"rewards": expanded})
    elif kwargs["adv_type"] == "grpo":
        kwargs.update({"rewards": rewards.reshape(-1, kwargs["group_size"]).contiguous()})
    elif kwargs["adv_type"] == "grpo_dynamic":
        kwargs.update({"rewards": rewards.reshape(-1, kwargs["num_sequence"]).transpose(0, 1).contiguous()})
    elif kwargs["adv_type"] == "reinpp":
        kwargs.update({"rewards": rewards.unsqueeze(0)})
    elif kwargs["adv_type"] == "raw":
        kwargs.update({"rewards": rewards})
    else:
        assert False, f"Unsupported adv_type {kwargs['adv_type']}"
    if values is not None:
        assert values.ndim == 2, f"Unsupported values shape {values.shape}"
        values = L.to_device(values, dev, torch.float32).transpose(0, 1)
        values = torch.cat([values, torch.zeros((1, values.shape[-1]), dtype=values.dtype, device=dev)], dim=0)
        kwargs.update({"values": values})
    if logprob is not None:
        kwargs.update({"logprob": L.to_device(logprob, dev, torch.float32).transpose(0, 1)})
    if ref_logprob is not None:
        kwargs.update({"ref_logprob": L.to_device(ref_logprob, dev, torch.float32).transpose(0, 1)})
    dones = torch.zeros(seq_len + 1, bsz, dtype=torch.bool, device=dev)
    dones[-1] = True
    kwargs.update({"dones": dones, "loss_mask": loss_mask})
    return kwargs


def postprocess_reasoning_advantages_outputs(advantages, returns=None):
    advantages = advantages.transpose(0, 1).contiguous()
    if returns is not None:
        returns = returns.transpose(0, 1).contiguous()
    return advantages, returns


def safe_normalize(array: torch.Tensor, loss_mask: Optional[torch.Tensor], stats: Optional[torch.Tensor] = None):
    """(x - mean(valid)) / (std_unbiased(valid) + 1e-5); unchanged if no valid entry (utils.py:397-404).
    `stats` = {n,sum,sumsq} already produced by the scan kernel avoids a second reduction pass."""
    if stats is None:
        raise ValueError("safe_normalize needs the {n,sum,sumsq} statistics produced by the scan kernel")
    return ops.normalize_(array, stats, 1e-5)


class _KlPenalty(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logprob, ref_logprob, kind):
        out, g = ops.kl_penalty_raw(logprob.detach(), ref_logprob.detach(), kind, want_grad=logprob.requires_grad)
        ctx.save_for_backward(g)
        return out.view(logprob.shape)

    @staticmethod
    def backward(ctx, grad_out):
        (g,) = ctx.saved_tensors
        return (grad_out * g.view(grad_out.shape)) if g is not None else None, None, None


def kl_penalty(logprob: torch.Tensor, ref_logprob: torch.Tensor, kl_penalty: str) -> torch.Tensor:
    """KL estimators k1/abs/k2/k3 (rlinf/algorithms/utils.py:26-64), one elementwise kernel (forward value and
    d/dlogprob in the same pass); used as a loss term by the reasoning actor (fsdp_actor_worker.py:762-765)."""
    if kl_penalty == "full":
        raise NotImplementedError
    return _KlPenalty.apply(logprob, ref_logprob, kl_penalty)
