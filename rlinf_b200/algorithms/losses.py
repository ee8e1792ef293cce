"""Registered policy losses (CUDA-backed), surface of rlinf/algorithms/losses.py.

`compute_ppo_actor_critic_loss` ("actor_critic", losses.py:396-424) and `compute_grpo_actor_loss_fn`
("actor", :508-535) take the kwargs produced by preprocess_loss_inputs (so they can be dropped into
the REFERENCE's own LOSS_REGISTRY); `fused_embodied_policy_loss` is what this package's `policy_loss`
uses for the embodied task type: it skips the torch-side reduction and hands the raw [bsz, C*A]
log-probs to the kernel.  Either way: one fused forward+backward kernel group, gradients delivered to
autograd through a Function whose backward only rescales the stored gradients.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import _lib as L
from .. import ops
from .registry import register_policy_loss

_HP_KEYS = ("clip_ratio_low", "clip_ratio_high", "clip_ratio_c", "clip_log_ratio_min", "clip_log_ratio_max",
            "value_clip", "huber_delta", "max_episode_steps", "critic_warmup")


class _FusedPpoLoss(torch.autograd.Function):
    """forward: launch the fused kernel (loss, metrics AND gradients); backward: grads x upstream scalar."""

    @staticmethod
    def forward(ctx, logprobs, values, entropy, cfg):
        loss, metrics, d_lp, d_v, d_e = ops.ppo_loss(
            logprobs=logprobs.detach(), values=None if values is None else values.detach(),
            entropy=None if entropy is None else entropy.detach(), want_grads=True, **cfg)
        ctx.grads = (d_lp, d_v, d_e)
        ctx.shapes = (logprobs.shape, None if values is None else values.shape,
                      None if entropy is None else entropy.shape)
        ctx.mark_non_differentiable(metrics)
        return loss.reshape(()), metrics

    @staticmethod
    def backward(ctx, g_loss, _g_metrics):
        d_lp, d_v, d_e = ctx.grads
        ctx.grads = None
        outs = []
        for g, shp in zip((d_lp, d_v, d_e), ctx.shapes):
            if g is None or shp is None:
                outs.append(None)
            else:
                outs.append(ops.scale_by_(g, g_loss).view(shp))
        return outs[0], outs[1], outs[2], None


def _check_fp32(**named):
    for k, v in named.items():
        if v is not None and v.dtype != torch.float32:
            raise AssertionError(f"{k} must be float32 to keep numerical stability")  # losses.py:232-240


def _metrics_dict(metrics: torch.Tensor, with_critic: bool, as_float: bool) -> dict:
    slots = L.ACTOR_SLOTS + (L.CRITIC_SLOTS if with_critic else ())
    if as_float:
        host = metrics.tolist()  # ONE device->host copy for all metrics
        return {L.M_KEYS[s]: host[s] for s in slots}
    return {L.M_KEYS[s]: metrics[s] for s in slots}


def postprocess_loss_metric(metrics_data: dict) -> dict:
    """rlinf/algorithms/utils.py:379-385."""
    for k, v in metrics_data.items():
        if isinstance(v, torch.Tensor):
            metrics_data[k] = v.detach().item()
    return metrics_data


def _run(logprobs, values, entropy, cfg, with_critic, as_float):
    needs_grad = torch.is_grad_enabled() and (
        logprobs.requires_grad or (values is not None and values.requires_grad))
    if needs_grad:
        loss, metrics = _FusedPpoLoss.apply(logprobs, values, entropy, cfg)
    else:
        loss, metrics, *_ = ops.ppo_loss(logprobs=logprobs.detach(), values=None if values is None else values.detach(),
                                         entropy=entropy, want_grads=False, **cfg)
        loss = loss.reshape(())
    return loss, _metrics_dict(metrics, with_critic, as_float), metrics


def fused_embodied_policy_loss(**kwargs):
    """policy_loss for task_type == "embodied" (registry.py:77-92 + utils.py:280-376 fused).
    kwargs are the ones EmbodiedFSDPActor.train_micro_batch builds
    (workers/actor/embodied_fsdp_actor_worker.py:642-676)."""
    loss_type = kwargs["loss_type"]
    logprobs = kwargs["logprobs"]
    dev = logprobs.device if logprobs.is_cuda else L.default_device()
    logprobs = logprobs if logprobs.is_cuda else logprobs.to(dev)
    with_critic = loss_type == "actor_critic"
    values = kwargs.get("values") if with_critic else None
    _check_fp32(logprobs=logprobs, old_logprobs=kwargs["old_logprobs"], advantages=kwargs["advantages"])
    A = int(kwargs.get("single_action_dim") or logprobs.shape[-1])
    bsz = logprobs.shape[0]
    total = logprobs[0].numel()
    if total % A != 0:
        raise RuntimeError(f"logprobs row of {total} entries is not a multiple of single_action_dim={A}")
    Cc = total // A
    logprob_type = kwargs.get("logprob_type")
    if logprob_type not in L.LOGPROB_TYPES:
        raise ValueError(f"unsupported logprob_type {logprob_type!r}")
    U = 1 if logprob_type == "chunk_level" else Cc

    def per_unit(t, name):
        if t is None:
            return None
        t = L.to_device(t, dev)
        if t.numel() != bsz * U:
            raise RuntimeError(f"{name} has {t.numel()} entries, expected bsz*{U}={bsz * U} "
                               f"(logprob_type={logprob_type}, reward_type={kwargs.get('reward_type')})")
        return t.reshape(bsz, U)

    lms = kwargs.get("loss_mask_sum")
    cfg = dict(
        old_logprobs=L.to_device(kwargs["old_logprobs"], dev).reshape(bsz, Cc * A),
        advantages=per_unit(kwargs["advantages"], "advantages"),
        returns=per_unit(kwargs.get("returns"), "returns") if with_critic else None,
        prev_values=per_unit(kwargs.get("prev_values"), "prev_values") if with_critic else None,
        loss_mask=per_unit(kwargs.get("loss_mask"), "loss_mask"),
        loss_mask_sum=per_unit(lms, "loss_mask_sum"),
        C_chunks=Cc, A_dim=A, logprob_type=logprob_type,
    )
    for k in _HP_KEYS:
        if kwargs.get(k) is not None:
            cfg[k] = kwargs[k]
    if with_critic and values is not None:
        values = values if values.is_cuda else values.to(dev)
        _check_v = per_unit(values.detach(), "values")  # shape check only
        del _check_v
    loss, metrics, _ = _run(logprobs.reshape(bsz, Cc * A), None if values is None else values.reshape(bsz, U), None,
                            cfg, with_critic, as_float=True)
    return loss, metrics


def _from_preprocessed(kwargs, with_critic):
    """Map the kwargs of preprocess_loss_inputs (already-reduced log-probs) onto kernel arguments."""
    logprobs = kwargs["logprobs"]
    old = kwargs["old_logprobs"]
    adv = kwargs["advantages"]
    _check_fp32(logprobs=logprobs, old_logprobs=old, advantages=adv)
    dev = logprobs.device if logprobs.is_cuda else L.default_device()
    if adv.shape == logprobs.shape or adv.numel() == logprobs.numel():
        g, mode = 1, "action_level"  # one ratio per entry
    elif adv.dim() == logprobs.dim() and adv.shape[-1] == 1:
        g, mode = logprobs.shape[-1], "token_level"
    else:
        raise RuntimeError(f"advantages {tuple(adv.shape)} cannot be matched to logprobs {tuple(logprobs.shape)}")
    n_units = logprobs.numel() // g

    def unit(t):
        if t is None:
            return None
        t = L.to_device(t, dev)
        if t.numel() != n_units:
            t = t.expand(adv.shape)
        return t.reshape(n_units, 1)

    cfg = dict(old_logprobs=L.to_device(old, dev).reshape(n_units, g), advantages=unit(adv),
               returns=unit(kwargs.get("returns")) if with_critic else None,
               prev_values=unit(kwargs.get("prev_values")) if with_critic else None,
               loss_mask=unit(kwargs.get("loss_mask")), loss_mask_sum=unit(kwargs.get("loss_mask_sum")),
               C_chunks=1, A_dim=g, logprob_type=mode)
    for k in _HP_KEYS:
        if kwargs.get(k) is not None:
            cfg[k] = kwargs[k]
    values = kwargs.get("values") if with_critic else None
    lp2 = (logprobs if logprobs.is_cuda else logprobs.to(dev)).reshape(n_units, g)
    v2 = None if values is None else (values if values.is_cuda else values.to(dev)).reshape(n_units, 1)
    return lp2, v2, cfg


@register_policy_loss("actor_critic")
def compute_ppo_actor_critic_loss(**kwargs) -> tuple[torch.Tensor, dict]:
    """PPO actor + critic loss, no value coefficient (losses.py:396-424). Metrics are 0-dim tensors
    (the caller - registry.policy_loss - converts them for the embodied task type)."""
    lp, v, cfg = _from_preprocessed(kwargs, True)
    loss, metrics, _ = _run(lp, v, None, cfg, True, as_float=False)
    return loss, metrics


@register_policy_loss("actor")
def compute_grpo_actor_loss_fn(**kwargs) -> tuple[torch.Tensor, dict]:
    """PPO-clip actor loss for GRPO (losses.py:508-535)."""
    lp, _, cfg = _from_preprocessed(kwargs, False)
    loss, metrics, _ = _run(lp, None, None, cfg, False, as_float=False)
    return loss, metrics


def compute_ppo_actor_loss(**kwargs):
    """losses.py:170-312 - same kernel, actor half only."""
    return compute_grpo_actor_loss_fn(**kwargs)
