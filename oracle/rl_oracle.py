"""CPU oracle for the actor-learner hot path (advantages, PPO/GRPO loss, policy, optimiser).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  Nothing under `rlinf_b200/` imports it: the product
path is the CUDA library and fails loudly when that library is missing.

It is a restatement, in plain fp32 PyTorch-on-CPU arithmetic (the arithmetic the
reference itself uses: the reference is 100 % Python/PyTorch and its embodied
path runs these functions on CPU tensors), of the reference algorithm
(RLinf v0.4.0).  Each function cites the reference file:line it follows.

Parity pinning: the reference's own tests hold NO golden vector for this path
 -> "parity unpinned by the reference's tests".  The oracle
is instead pinned against the reference ITSELF, executed unmodified in the
build container (`tests/golden/make_golden.py` -> `tests/golden/*.npz`,
checked by `tests/test_oracle_golden.py`).

Third-party arithmetic on the path: torch==2.11.0 (pyproject.toml:151 of the
reference; same version here) supplies `torch.optim.AdamW`,
`torch.nn.utils.clip_grad_norm_`, `torch.randperm` (CPU mt19937) and
`torch.distributions.Normal`; the oracle calls the same library routines.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

# --------------------------------------------------------------------------
# masks, reductions
# --------------------------------------------------------------------------


def loss_mask_from_dones(dones: torch.Tensor):
    """rlinf/utils/metric_utils.py:516-537 (compute_loss_mask).

    dones: bool [nc+1, B, C].  Step-major flatten, keep the last nc*C+1 rows,
    a step is valid while no done has been seen in rows [0..t]; drop last row.
    Returns mask bool [nc, B, C] and mask_sum int64 [nc, B, C] (per-env count).
    """
    ncp1, bsz, csz = dones.shape
    nc = ncp1 - 1
    flat = dones.permute(0, 2, 1).reshape(ncp1 * csz, bsz)[-(nc * csz + 1):]
    seen = torch.cumsum(flat.to(torch.int64), dim=0)
    valid = (seen == 0)[:-1]
    mask = valid.reshape(nc, csz, bsz).permute(0, 2, 1)
    per_env = mask.sum(dim=(0, 2), keepdim=True)
    return mask, per_env.expand_as(mask)


def masked_mean(values: torch.Tensor, mask: Optional[torch.Tensor]):
    """rlinf/utils/utils.py:323-330: None -> mean; all-False -> sum(v*m) (=0)."""
    if mask is None:
        return values.mean()
    if bool((~mask).all()):
        return (values * mask).sum()
    return (values * mask).sum() / mask.sum()


def masked_mean_ratio(values, mask, ratio):
    """rlinf/utils/utils.py:352-356: divides by the TOTAL element count."""
    return (values / ratio * mask).mean()


def huber(err: torch.Tensor, delta: float):
    """rlinf/algorithms/utils.py:20-23."""
    a = err.abs()
    return torch.where(a < delta, 0.5 * err**2, delta * (a - 0.5 * delta))


def kl_penalty(logprob, ref_logprob, kind: str):
    """rlinf/algorithms/utils.py:26-64."""
    if kind in ("kl", "k1"):
        return logprob - ref_logprob
    if kind == "abs":
        return (logprob - ref_logprob).abs()
    if kind in ("mse", "k2"):
        return 0.5 * (logprob - ref_logprob).square()
    if kind in ("low_var_kl", "k3"):
        d = torch.clamp(ref_logprob - logprob, min=-20, max=20)
        return torch.clamp(torch.exp(d) - d - 1, min=-10, max=10)
    raise NotImplementedError(kind)


def normalize_valid(x: torch.Tensor, mask: Optional[torch.Tensor]):
    """rlinf/algorithms/utils.py:397-404 (safe_normalize): unbiased std, eps on std."""
    sel = x[mask] if mask is not None else x.reshape(-1)
    if sel.numel() > 0:
        x = (x - sel.mean()) / (sel.std() + 1e-5)
    return x


# --------------------------------------------------------------------------
# advantages
# --------------------------------------------------------------------------


def chunks_to_steps(rewards, dones, values=None, loss_mask=None, loss_mask_sum=None,
                    reward_type="action_level", need_values=True):
    """rlinf/algorithms/utils.py:67-131 (preprocess_embodied_advantages_inputs).

    [nc,B,C] -> [T=nc*C, B]; dones/values [(nc+1),B,C] -> last/first T+1 rows.
    """
    if reward_type == "chunk_level":
        rewards = rewards.sum(dim=-1, keepdim=True)
        dones = dones.max(dim=-1, keepdim=True)[0]
        if loss_mask is not None:
            loss_mask = loss_mask.max(dim=-1, keepdim=True)[0]
        if loss_mask_sum is not None:
            loss_mask_sum = loss_mask_sum.max(dim=-1, keepdim=True)[0]
    nc, bsz, csz = rewards.shape
    n_steps = nc * csz
    r = rewards.transpose(1, 2).reshape(n_steps, bsz)
    m = loss_mask.transpose(1, 2).reshape(n_steps, bsz) if loss_mask is not None else None
    d = dones.transpose(1, 2).reshape((nc + 1) * csz, bsz)[-(n_steps + 1):]
    v = None
    if need_values and values is not None:
        v = values.transpose(1, 2).reshape((nc + 1) * csz, bsz)[: n_steps + 1]
    return dict(rewards=r, dones=d, values=v, loss_mask=m, loss_mask_sum=loss_mask_sum,
                num_chunk=nc, chunk_size=csz, batch_size=bsz, n_steps=n_steps)


def steps_to_chunks(x, num_chunk, chunk_size):
    """rlinf/algorithms/utils.py:155-174."""
    return x.reshape(num_chunk, chunk_size, -1).transpose(1, 2)


def gae(rewards, values=None, dones=None, gamma=1.0, gae_lambda=1.0, loss_mask=None,
        normalize_advantages=True, normalize_returns=False):
    """rlinf/algorithms/advantages.py:24-86.  [T,B] tensors, reverse recurrence.

    Op order is kept exactly (each op rounds to fp32 separately):
      nd = ~done[t+1]; delta = r[t] + gamma*V[t+1]*nd - V[t]
      g = delta + (gamma*lambda)*nd*g ; ret[t] = g + V[t]; adv = ret - V[:-1]
    """
    n_steps = rewards.shape[0]
    returns = torch.zeros_like(rewards)
    critic_free = values is None
    if critic_free:
        gamma, gae_lambda = 1, 1
    g = 0
    for t in range(n_steps - 1, -1, -1):
        nd = ~dones[t + 1]
        delta = rewards[t] if critic_free else rewards[t] + gamma * values[t + 1] * nd - values[t]
        g = delta + gamma * gae_lambda * nd * g
        returns[t] = g if critic_free else g + values[t]
    adv = returns if critic_free else returns - values[:-1]
    if normalize_advantages:
        adv = normalize_valid(adv, loss_mask)
    if normalize_returns:
        returns = normalize_valid(returns, loss_mask)
    return adv, returns


def first_episode_scores(rewards, dones):
    """rlinf/algorithms/utils.py:134-152 (calculate_scores): reverse accumulation
    that is reset by every done -> return of the FIRST episode of each env."""
    n_steps, bsz = rewards.shape
    s = torch.zeros(bsz)
    for t in range(n_steps - 1, -1, -1):
        s = s * ~dones[t + 1]
        s += rewards[t]
    return s


def grpo_group_advantages(scores, loss_mask, group_size):
    """rlinf/algorithms/advantages.py:89-121: unbiased group std, eps 1e-6 on std,
    broadcast over T through the (bool) loss mask."""
    grp = scores.reshape(-1, group_size)
    a = (grp - grp.mean(dim=-1, keepdim=True)) / (grp.std(dim=-1, keepdim=True) + 1e-6)
    return (torch.zeros_like(loss_mask) + a.reshape(1, -1)) * loss_mask


def grpo_video_advantages(rewards, loss_mask, group_size, advantage_mode):
    """rlinf/algorithms/advantages.py:124-164 (compute_grpo_video_advantages): per-frame ("frame") or per-video
    ("video") group normalisation of step-level rewards [num_steps, B], unbiased std, eps 1e-6 on the std."""
    num_steps, batch = rewards.shape
    g = rewards.reshape(num_steps, -1, group_size)
    if advantage_mode == "frame":
        mean, std = g.mean(dim=-1, keepdim=True), g.std(dim=-1, keepdim=True)
    elif advantage_mode == "video":
        mean, std = g.mean(dim=(0, 2), keepdim=True), g.std(dim=(0, 2), keepdim=True)
    else:
        raise ValueError(f"Unsupported grpo_video advantage_mode: {advantage_mode}")
    return ((g - mean) / (std + 1e-6)).reshape(num_steps, batch) * loss_mask


def opd_advantages(prev_logprobs, teacher_logprobs, num_action_chunks, loss_mask=None):
    """rlinf/algorithms/advantages.py:367-407 (compute_opd_advantages): dense reverse-KL reward
    teacher_logp - student_logp per token, reshaped to [..., num_action_chunks, tokens_per_chunk]; a bootstrap row
    beyond the loss mask's time dimension is dropped."""
    adv = teacher_logprobs.float() - prev_logprobs.float()
    assert adv.shape[-1] % num_action_chunks == 0
    adv = adv.reshape(*adv.shape[:-1], num_action_chunks, -1)
    if loss_mask is not None:
        adv = adv[: loss_mask.shape[0]]
    return adv


def opd_actor_loss(logprobs, advantages, loss_mask, loss_mask_sum, max_episode_steps=None):
    """rlinf/algorithms/losses.py:427-505 (compute_opd_actor_loss): -logp * stop_grad(reward), masked mean (or the
    per-episode ratio aggregation); the mask / mask_sum broadcast over the token dimension."""
    if loss_mask.dim() == logprobs.dim() - 1:
        loss_mask = loss_mask.unsqueeze(-1)
    if loss_mask_sum.dim() == logprobs.dim() - 1:
        loss_mask_sum = loss_mask_sum.unsqueeze(-1)
    loss_mask = loss_mask.expand_as(logprobs)
    loss_mask_sum = loss_mask_sum.expand_as(logprobs)
    r = advantages.detach()
    if max_episode_steps is not None:
        loss = masked_mean_ratio(-logprobs * r, loss_mask, (loss_mask_sum * 1.0) / max_episode_steps)
    else:
        loss = masked_mean(-logprobs * r, loss_mask)
    metrics = {"actor/policy_loss": loss.detach(), "actor/opd_reward": masked_mean(r, loss_mask).detach(),
               "actor/opd_reverse_kl": masked_mean(-r, loss_mask).detach()}
    return loss, metrics


def raw_advantages(rewards, loss_mask, normalize_advantages=False):
    """rlinf/algorithms/advantages.py:410-438 (compute_raw_advantages): scores broadcast over the sequence,
    optionally normalised over the valid entries (unbiased std, eps 1e-5).  [SURVEY §8(f) rank 4]"""
    if rewards.ndim == 2:
        rewards = rewards.reshape(-1)
    adv = rewards.unsqueeze(0).expand_as(loss_mask) * loss_mask
    if normalize_advantages:
        valid = adv[loss_mask.bool()]
        if valid.numel() > 0:
            adv = (adv - valid.mean()) / (valid.std() + 1e-5)
    return adv


def reinpp_advantages(rewards, loss_mask, group_size, use_reinpp_baseline=False, kl_beta=0.0, logprob=None,
                      ref_logprob=None, kl_penalty_type=""):
    """rlinf/algorithms/advantages.py:302-364 (compute_reinpp_advantages): reward at the last valid token,
    optional per-token KL penalty, reverse cumulative return over L, masked whitening with the BIASED variance
    clamped at 1e-8 (rsqrt).  loss_mask is [L, B].  [SURVEY §8(f) rank 4]"""
    if use_reinpp_baseline:
        # the reference flattens the baselined rewards to 1-D and then scatters them with a 2-D index
        # (advantages.py:331-345), which torch rejects: the option cannot be used in RLinf v0.4.0 either
        raise IndexError("Dimension out of range (expected to be in range of [-1, 0], but got 1)")
    L = loss_mask.size(0)
    r = torch.zeros_like(loss_mask).float()
    # NOTE the reference flips left-right (dim 1) and then argmax'es over dim 0 (advantages.py:339-341)
    eos = L - 1 - loss_mask.long().fliplr().argmax(dim=0, keepdim=True)
    r = r.scatter_(dim=0, index=eos, src=rewards.view(1, -1).float())
    if kl_beta > 0:
        r = r - kl_beta * kl_penalty(logprob, ref_logprob, kl_penalty_type)
    ret = torch.cumsum(r.flip(dims=[0]), dim=0).flip(dims=[0])
    mean = masked_mean(ret, loss_mask)
    var = masked_mean((ret - mean).pow(2), loss_mask)
    return (ret - mean) * var.clamp(min=1e-8).rsqrt()


def grpo_dynamic_advantages(rewards, loss_mask, group_size, idx_to_traj, advantage_mode="turn"):
    """rlinf/algorithms/advantages.py:167-299 (compute_grpo_dynamic_advantages): GRPO per question over trajectory
    rewards ("trajectory": mean of a trajectory's turn rewards, broadcast back to its turns) or over all turns of the
    question ("turn"); unbiased std, eps 1e-6; broadcast over the sequence through the mask [seq_len, num_sequence]."""
    n = len(idx_to_traj)
    r = rewards.reshape(-1)
    assert r.numel() == n
    n_traj = max(idx_to_traj) + 1
    assert n_traj % group_size == 0
    turn_adv = torch.zeros(n, dtype=r.dtype)
    if advantage_mode == "trajectory":
        tot, cnt = torch.zeros(n_traj, dtype=r.dtype), torch.zeros(n_traj, dtype=torch.long)
        for i, t in enumerate(idx_to_traj):
            tot[t] += r[i]
            cnt[t] += 1
        per_traj = (tot / cnt.clamp(min=1).float()).view(-1, group_size)
        norm = ((per_traj - per_traj.mean(-1, keepdim=True)) / (per_traj.std(-1, keepdim=True) + 1e-6)).view(-1)
        for i, t in enumerate(idx_to_traj):
            turn_adv[i] = norm[t]
    elif advantage_mode == "turn":
        question = torch.tensor([t // group_size for t in idx_to_traj])
        for q in range(n_traj // group_size):
            sel = question == q
            x = r[sel]
            turn_adv[sel] = (x - x.mean()) / (x.std() + 1e-6)
    else:
        raise ValueError(f"Invalid advantage_mode: {advantage_mode}. Must be 'trajectory' or 'turn'")
    return (torch.zeros_like(loss_mask, dtype=r.dtype) + turn_adv.view(1, -1)) * loss_mask


def masked_normalization(x, mask=None, unbiased=False, eps=1e-5, reduce=None):
    """rlinf/utils/distributed.py:866-939 with dim=None: fp64; the input is multiplied by the mask first; `reduce`
    stands for the three SUM all-reduces (a callable applied to the [factor, sum, sumsq] triple)."""
    x = x.to(torch.float64).clone()
    if mask is None:
        factor = torch.tensor(float(x.numel()), dtype=torch.float64)
    else:
        m = mask.to(torch.float64)
        x = x * m
        factor = m.sum()
    s, ss = x.sum(), x.square().sum()
    if reduce is not None:
        factor, s, ss = reduce(factor), reduce(s), reduce(ss)
    mean = s / factor
    var = ss / factor - mean**2
    if unbiased:
        var = var * factor / (factor - 1)
    return ((x - mean) / (var.sqrt() + eps)).float()


def masked_stats(x, mask=None):
    """rlinf/utils/distributed.py:942-954."""
    x = x.to(torch.float64)
    x = x[mask.bool()] if mask is not None else x.reshape(-1)
    return torch.tensor([x.numel(), x.sum(), x.square().sum()], dtype=torch.float64)


def normalize_from_stats(x, stats):
    """rlinf/utils/distributed.py:957-965."""
    stats = stats.to(torch.float64)
    count = stats[0].clamp_min(1.0)
    mean = stats[1] / count
    var = stats[2] / count - mean.square()
    return ((x.to(torch.float64) - mean) * torch.rsqrt(var.clamp_min(0.0) + 1e-5)).float()


def logprobs_entropy_from_logits(logits, target, temperature=1.0, window=None, g_logprobs=None, g_entropy=None):
    """Token log-probabilities and entropies from logits as the reference's callers compute them:
    `logits / temperature` (workers/actor/fsdp_actor_worker.py:478), optional OpenVLA action-bin window - every logit
    outside [lo, hi) set to -inf (models/embodiment/openvla_oft/rlinf/openvla_oft_action_model.py:546-551) - then
    compute_logprobs_from_logits = -cross_entropy (rlinf/utils/utils.py:454-492) and compute_entropy_from_logits =
    -sum(where(p > 0, p * logp, 0)) with logp = log_softmax (:495-512).  fp32 throughout, like the reference on fp32
    logits.  With upstream gradients given, also returns d(sum g_lp*logp + g_h*H)/d logits in closed form:
        inv_T * (g_lp * (onehot - p) - g_h * p * (logp + H))      (zero outside the window).
    NOTE (reference quirk): autograd through the reference's entropy with -inf logits yields NaN for every in-window
    logit (0 * -inf in the backward of p * logp); the closed form here is the finite mathematical gradient."""
    z = logits.float() / temperature
    V = z.shape[-1]
    lo, hi = window if window is not None else (0, V)
    inside = torch.zeros(V, dtype=torch.bool)
    inside[lo:hi] = True
    z = torch.where(inside, z, torch.full_like(z, -math.inf))
    m = z.max(dim=-1, keepdim=True).values
    lse = m + torch.log(torch.exp(z - m).sum(dim=-1, keepdim=True))
    logp = z - lse
    p = torch.exp(logp)
    lp_t = torch.gather(logp, -1, target.unsqueeze(-1)).squeeze(-1)
    ent = -torch.where(p > 0, p * logp, torch.zeros_like(p)).sum(dim=-1)
    if g_logprobs is None and g_entropy is None:
        return lp_t, ent
    onehot = torch.zeros_like(z).scatter_(-1, target.unsqueeze(-1), 1.0)
    grad = torch.zeros_like(z)
    if g_logprobs is not None:
        grad = grad + g_logprobs.unsqueeze(-1) * (onehot - p)
    if g_entropy is not None:
        safe_logp = torch.where(p > 0, logp, torch.zeros_like(logp))
        grad = grad - g_entropy.unsqueeze(-1) * p * (safe_logp + ent.unsqueeze(-1))
    grad = torch.where(inside, grad, torch.zeros_like(grad)) / temperature
    return lp_t, ent, grad


def adv_and_returns_embodied(adv_type, rewards, dones, values=None, loss_mask=None,
                             loss_mask_sum=None, gamma=1.0, gae_lambda=1.0, group_size=8,
                             reward_type="action_level", **kw):
    """rlinf/algorithms/registry.py:95-118 (embodied branch) -> dict."""
    p = chunks_to_steps(rewards, dones, values, loss_mask, loss_mask_sum, reward_type,
                        need_values=(adv_type == "gae"))
    if adv_type == "gae":
        extra = {k: kw[k] for k in ("normalize_advantages", "normalize_returns") if k in kw}
        adv, ret = gae(p["rewards"], p["values"], p["dones"], gamma, gae_lambda,
                       p["loss_mask"], **extra)
    elif adv_type == "grpo":
        sc = first_episode_scores(p["rewards"], p["dones"])
        adv, ret = grpo_group_advantages(sc, p["loss_mask"], group_size), None
    else:
        raise ValueError(adv_type)
    out = {"advantages": steps_to_chunks(adv, p["num_chunk"], p["chunk_size"])}
    if ret is not None:
        out["returns"] = steps_to_chunks(ret, p["num_chunk"], p["chunk_size"])
    return out


def adv_and_returns_reasoning(adv_type, rewards, loss_mask, values=None, gamma=1.0,
                              gae_lambda=1.0, group_size=8, **kw):
    """rlinf/algorithms/registry.py:119-124 + utils.py:177-277 (reasoning branch).

    rewards [bsz]; loss_mask/values [bsz, L] -> (adv [bsz,L], ret [bsz,L] | None).
    """
    bsz, seqlen = loss_mask.shape
    m = loss_mask.transpose(0, 1)
    if adv_type == "gae":
        r = torch.zeros((seqlen, bsz), dtype=rewards.dtype)
        r[-1] = rewards
        v = None
        if values is not None:
            v = torch.cat([values.transpose(0, 1), torch.zeros((1, bsz), dtype=values.dtype)], 0)
        d = torch.zeros(seqlen + 1, bsz, dtype=torch.bool)
        d[-1] = True
        extra = {k: kw[k] for k in ("normalize_advantages", "normalize_returns") if k in kw}
        adv, ret = gae(r, v, d, gamma, gae_lambda, m, **extra)
    elif adv_type == "grpo":
        adv, ret = grpo_group_advantages(rewards.reshape(-1, group_size), m, group_size), None
    else:
        raise ValueError(adv_type)
    adv = adv.transpose(0, 1).contiguous()
    if ret is not None:
        ret = ret.transpose(0, 1).contiguous()
    return adv, ret


def reward_filter_mask(rewards, loss_mask, group_size, lower, upper):
    """rlinf/workers/actor/embodied_fsdp_actor_worker.py:236-282 -> bool [nc,B,1]."""
    if loss_mask is not None:
        rewards = rewards * loss_mask
    nc, bsz, _ = rewards.shape
    per_env = rewards.transpose(0, 1).reshape(bsz, -1)
    grp_mean = per_env.reshape(bsz // group_size, group_size, -1).sum(-1).mean(1)
    keep = ((grp_mean >= lower) & (grp_mean <= upper)).repeat_interleave(group_size)
    keep = keep.unsqueeze(0).expand(nc, -1).unsqueeze(-1)
    return (keep & loss_mask) if loss_mask is not None else keep


def rollout_metrics(buf: dict) -> dict:
    """rlinf/utils/metric_utils.py:422-506 (compute_rollout_metrics) for one rank."""
    mask = buf.get("loss_mask")

    def valid(x):
        if mask is None:
            return x.reshape(-1)
        m = mask.bool()
        if m.ndim == x.ndim - 1:
            m = m.unsqueeze(-1)
        return x[torch.broadcast_to(m, x.shape)]

    out = {}
    if "rewards" in buf:
        v = valid(buf["rewards"])
        out["rewards"] = float(v.float().sum() / v.numel()) if v.numel() else float("nan")
    for k in ("advantages", "returns"):
        if buf.get(k) is not None:
            v = valid(buf[k])
            if v.numel():
                out[f"{k}_mean"], out[f"{k}_max"], out[f"{k}_min"] = float(v.float().sum() / v.numel()), float(v.max()), float(v.min())
            else:
                out[f"{k}_mean"] = out[f"{k}_max"] = out[f"{k}_min"] = float("nan")
    return out


# --------------------------------------------------------------------------
# trajectory indexing
# --------------------------------------------------------------------------

_T_PLUS_ONE_KEYS = ("dones", "terminations", "truncations", "prev_values")


def merge_rollout_epochs(batch, rollout_epoch):
    """rlinf/utils/nested_dict_process.py:251-269: [E*nc,B,..] -> [nc,E*B,..]."""
    out = {}
    for k, v in batch.items():
        if isinstance(v, dict):
            out[k] = merge_rollout_epochs(v, rollout_epoch)
        elif isinstance(v, torch.Tensor):
            x = v.reshape(rollout_epoch, -1, *v.shape[1:]).transpose(0, 1)
            out[k] = x.reshape(x.shape[0], -1, *x.shape[3:])
    return out


def shuffle_indices(n, seed):
    """embodied_fsdp_actor_worker.py:511-513: CPU mt19937 randperm, seed+rank."""
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def flatten_and_shuffle(batch, perm):
    """rlinf/utils/nested_dict_process.py:272-285: drop the bootstrap row of the
    (T+1)-row tensors, flatten [T,B,..] -> [T*B,..] (index t*B+b), gather by perm."""
    out = {}
    for k, v in batch.items():
        if v is None:
            out[k] = None
        elif isinstance(v, dict):
            out[k] = flatten_and_shuffle(v, perm)
        elif isinstance(v, torch.Tensor):
            if k in _T_PLUS_ONE_KEYS:
                v = v[:-1]
            out[k] = v.reshape(-1, *v.shape[2:])[perm]
    return out


# --------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------


def _to_rank(t, shape):
    if t is None:
        return None
    while t.dim() < len(shape) and t.shape != shape:
        t = t.unsqueeze(-1)
    return t


def reduce_loss_inputs(logprobs, old_logprobs, advantages, logprob_type, single_action_dim,
                       loss_mask=None, loss_mask_sum=None, values=None, prev_values=None,
                       returns=None, reward_type="action_level", proximal_logprobs=None, versions=None):
    """rlinf/algorithms/utils.py:280-376 (preprocess_loss_inputs)."""
    if reward_type == "chunk_level":
        flat = lambda t: None if t is None else t.flatten()  # noqa: E731
        advantages, loss_mask, loss_mask_sum = flat(advantages), flat(loss_mask), flat(loss_mask_sum)
        values, prev_values, returns = flat(values), flat(prev_values), flat(returns)
    bsz = logprobs.shape[0]
    if logprob_type == "token_level":
        logprobs = logprobs.reshape(bsz, -1, single_action_dim)
        old_logprobs = old_logprobs.reshape(bsz, -1, single_action_dim)
        advantages = advantages.unsqueeze(-1)
        loss_mask = None if loss_mask is None else loss_mask.unsqueeze(-1)
        loss_mask_sum = None if loss_mask_sum is None else loss_mask_sum.unsqueeze(-1)
        if proximal_logprobs is not None:
            proximal_logprobs = proximal_logprobs.reshape(bsz, -1, single_action_dim)
        if versions is not None:
            versions = versions.reshape(bsz, -1, single_action_dim)
    elif logprob_type == "action_level":
        logprobs = logprobs.reshape(bsz, -1, single_action_dim).sum(-1)
        old_logprobs = old_logprobs.reshape(bsz, -1, single_action_dim).sum(-1)
        if proximal_logprobs is not None:
            proximal_logprobs = proximal_logprobs.reshape(bsz, -1, single_action_dim).sum(-1)
        if versions is not None:
            versions = versions.reshape(bsz, -1, single_action_dim)[..., 0]
    elif logprob_type == "chunk_level":
        logprobs = logprobs.reshape(bsz, -1, single_action_dim).sum(dim=[1, 2])
        old_logprobs = old_logprobs.reshape(bsz, -1, single_action_dim).sum(dim=[1, 2])
        if proximal_logprobs is not None:
            proximal_logprobs = proximal_logprobs.reshape(bsz, -1, single_action_dim).sum(dim=[1, 2])
        if versions is not None:
            versions = versions.reshape(bsz, -1, single_action_dim)[:, 0, 0]
    shp = logprobs.shape
    return dict(logprobs=logprobs, old_logprobs=old_logprobs, proximal_logprobs=proximal_logprobs,
                versions=_to_rank(versions, shp),
                advantages=_to_rank(advantages, shp), loss_mask=_to_rank(loss_mask, shp),
                loss_mask_sum=_to_rank(loss_mask_sum, shp), values=_to_rank(values, shp),
                prev_values=_to_rank(prev_values, shp), returns=_to_rank(returns, shp))


def ppo_actor_loss(logprobs, old_logprobs, advantages, clip_ratio_low, clip_ratio_high,
                   loss_mask=None, clip_ratio_c=None, max_episode_steps=None,
                   loss_mask_sum=None, critic_warmup=False, clip_log_ratio_min=None,
                   clip_log_ratio_max=None, **_):
    """rlinf/algorithms/losses.py:170-312 (compute_ppo_actor_loss)."""
    agg, wratio = masked_mean, None
    if max_episode_steps is not None and loss_mask_sum is not None and loss_mask is not None:
        wratio = (loss_mask_sum * 1.0) / max_episode_steps
        agg = masked_mean_ratio
    if loss_mask is None:
        loss_mask = torch.ones_like(logprobs).bool()
    cnt = float(int(loss_mask.count_nonzero()) or 1)
    lr = logprobs - old_logprobs
    if clip_log_ratio_min is not None:
        lr = torch.clamp(lr, min=clip_log_ratio_min)
    if clip_log_ratio_max is not None:
        lr = torch.clamp(lr, max=clip_log_ratio_max)
    ratio = torch.where(loss_mask, torch.exp(lr), 0)
    kl_terms = torch.where(loss_mask, lr.detach(), 0.0)
    clipped = torch.clamp(ratio, 1.0 - clip_ratio_low, 1.0 + clip_ratio_high)
    l1, l2 = -advantages * ratio, -advantages * clipped
    clip_hit = l1.detach() < l2.detach()
    loss_e = torch.max(l1, l2)
    if clip_ratio_c is not None:
        assert clip_ratio_c > 1.0
        l3 = torch.sign(advantages) * clip_ratio_c * advantages
        dual_hit = l3.detach() < loss_e.detach()
        loss_e = torch.min(loss_e, l3)
    else:
        dual_hit = torch.zeros_like(clip_hit)
    args = (loss_mask,) if agg is masked_mean else (loss_mask, wratio)
    loss_abs = agg(loss_e.abs(), *args)
    loss = agg(loss_e, *args)
    dual_hit = (dual_hit * loss_mask).bool()
    clip_fraction = (clip_hit * loss_mask).sum() / cnt
    approx_kl = -kl_terms.sum() / cnt
    dual_ratio = torch.where(dual_hit, ratio, 0)
    if critic_warmup:
        loss = torch.tensor(0.0)
    mm = loss_mask
    if ratio.dim() > 2 and loss_mask.shape[-1] == 1 and ratio.shape[-1] > 1:
        mm = loss_mask.expand_as(ratio)
    rd = ratio.detach()
    metrics = {
        "actor/policy_loss": loss.detach(),
        "actor/policy_loss_abs": loss_abs.detach(),
        "actor/ratio": masked_mean(rd, mm),
        "actor/ratio_abs": masked_mean((rd - 1).abs(), mm),
        "actor/clipped_ratio": masked_mean(clipped.detach(), mm),
        "actor/dual_cliped_ratio": masked_mean(dual_ratio.detach(), mm),
        "actor/approx_kl": approx_kl.detach(),
        "actor/clip_fraction": clip_fraction.detach(),
    }
    return loss, metrics


def decoupled_ppo_actor_loss(logprobs, old_logprobs, advantages, clip_ratio_low, clip_ratio_high,
                             proximal_logprobs=None, versions=None, current_version=None, loss_mask=None,
                             clip_ratio_c=None, max_episode_steps=None, loss_mask_sum=None,
                             critic_warmup=False, behave_weight_threshold=None, **_):
    """rlinf/algorithms/losses.py:27-167 (compute_decoupled_ppo_actor_loss): PPO clipped around a proximal
    policy (given, or interpolated between behaviour and current policy from the weight versions), importance
    weight exp(prox - old) towards the behaviour policy with an optional cut-off.  [SURVEY §8(f) rank 4]"""
    agg, wratio = masked_mean, None
    if loss_mask is None:
        loss_mask = torch.ones_like(logprobs).bool()
    if max_episode_steps is not None and loss_mask_sum is not None and loss_mask is not None:
        wratio = (loss_mask_sum * 1.0) / max_episode_steps
        agg = masked_mean_ratio
    if proximal_logprobs is None:
        anchor = old_logprobs.detach()
        if versions is not None and current_version is not None:
            # anchor = behaviour policy moved a fraction w towards the current one, w = (age - 1) / age with
            # age = current weight version - version that generated the sample (0 where age <= 0 or version < 0)
            cur = float(current_version)
            born = versions.float()
            age = cur - born
            w = torch.where((age > 0) & (versions >= 0), ((cur - 1.0) - born) / age, torch.zeros_like(born))
            w = w.reshape(w.shape + (1,) * (logprobs.dim() - w.dim())).clamp(0.0, 1.0)
            anchor = (old_logprobs + w * (logprobs - old_logprobs)).detach()
        proximal_logprobs = anchor
    cnt = loss_mask.count_nonzero() or 1  # int64 tensor (fp32 ratios below), as in the reference
    prox_ratio = torch.where(loss_mask, torch.exp(logprobs - proximal_logprobs), 0.0)
    clipped = torch.clamp(prox_ratio, 1.0 - clip_ratio_low, 1.0 + clip_ratio_high)
    l1, l2 = -advantages * prox_ratio, -advantages * clipped
    loss_e = torch.max(l1, l2)
    if clip_ratio_c is not None:
        assert clip_ratio_c > 1.0
        l3 = torch.sign(advantages) * clip_ratio_c * advantages
        dual_hit = l3.detach() < loss_e.detach()
        loss_e = torch.min(loss_e, l3)
    else:
        dual_hit = torch.zeros_like(loss_e, dtype=torch.bool)
    behav_weight = torch.exp(proximal_logprobs - old_logprobs)
    behav_mask = ((behav_weight <= behave_weight_threshold).logical_and(loss_mask)
                  if behave_weight_threshold is not None else loss_mask)
    bcnt = behav_mask.count_nonzero() or 1
    args = (behav_mask,) if agg is masked_mean else (behav_mask, wratio)
    loss = agg(loss_e * behav_weight, *args)
    if critic_warmup:
        loss = torch.tensor(0.0)
    with torch.no_grad():
        metrics = {
            "actor/policy_loss": loss.detach(),
            "actor/proximal_ratio": masked_mean(prox_ratio.detach(), loss_mask),
            "actor/clipped_proximal_ratio": masked_mean(clipped.detach(), loss_mask),
            "actor/clip_fraction": (l1 < l2).logical_and(loss_mask).count_nonzero() / cnt,
            "actor/dual_clip_fraction": dual_hit.logical_and(loss_mask).count_nonzero() / cnt,
            "actor/behav_clip_fraction": 1.0 - (bcnt / cnt),
            "actor/proximal_approx_kl": -torch.where(loss_mask, logprobs - proximal_logprobs, 0.0).sum() / cnt,
            "actor/behav_approx_kl": -torch.where(behav_mask, proximal_logprobs - old_logprobs, 0.0).sum() / bcnt,
        }
        if (versions is not None and current_version is not None and versions.shape == loss_mask.shape
                and bool(loss_mask.any())):
            metrics["actor/average_version"] = versions[loss_mask].float().mean()
            metrics["actor/current_version"] = torch.tensor(float(current_version))
    return loss, metrics


EV_PREFIX = "__sum__/_critic_explained_variance/"


def ppo_critic_loss(values, returns, prev_values, value_clip, huber_delta, loss_mask=None,
                    max_episode_steps=None, loss_mask_sum=None, **_):
    """rlinf/algorithms/losses.py:315-380 + metric_utils.py:232-258."""
    agg, wratio = masked_mean, None
    if max_episode_steps is not None and loss_mask_sum is not None and loss_mask is not None:
        wratio = (loss_mask_sum * 1.0) / max_episode_steps
        agg = masked_mean_ratio
    v_clipped = prev_values + (values - prev_values).clamp(-value_clip, value_clip)
    l_orig = huber(returns - values, huber_delta)
    l_clip = huber(returns - v_clipped, huber_delta)
    le = torch.max(l_orig, l_clip)
    loss = agg(le, loss_mask) if agg is masked_mean else agg(le, loss_mask, wratio)
    clip_ratio = ((v_clipped - prev_values).abs() > value_clip).float().mean()
    r, v = returns.detach().float(), values.detach().float()
    if loss_mask is not None:
        mk = torch.broadcast_to(loss_mask.bool(), r.shape)
        r, v = r[mk], v[mk]
    else:
        r, v = r.reshape(-1), v.reshape(-1)
    e = r - v
    metrics = {
        "critic/value_loss": loss.detach(),
        "critic/value_clip_ratio": clip_ratio.detach(),
        EV_PREFIX + "count": torch.tensor(float(r.numel())),
        EV_PREFIX + "returns_sum": r.sum(),
        EV_PREFIX + "returns_sq_sum": (r * r).sum(),
        EV_PREFIX + "errors_sum": e.sum(),
        EV_PREFIX + "errors_sq_sum": (e * e).sum(),
    }
    return loss, metrics


def policy_loss_embodied(loss_type, logprobs, old_logprobs, advantages, logprob_type,
                         single_action_dim, loss_mask=None, loss_mask_sum=None, values=None,
                         prev_values=None, returns=None, reward_type="action_level", proximal_logprobs=None,
                         versions=None, **hp):
    """rlinf/algorithms/registry.py:77-92 (embodied) + losses.py:383-424,508-535."""
    p = reduce_loss_inputs(logprobs, old_logprobs, advantages, logprob_type, single_action_dim,
                           loss_mask, loss_mask_sum, values, prev_values, returns, reward_type,
                           proximal_logprobs, versions)
    if loss_type == "decoupled_actor_critic":
        loss, metrics = decoupled_ppo_actor_loss(p["logprobs"], p["old_logprobs"], p["advantages"],
                                                 proximal_logprobs=p["proximal_logprobs"], versions=p["versions"],
                                                 loss_mask=p["loss_mask"], loss_mask_sum=p["loss_mask_sum"], **hp)
    else:
        loss, metrics = ppo_actor_loss(p["logprobs"], p["old_logprobs"], p["advantages"],
                                       loss_mask=p["loss_mask"], loss_mask_sum=p["loss_mask_sum"], **hp)
    if loss_type in ("actor_critic", "decoupled_actor_critic"):
        closs, cmetrics = ppo_critic_loss(p["values"], p["returns"], p["prev_values"],
                                          hp["value_clip"], hp["huber_delta"],
                                          loss_mask=p["loss_mask"], loss_mask_sum=p["loss_mask_sum"],
                                          max_episode_steps=hp.get("max_episode_steps"))
        loss = loss + closs
        metrics.update(cmetrics)
    elif loss_type != "actor":
        raise ValueError(loss_type)
    return loss, {k: (float(v) if isinstance(v, torch.Tensor) else v) for k, v in metrics.items()}


def entropy_term(entropy, entropy_type, action_dim, batch_size, loss_mask):
    """rlinf/utils/utils.py:384-408 + embodied_fsdp_actor_worker.py:680-689."""
    if entropy_type == "action_level":
        entropy = entropy.reshape(batch_size, -1, action_dim).sum(-1)
    elif entropy_type == "chunk_level":
        entropy = entropy.sum(-1)
    return masked_mean(entropy, loss_mask)


# --------------------------------------------------------------------------
# MLP policy (models/embodiment/mlp_policy/mlp_policy.py, modules/value_head.py)
# --------------------------------------------------------------------------

HIDDEN = 256
_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def mlp_param_shapes(obs_dim, act_dim, num_action_chunks=1, value_out=None):
    """Parameter names/shapes in `named_parameters()` order of the reference
    MLPPolicy (mlp_policy.py:28-105): own Parameter first, then children in
    construction order (value_head, backbone, actor_mean)."""
    ca = num_action_chunks * act_dim
    vo = num_action_chunks if value_out is None else value_out
    return [
        ("actor_logstd", (1, ca)),
        ("value_head.mlp.0.weight", (HIDDEN, obs_dim)), ("value_head.mlp.0.bias", (HIDDEN,)),
        ("value_head.mlp.2.weight", (HIDDEN, HIDDEN)), ("value_head.mlp.2.bias", (HIDDEN,)),
        ("value_head.mlp.4.weight", (HIDDEN, HIDDEN)), ("value_head.mlp.4.bias", (HIDDEN,)),
        ("value_head.mlp.6.weight", (vo, HIDDEN)),
        ("backbone.0.weight", (HIDDEN, obs_dim)), ("backbone.0.bias", (HIDDEN,)),
        ("backbone.2.weight", (HIDDEN, HIDDEN)), ("backbone.2.bias", (HIDDEN,)),
        ("backbone.4.weight", (HIDDEN, HIDDEN)), ("backbone.4.bias", (HIDDEN,)),
        ("actor_mean.weight", (ca, HIDDEN)), ("actor_mean.bias", (ca,)),
    ]


def mlp_init(obs_dim, act_dim, num_action_chunks=1, seed=0):
    """Random init with the reference's distributions (mlp_policy.py:91-105 ->
    modules/utils.py layer_init: orthogonal(sqrt 2) weights, zero bias;
    value_head.py:50-63: kaiming_normal fan_out / N(0,0.02) last layer)."""
    g = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in mlp_param_shapes(obs_dim, act_dim, num_action_chunks):
        if name == "actor_logstd":
            p[name] = torch.full(shape, -0.5)
        elif name.endswith("bias"):
            p[name] = torch.zeros(shape)
        elif name.startswith("value_head"):
            if name == "value_head.mlp.6.weight":
                p[name] = torch.randn(shape, generator=g) * 0.02
            else:  # kaiming_normal_(mode="fan_out", nonlinearity="tanh"): gain 5/3
                p[name] = torch.randn(shape, generator=g) * ((5.0 / 3.0) / math.sqrt(shape[0]))
        else:
            w = torch.empty(shape)
            torch.nn.init.orthogonal_(w, gain=(0.01 * math.sqrt(2)) if name.startswith("actor_mean") else math.sqrt(2), generator=g)
            p[name] = w
    return p


def mlp_forward(params, states, action=None, want_entropy=True, want_values=True):
    """mlp_policy.py:202-236 (default_forward): tanh MLP 3x256 -> mean;
    state-independent logstd; Normal log_prob / entropy; value MLP 3x256 -> C."""
    F = torch.nn.functional
    h = states
    for i in (0, 2, 4):
        h = torch.tanh(F.linear(h, params[f"backbone.{i}.weight"], params[f"backbone.{i}.bias"]))
    mean = F.linear(h, params["actor_mean.weight"], params["actor_mean.bias"])
    logstd = params["actor_logstd"].expand_as(mean)
    std = torch.exp(logstd)
    out = {"mean": mean, "logstd": logstd}
    if action is not None:
        var = std**2
        out["logprobs"] = -((action - mean) ** 2) / (2 * var) - std.log() - _HALF_LOG_2PI
    if want_entropy:
        out["entropy"] = 0.5 + _HALF_LOG_2PI + torch.log(std)
    if want_values:
        v = states
        for i in (0, 2, 4):
            v = torch.tanh(F.linear(v, params[f"value_head.mlp.{i}.weight"], params[f"value_head.mlp.{i}.bias"]))
        out["values"] = F.linear(v, params["value_head.mlp.6.weight"])
    return out


def mlp_sample(params, states, noise):
    """mlp_policy.py:256-293 (_generate_actions, mode="train") with the N(0,1)
    draw supplied by the caller (parity is on GIVEN noise, not on the RNG):
    action = mean + std * noise; log_prob; value."""
    out = mlp_forward(params, states, None, want_entropy=False, want_values=True)
    std = torch.exp(out["logstd"])
    action = out["mean"] + std * noise
    var = std**2
    logp = -((action - out["mean"]) ** 2) / (2 * var) - std.log() - _HALF_LOG_2PI
    return action, logp, out["values"]


# --------------------------------------------------------------------------
# optimiser (hybrid_engines/fsdp/fsdp_model_manager.py:429-463,501-590; no_shard
# path of strategy/fsdp.py:363-369 = torch.nn.utils.clip_grad_norm_)
# --------------------------------------------------------------------------


def build_adamw(params: dict, lr, value_lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                enable_critic_warmup=False):
    """build_optimizer (fsdp_model_manager.py:501-590). enable_critic_warmup: only the value-head parameters are
    optimised, everything else is frozen (:523-531)."""
    actor = [p for n, p in params.items() if "value_head" not in n]
    critic = [p for n, p in params.items() if "value_head" in n]
    groups = []
    if enable_critic_warmup:
        for p in actor:
            p.requires_grad_(False)
    else:
        for p in params.values():
            p.requires_grad_(True)
        groups.append({"params": actor, "lr": lr, "betas": betas})
    if critic:
        groups.append({"params": critic, "lr": value_lr, "betas": betas})
    return torch.optim.AdamW(groups, eps=eps, weight_decay=weight_decay)


def prime_optimizer_state(optimizer):
    """warmup_optimizer_state (rlinf/utils/utils.py:594-663), the last call of build_optimizer
    (fsdp_model_manager.py:589): one step with every lr set to 0 over the CURRENT .grad tensors (zeros where .grad is
    None), then lr restored and every `step` counter reset to 0.  Parameters do not move (lr = 0 also disables the
    decoupled decay), but the moments do absorb whatever gradient is present: (1-b1) g and (1-b2) g^2.  At construction
    time there are no gradients and this is a no-op; when the optimiser is rebuilt at the end of critic warm-up the
    value-head parameters still carry the clipped gradient of the last warm-up step (golden_r5 "warmup")."""
    saved_lr = [g["lr"] for g in optimizer.param_groups]
    saved_grad = {}
    for g in optimizer.param_groups:
        g["lr"] = 0.0
        for p in g["params"]:
            saved_grad[p] = p.grad
            if p.grad is None:
                p.grad = torch.zeros_like(p)
    with torch.no_grad():
        optimizer.step()
    for g, lr in zip(optimizer.param_groups, saved_lr):
        g["lr"] = lr
    for p, gr in saved_grad.items():
        p.grad = gr
        st = optimizer.state.get(p, {})
        if torch.is_tensor(st.get("step")):
            st["step"].zero_()
        elif "step" in st:
            st["step"] = 0


def lr_lambda(optim_cfg: dict, base_lr: float):
    """build_lr_scheduler + get_lr_scheduler (fsdp_model_manager.py:465-499, fsdp/utils.py:522-604): the LambdaLR
    multiplier as a function of the scheduler step (constant / cosine / openpi_cosine)."""
    g = optim_cfg.get
    total = g("total_training_steps", 0)
    warm = int(g("lr_warmup_steps", -1))
    if warm < 0:
        warm = int(g("lr_warmup_steps_ratio", 0.0) * total)
    kind = g("lr_scheduler", "constant")
    cycles = g("num_cycles", 0.5)
    min_lr, min_lr_rate = g("min_lr", 0.0), g("min_lr_rate", None)

    def constant(s):
        return float(s) / float(max(1.0, warm)) if s < warm else 1.0

    def cosine(s):
        rate = min_lr_rate if min_lr_rate is not None else min_lr / base_lr
        if s < warm:
            return float(s) / float(max(1, warm))
        progress = float(s - warm) / float(max(1, total - warm))
        return max(0, 0.5 * (1.0 + math.cos(math.pi * float(cycles) * 2.0 * progress)) * (1 - rate) + rate)

    def openpi(s):
        mm = min_lr_rate if min_lr_rate is not None else (min_lr / base_lr if (min_lr and base_lr > 0) else 0.0)
        if s < warm:
            init = 1.0 / (warm + 1)
            return init + (1.0 - init) * s / max(1, warm)
        progress = min(1.0, (s - warm) / max(1, total - warm))
        return mm + (1.0 - mm) * 0.5 * (1.0 + math.cos(math.pi * progress))

    return {"constant": constant, "cosine": cosine, "openpi_cosine": openpi, "ref_warmup_cosine": openpi}[kind]


def optimizer_step(optimizer, params: dict, clip_grad):
    gn = torch.nn.utils.clip_grad_norm_([p for p in params.values() if p.grad is not None], clip_grad)
    if torch.isfinite(gn):
        optimizer.step()
    return float(gn)


def adamw_reference_math(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, wd=1e-2):
    """Closed form of one torch.optim.AdamW step (single-tensor path) for
    known-answer checks of the flat-buffer kernel."""
    p = p * (1 - lr * wd)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    bc1, bc2 = 1 - beta1**step, 1 - beta2**step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v
