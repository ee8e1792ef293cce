"""CPU restatement of one EmbodiedRunner iteration (rollout + advantages + PPO update), single process.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/rl_oracle.py header).  It is the `cpu_baseline` and the
`--impl reference` arm of bench.py: the reference's own runner cannot be launched here (needs Ray, Hydra,
OmegaConf and a simulator - SURVEY.md §8c), so this is the loop of rlinf/runners/embodied_runner.py:478-563
restated with the worker-level steps in the reference's order and with the reference's data movement
kept where it is cheap to keep (per-step `.cpu().contiguous()` staging is a no-op on CPU tensors; list
appends + torch.stack are kept).  The Ray/Channel hops are omitted, which FAVOURS the baseline.

Steps follow: env_worker.py:1059-1349 (_run_interact_once), huggingface_worker.py:678-781
(generate_one_epoch), embodied_fsdp_actor_worker.py:187-321 (recv / advantages), :484-699 (run_training /
train_micro_batch), fsdp_model_manager.py:429-463 (optimizer_step).
"""
from __future__ import annotations

import math
import time

import torch

from . import rl_oracle as O


def bootstrap_rewards(rewards, dones, truncations, bootstrap_values, gamma, bootstrap_type="standard", auto_reset=True):
    """EnvWorker.compute_bootstrap_rewards (rlinf/workers/env/env_worker.py:719-758) for env rewards only (no reward
    model): rewards [B, C] are returned untouched without bootstrap values or without auto_reset; else
    r[:, -1] += gamma * V(final_obs) where the chunk's LAST sub-step is truncated ("standard") or done (any other type).
    Pinned against the reference function body itself (tests/golden/make_golden_r4.py)."""
    adj = rewards.clone()
    if bootstrap_values is None or not auto_reset or dones is None:
        return adj
    flag = (truncations if bootstrap_type == "standard" else dones)[:, -1]
    if not bool(flag.any()):
        return adj
    fv = torch.zeros_like(adj[:, -1], dtype=torch.float32)
    fv[flag] = bootstrap_values[flag].reshape(-1).to(torch.float32)
    adj[:, -1] += gamma * fv
    return adj


def chunk_flags(raw_terminations, raw_truncations):
    """Flag aggregation of env.chunk_step (rlinf/envs/maniskill/maniskill_env.py:355-369): raw per-sub-step flags [B, C]
    are any-reduced over the chunk and reported on its last sub-step only.  Returns (terminations, truncations, past_dones)."""
    past_term = raw_terminations.any(dim=1)
    past_trunc = raw_truncations.any(dim=1)
    chunk_term = torch.zeros_like(raw_terminations)
    chunk_trunc = torch.zeros_like(raw_truncations)
    chunk_term[:, -1] = past_term
    chunk_trunc[:, -1] = past_trunc
    return chunk_term, chunk_trunc, past_term | past_trunc


class SyntheticEnvCPU:
    """Same dynamics as rlinf_b200.envs.SyntheticVectorEnv (same W_s/W_a from the same seed); own RNG."""

    def __init__(self, num_envs, obs_dim, action_dim, max_episode_steps, auto_reset=True, p_term=0.005,
                 noise_std=0.1, reward_noise_std=0.01, seed=1234):
        g = torch.Generator().manual_seed(seed)
        self.w_s = torch.randn(obs_dim, obs_dim, generator=g) / math.sqrt(obs_dim)
        self.w_a = torch.randn(action_dim, obs_dim, generator=g) / math.sqrt(action_dim)
        self.B, self.obs_dim = num_envs, obs_dim
        self.max_episode_steps, self.auto_reset = max_episode_steps, auto_reset
        self.p_term, self.noise_std, self.reward_noise_std = p_term, noise_std, reward_noise_std
        self.gen = torch.Generator().manual_seed(seed + 1)
        self.state = torch.zeros(num_envs, obs_dim)
        self.elapsed = torch.zeros(num_envs, dtype=torch.int32)

    def reset(self):
        self.state = torch.randn(self.B, self.obs_dim, generator=self.gen)
        self.elapsed.zero_()
        return {"states": self.state}, {}

    def step_given_noise(self, state, action, noise):
        """Deterministic step with pre-drawn noise [B, 2*obs+2] (layout of rb200_synth_env_step)."""
        obs = self.obs_dim
        z = state @ self.w_s + action @ self.w_a
        s = torch.tanh(z + self.noise_std * noise[:, :obs])
        reward = -(s * s).sum(-1) / obs + self.reward_noise_std * noise[:, obs]
        self.elapsed += 1
        term = noise[:, obs + 1] < self.p_term
        trunc = (self.elapsed >= self.max_episode_steps) if self.max_episode_steps > 0 else torch.zeros_like(term)
        done = term | trunc
        final = s
        nxt = s
        if self.auto_reset:
            nxt = torch.where(done.unsqueeze(-1), noise[:, obs + 2:], s)
            self.elapsed[done] = 0
        return nxt, final, reward, term, trunc, done

    def substep_given_noise(self, state, action, noise):
        """One sub-step of a chunk WITHOUT auto-reset (maniskill_env.py:339-343: self.step(actions, auto_reset=False));
        noise [B, obs+2] = eps[obs] | eps_r | u_term."""
        obs = self.obs_dim
        z = state @ self.w_s + action @ self.w_a
        s = torch.tanh(z + self.noise_std * noise[:, :obs])
        reward = -(s * s).sum(-1) / obs + self.reward_noise_std * noise[:, obs]
        self.elapsed += 1
        term = noise[:, obs + 1] < self.p_term
        trunc = (self.elapsed >= self.max_episode_steps) if self.max_episode_steps > 0 else torch.zeros_like(term)
        return s, reward, term, trunc

    def chunk_step_multi(self, chunk_actions, noise=None):
        """chunk_step for num_action_chunks = C > 1 (maniskill_env.py:327-375): C sub-steps, raw flags any-reduced over
        the chunk and reported on its last sub-step, one auto-reset after the chunk.  noise [B, C*(obs+2) + obs]."""
        B, C, _ = chunk_actions.shape
        obs = self.obs_dim
        if noise is None:
            parts = []
            for _ in range(C):
                parts += [torch.randn(B, obs + 1, generator=self.gen), torch.rand(B, 1, generator=self.gen)]
            parts.append(torch.randn(B, obs, generator=self.gen))
            noise = torch.cat(parts, dim=1)
        state = self.state
        rewards, terms, truncs = [], [], []
        for c in range(C):
            state, r, t, tr = self.substep_given_noise(state, chunk_actions[:, c], noise[:, c * (obs + 2):(c + 1) * (obs + 2)])
            rewards.append(r)
            terms.append(t)
            truncs.append(tr)
        rewards = torch.stack(rewards, dim=1)
        chunk_term, chunk_trunc, past_done = chunk_flags(torch.stack(terms, dim=1), torch.stack(truncs, dim=1))
        final = state
        nxt = state
        if self.auto_reset:
            nxt = torch.where(past_done.unsqueeze(-1), noise[:, C * (obs + 2):], state)
            self.elapsed[past_done] = 0
        self.state = nxt
        return ([{"states": nxt}], rewards, chunk_term, chunk_trunc, [{"final_observation": {"states": final}}])

    def chunk_step(self, chunk_actions, noise=None):
        B = self.B
        if chunk_actions.shape[1] != 1:
            return self.chunk_step_multi(chunk_actions, noise)
        if noise is None:
            noise = torch.cat([torch.randn(B, self.obs_dim + 1, generator=self.gen),
                               torch.rand(B, 1, generator=self.gen),
                               torch.randn(B, self.obs_dim, generator=self.gen)], dim=1)
        nxt, final, reward, term, trunc, done = self.step_given_noise(self.state, chunk_actions.reshape(B, -1), noise)
        self.state = nxt
        return ([{"states": nxt}], reward.view(B, 1), term.view(B, 1), trunc.view(B, 1),
                [{"final_observation": {"states": final}}])


class RunnerOracle:
    def __init__(self, cfg, params=None):
        self.cfg = cfg
        m, et = cfg["actor"]["model"], cfg["env"]["train"]
        self.B, self.T = et["total_num_envs"], et["max_steps_per_rollout_epoch"]
        self.obs_dim, self.act_dim = m["obs_dim"], m["action_dim"] * m.get("num_action_chunks", 1)
        self.params = params or O.mlp_init(self.obs_dim, m["action_dim"], m.get("num_action_chunks", 1),
                                           seed=cfg["actor"]["seed"])
        for p in self.params.values():
            p.requires_grad_(True)
        o = cfg["actor"]["optim"]
        self.optimizer_steps = 0
        self.critic_warmup_steps = int(o.get("critic_warmup_steps", 0) or 0)  # fsdp_model_manager.py:88-93
        self.opt = self._build_optimizer(self.critic_warmup_steps > 0)
        self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, O.lr_lambda(o, o["lr"]))
        self.num_action_chunks = m.get("num_action_chunks", 1)
        self.n_chunk_steps = self.T // self.num_action_chunks  # env_worker.py: max_steps_per_rollout_epoch // chunks
        self.env = SyntheticEnvCPU(self.B, self.obs_dim, m["action_dim"], et["max_episode_steps"], et["auto_reset"],
                                   et.get("p_term", 0.005), et.get("noise_std", 0.1),
                                   et.get("reward_noise_std", 0.01), et.get("seed", 1234))
        self.gen = torch.Generator().manual_seed(cfg["actor"]["seed"])
        self.obs = None
        self.timers = {}

    def _build_optimizer(self, warmup):
        o = self.cfg["actor"]["optim"]
        opt = O.build_adamw(self.params, o["lr"], o.get("value_lr", o["lr"]),
                            (o.get("adam_beta1", 0.9), o.get("adam_beta2", 0.999)), o.get("adam_eps", 1e-8),
                            o.get("weight_decay", 1e-2), enable_critic_warmup=warmup)
        O.prime_optimizer_state(opt)  # build_optimizer's closing warmup_optimizer_state call
        return opt

    def _optimizer_step(self):
        """FSDPModelManager.optimizer_step (fsdp_model_manager.py:429-463)."""
        self.optimizer_steps += 1
        gn = O.optimizer_step(self.opt, self.params, self.cfg["actor"]["optim"]["clip_grad"])
        if self.critic_warmup_steps > 0:
            lr_list = [0.0 for _ in self.opt.param_groups]
            if self.optimizer_steps >= self.critic_warmup_steps:
                self.opt = self._build_optimizer(False)
                self.critic_warmup_steps = 0
                o = self.cfg["actor"]["optim"]
                self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, O.lr_lambda(o, o["lr"]))
        else:
            lr_list = [g["lr"] for g in self.opt.param_groups]
        return gn, lr_list

    # -- rollout (env_worker.py:1059-1349 / huggingface_worker.py:678-781) -----------------------------
    @torch.no_grad()
    def rollout(self, policy_noise=None, env_noise=None):
        """policy_noise [nc+1,B,C*A] / env_noise [nc,B,2*obs+2] (C = 1) or [nc,B,C*(obs+2)+obs]: pre-drawn draws for the
        parity tests; nc = chunk steps."""
        a = self.cfg["algorithm"]
        gamma, boot_always = a.get("gamma", 1), a.get("bootstrap_type", "standard") != "standard"
        B, T, Cn = self.B, self.n_chunk_steps, self.num_action_chunks
        if self.obs is None or not self.env.auto_reset:  # bootstrap_step, env_worker.py:908-935
            self.obs, _ = self.env.reset()
        lists = {k: [] for k in ("rewards", "dones", "terminations", "truncations", "prev_values", "prev_logprobs",
                                 "states", "action")}
        dones = torch.zeros(B, Cn, dtype=torch.bool)
        term, trunc = dones.clone(), dones.clone()
        rewards, final_obs = None, None
        for t in range(T + 1):
            states = self.obs["states"]
            noise = policy_noise[t] if policy_noise is not None else torch.randn(B, self.act_dim, generator=self.gen)
            action, logp, values = O.mlp_sample(self.params, states, noise)
            boot = None
            if final_obs is not None:  # get_bootstrap_values: second forward on final_obs
                boot = O.mlp_forward(self.params, final_obs["states"], None, want_entropy=False)["values"][:, :1]
            if rewards is not None:  # compute_bootstrap_rewards
                adj = bootstrap_rewards(rewards, dones, trunc, boot, gamma, "always" if boot_always else "standard",
                                        self.env.auto_reset)
                lists["rewards"].append(adj.contiguous())
            lists["dones"].append(dones.contiguous())
            lists["terminations"].append(term.contiguous())
            lists["truncations"].append(trunc.contiguous())
            lists["prev_values"].append(values.contiguous())
            if t == T:
                break
            lists["prev_logprobs"].append(logp.contiguous())
            lists["states"].append(states.contiguous())
            lists["action"].append(action.contiguous())
            obs_list, rewards, term, trunc, infos = self.env.chunk_step(
                action.reshape(B, Cn, -1), None if env_noise is None else env_noise[t])
            self.obs = obs_list[-1]
            dones = term | trunc
            final_obs = infos[-1]["final_observation"]
        batch = {k: torch.stack(v, dim=0) for k, v in lists.items() if k not in ("states", "action")}
        batch["forward_inputs"] = {"states": torch.stack(lists["states"], 0), "action": torch.stack(lists["action"], 0)}
        return batch

    # -- advantages + update ---------------------------------------------------------------------------
    def _prepare(self, batch, rank=0):
        """recv_rollout_trajectories + compute_advantages_and_returns + the seeded shuffle of run_training."""
        cfg, a = self.cfg, self.cfg["algorithm"]
        E = cfg["env"]["train"].get("rollout_epoch", 1)
        if E != 1:
            batch = O.merge_rollout_epochs(batch, E)
        if not cfg["env"]["train"]["auto_reset"] and not cfg["env"]["train"].get("ignore_terminations", False):
            batch["loss_mask"], batch["loss_mask_sum"] = O.loss_mask_from_dones(batch["dones"])
        if a.get("filter_rewards", False):  # embodied_fsdp_actor_worker.py:236-282
            batch["loss_mask"] = O.reward_filter_mask(batch["rewards"], batch.get("loss_mask"), a["group_size"],
                                                      a["rewards_lower_bound"], a["rewards_upper_bound"])
        res = O.adv_and_returns_embodied(a["adv_type"], batch["rewards"], batch["dones"], batch.get("prev_values"),
                                         batch.get("loss_mask"), batch.get("loss_mask_sum"), a.get("gamma", 1),
                                         a.get("gae_lambda", 1), a.get("group_size", 8), a["reward_type"])
        batch.update(res)
        n = batch["prev_logprobs"].shape[0] * batch["prev_logprobs"].shape[1]
        perm = O.shuffle_indices(n, cfg["actor"]["seed"] + rank)
        with torch.no_grad():
            return O.flatten_and_shuffle(batch, perm), n

    def _micro_loss(self, flat, sl, accum, metrics):
        """train_micro_batch (embodied_fsdp_actor_worker.py:591-699): returns the loss to call .backward() on."""
        cfg, a = self.cfg, self.cfg["algorithm"]
        with_critic = a["adv_type"] == "gae"
        A = cfg["actor"]["model"]["action_dim"]
        ent_bonus = a.get("entropy_bonus", 0) or 0
        warm = self.optimizer_steps < self.critic_warmup_steps
        out = O.mlp_forward(self.params, flat["forward_inputs"]["states"][sl], flat["forward_inputs"]["action"][sl],
                            want_entropy=ent_bonus > 0, want_values=with_critic)
        loss, md = O.policy_loss_embodied(
            a["loss_type"], out["logprobs"], flat["prev_logprobs"][sl], flat["advantages"][sl],
            a["logprob_type"], A, loss_mask=None if flat.get("loss_mask") is None else flat["loss_mask"][sl],
            loss_mask_sum=None if flat.get("loss_mask_sum") is None else flat["loss_mask_sum"][sl],
            values=out.get("values") if with_critic else None,
            prev_values=flat["prev_values"][sl] if with_critic else None,
            returns=flat["returns"][sl] if with_critic else None, reward_type=a["reward_type"],
            clip_ratio_low=a["clip_ratio_low"], clip_ratio_high=a["clip_ratio_high"],
            value_clip=a.get("value_clip"), huber_delta=a.get("huber_delta"),
            max_episode_steps=cfg["env"]["train"]["max_episode_steps"] if flat.get("loss_mask_sum") is not None else None,
            critic_warmup=warm)
        md["actor/entropy_loss"] = 0.0
        if ent_bonus > 0 and not warm:
            ent = O.entropy_term(out["entropy"], a["entropy_type"], A, out["logprobs"].shape[0],
                                 None if flat.get("loss_mask") is None else flat["loss_mask"][sl])
            loss = loss - ent_bonus * ent
            md["actor/entropy_loss"] = float(ent.detach())
        loss = loss / accum
        md["actor/total_loss"] = float(loss.detach())
        for kk, vv in md.items():
            metrics.setdefault(kk, []).append(vv)
        return loss

    def _finish_metrics(self, metrics):
        ev = {k: sum(v) for k, v in metrics.items() if k.startswith(O.EV_PREFIX)}
        out = {k: sum(v) / len(v) for k, v in metrics.items() if not k.startswith(O.EV_PREFIX)}
        if ev:
            cnt = ev[O.EV_PREFIX + "count"]
            rc = ev[O.EV_PREFIX + "returns_sq_sum"] - ev[O.EV_PREFIX + "returns_sum"] ** 2 / max(cnt, 1)
            ec = ev[O.EV_PREFIX + "errors_sq_sum"] - ev[O.EV_PREFIX + "errors_sum"] ** 2 / max(cnt, 1)
            out["critic/explained_variance"] = (1 - ec / rc) if (cnt >= 2 and rc != 0) else float("nan")
        return out

    def update(self, batch, rank=0, world_size=1):
        """One rank's view of run_training (gradients of THIS rank only; see update_dp for the data-parallel average)."""
        return self.update_dp([batch], ranks=[rank], world_size=world_size)

    def update_dp(self, batches, ranks=None, world_size=None):
        """Data-parallel run_training emulated in one process: every rank holds the same parameters, shuffles its own
        shard with seed+rank (embodied_fsdp_actor_worker.py:511-513), and the gradients of one optimiser step are the
        MEAN over ranks (FSDP/DDP gradient averaging) - here: sum of the per-rank backward passes / number of ranks."""
        cfg, a = self.cfg, self.cfg["algorithm"]
        ranks = list(range(len(batches))) if ranks is None else ranks
        world_size = world_size or len(batches)
        t0 = time.perf_counter()
        prepared = [self._prepare(b, r) for b, r in zip(batches, ranks)]
        t1 = time.perf_counter()
        n = prepared[0][1]
        per_rank = cfg["actor"]["global_batch_size"] // world_size
        mbs = cfg["actor"]["micro_batch_size"]
        accum = per_rank // mbs
        metrics = {}
        for _ in range(a.get("update_epoch", 1)):
            for gb in range(n // per_rank):
                self.opt.zero_grad()
                for flat, _n in prepared:
                    for k in range(accum):
                        lo = gb * per_rank + k * mbs
                        self._micro_loss(flat, slice(lo, lo + mbs), accum, metrics).backward()
                if len(prepared) > 1:
                    for p in self.params.values():
                        if p.grad is not None:
                            p.grad.div_(len(prepared))
                gn, lr_list = self._optimizer_step()
                metrics.setdefault("actor/grad_norm", []).append(gn)
                metrics.setdefault("actor/lr", []).append(lr_list[0])
                if len(lr_list) > 1:
                    metrics.setdefault("critic/lr", []).append(lr_list[1])
        self.sched.step()
        t2 = time.perf_counter()
        self.timers = {"adv_s": t1 - t0, "train_s": t2 - t1}
        return self._finish_metrics(metrics)

    def run_iteration(self):
        t0 = time.perf_counter()
        batch = self.rollout()
        t1 = time.perf_counter()
        m = self.update(batch)
        self.timers["rollout_s"] = t1 - t0
        return m
