"""Config access with the reference's key names (rlinf/config.py schema is unchanged).

The runner/actor read `cfg.algorithm.*`, `cfg.actor.*`, `cfg.env.train.*`, `cfg.rollout.*`,
`cfg.runner.*` exactly as the reference workers do (examples/embodiment/config/maniskill_ppo_mlp.yaml).
Accepts an OmegaConf DictConfig when omegaconf is installed, or a plain nested dict (wrapped here).
"""
from __future__ import annotations

import copy


class Cfg(dict):
    """dict with attribute access and `.get`, nested."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def wrap(cfg):
    return cfg if not isinstance(cfg, dict) or isinstance(cfg, Cfg) else Cfg(cfg)


def synthetic_ppo_config(B=4096, T=512, obs_dim=128, action_dim=8, update_epoch=8, num_minibatches=8,
                         micro_batch_size=None, world_size=1, adv_type="gae", loss_type="actor_critic",
                         group_size=1, gamma=0.99, gae_lambda=0.95, seed=1234, **over) -> Cfg:
    """The BASELINE.json synthetic workloads expressed in the reference's schema; hyper-parameters follow
    examples/embodiment/config/maniskill_ppo_mlp.yaml (update_epoch 8, 8 mini-batches per epoch,
    clip 0.2/0.2, AdamW lr 3e-4 wd 0.01 clip_grad 0.5, seed 1234)."""
    n = B * T
    gbs = n // num_minibatches
    cfg = Cfg({
        "runner": {"task_type": "embodied", "max_epochs": 1000, "weight_sync_interval": 1},
        "algorithm": {
            "update_epoch": update_epoch, "normalize_advantages": True, "group_size": group_size,
            "reward_type": "action_level", "logprob_type": "action_level", "entropy_type": "action_level",
            "adv_type": adv_type, "loss_type": loss_type, "loss_agg_func": "token-mean", "bootstrap_type": "always",
            "kl_beta": 0.0, "entropy_bonus": 0, "clip_ratio_high": 0.2, "clip_ratio_low": 0.2, "clip_ratio_c": 3.0,
            "value_clip": 1.0, "huber_delta": 10.0, "gamma": gamma, "gae_lambda": gae_lambda,
        },
        "env": {"train": {"rollout_epoch": 1, "total_num_envs": B, "auto_reset": True, "ignore_terminations": False,
                          "max_episode_steps": max(T // 4, 1), "max_steps_per_rollout_epoch": T,
                          "env_type": "synthetic", "p_term": 0.005, "noise_std": 0.1, "reward_noise_std": 0.01,
                          "seed": 1234}},
        "rollout": {"pipeline_stage_num": 1, "enable_cuda_graph": True},
        "actor": {
            "micro_batch_size": micro_batch_size or gbs // world_size, "global_batch_size": gbs, "seed": seed,
            "model": {"model_type": "mlp_policy", "obs_dim": obs_dim, "action_dim": action_dim,
                      "num_action_chunks": 1, "hidden_dim": 256, "precision": "32", "add_value_head": True},
            "optim": {"lr": 3.0e-4, "value_lr": 3.0e-4, "adam_beta1": 0.9, "adam_beta2": 0.999, "adam_eps": 1.0e-8,
                      "weight_decay": 0.01, "clip_grad": 0.5},
        },
    })
    for k, v in over.items():
        node = cfg
        parts = k.split(".")
        for p in parts[:-1]:
            node = node[p]
        node[parts[-1]] = v
    return cfg


def clone(cfg):
    return Cfg(copy.deepcopy(dict(cfg)))
