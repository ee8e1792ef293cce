"""num_action_chunks = C > 1 through the device rollout and the update: env.chunk_step semantics
(rlinf/envs/maniskill/maniskill_env.py:327-375 - C sub-steps without auto-reset, flags on the last sub-step, one reset per
chunk), bootstrap on the chunk's last sub-step (env_worker.py:719-758), [nc,B,C] buffer rows, chunk -> step flattening of
the advantages (algorithms/utils.py:67-131), all against the oracle."""
import numpy as np
import pytest
import torch

from oracle.runner_oracle import RunnerOracle

pytestmark = pytest.mark.gpu


def _cpu_batch(b):
    return {k: (_cpu_batch(v) if isinstance(v, dict) else v.detach().cpu().clone()) for k, v in b.items()}


@pytest.mark.parametrize("bootstrap_type", ["standard", "always"])
def test_chunked_rollout_buffer_vs_oracle(bootstrap_type):
    from rlinf_b200.config import synthetic_ppo_config
    from rlinf_b200.runner import EmbodiedRunner

    B, T, obs, A, Cn = 96, 24, 8, 2, 4
    nc = T // Cn
    cfg = synthetic_ppo_config(B=B, T=T, obs_dim=obs, action_dim=A, **{"actor.model.num_action_chunks": Cn,
                                                                        "algorithm.bootstrap_type": bootstrap_type,
                                                                        "env.train.p_term": 0.03,
                                                                        "env.train.max_episode_steps": 10})
    run = EmbodiedRunner(cfg)
    assert not run.rollout._tc and not run.rollout._fused and run.buffer.rewards.shape == (nc, B, Cn)
    orc = RunnerOracle(cfg, params={n: p.detach().cpu().clone() for n, p in run.actor.model.named_parameters()})
    assert torch.equal(run.env.w_a.cpu(), orc.env.w_a) and orc.env.w_a.shape == (A, obs)
    g = torch.Generator().manual_seed(11)
    pn = torch.randn(nc + 1, B, Cn * A, generator=g)
    parts = []
    for _ in range(Cn):
        parts += [torch.randn(nc, B, obs + 1, generator=g), torch.rand(nc, B, 1, generator=g)]
    parts.append(torch.randn(nc, B, obs, generator=g))
    en = torch.cat(parts, -1)
    s0 = torch.randn(B, obs, generator=g)
    orc.env.state = s0.clone()
    orc.obs = {"states": orc.env.state}
    ob = orc.rollout(policy_noise=pn, env_noise=en)
    run.rollout.started = True
    run.buffer.states[0].copy_(s0)
    run.rollout._one_rollout(policy_noise=pn[:nc].cuda(), env_noise=en.cuda())
    torch.cuda.synchronize()
    b = _cpu_batch(run.buffer.as_batch())
    for k in ("dones", "terminations", "truncations"):
        assert b[k].shape == (nc + 1, B, Cn) and torch.equal(b[k], ob[k]), k
        assert not bool(b[k][:, :, :-1].any())  # flags only on the last sub-step of a chunk
    assert bool(ob["truncations"].any()) and bool(ob["terminations"].any())
    assert b["prev_values"].shape == (nc + 1, B, Cn) and b["prev_logprobs"].shape == (nc, B, Cn * A)
    for k in ("rewards", "prev_values", "prev_logprobs"):
        torch.testing.assert_close(b[k], ob[k], rtol=1e-4, atol=2e-5, msg=k)
    torch.testing.assert_close(b["forward_inputs"]["states"], ob["forward_inputs"]["states"], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(b["forward_inputs"]["action"], ob["forward_inputs"]["action"], rtol=1e-4, atol=2e-5)
    assert (run.env.elapsed.cpu() == orc.env.elapsed).all()


def test_chunked_update_matches_oracle_after_k_steps():
    """Rollout on the device RNG, then the same [nc,B,C] batch through RunnerOracle.update and update_phase: parameters
    and metrics after 4 optimiser steps within 1e-4."""
    from rlinf_b200.config import synthetic_ppo_config
    from rlinf_b200.runner import EmbodiedRunner

    B, T, obs, A, Cn = 64, 32, 4, 2, 4
    n = B * (T // Cn)
    cfg = synthetic_ppo_config(B=B, T=T, obs_dim=obs, action_dim=A, update_epoch=2, num_minibatches=2,
                               **{"actor.model.num_action_chunks": Cn, "actor.global_batch_size": n // 2,
                                  "actor.micro_batch_size": n // 2})
    run = EmbodiedRunner(cfg)
    run.rollout_phase()
    torch.cuda.synchronize()
    batch = _cpu_batch(run.buffer.as_batch())
    orc = RunnerOracle(cfg, params={k: p.detach().cpu().clone() for k, p in run.actor.model.named_parameters()})
    om = orc.update(batch)
    m = run.update_phase()
    for name, p in run.actor.model.named_parameters():
        torch.testing.assert_close(p.cpu(), orc.params[name].detach(), rtol=1e-4, atol=2e-5, msg=name)
    for k, v in om.items():
        if k in ("critic/value_clip_ratio",):
            continue
        assert k in m, k
        np.testing.assert_allclose(m[k], v, rtol=2e-4, atol=1e-6, err_msg=k)
    assert run.actor.optimizer.state[0].item() == 4
