"""The persistent tensor-core rollout kernel (csrc/rollout_tc.cu, rb200_rollout_tc) against the oracle's restatement of
the reference loop (EnvWorker.interact / MultiStepRolloutWorker.generate, env_worker.py:1059-1349,
huggingface_worker.py:678-781) with injected noise - flags bit-exact, floats within 1e-4 - and against the per-kernel
CUDA-graph path on the device Philox streams."""
import pytest
import torch

from oracle.runner_oracle import RunnerOracle

pytestmark = pytest.mark.gpu


def _cpu_batch(b):
    return {k: (_cpu_batch(v) if isinstance(v, dict) else v.detach().cpu().clone()) for k, v in b.items()}


def _compare(b, ob, rtol=1e-4, atol=2e-5):
    for k in ("dones", "terminations", "truncations"):
        assert torch.equal(b[k], ob[k]), k
    for k in ("rewards", "prev_values", "prev_logprobs"):
        torch.testing.assert_close(b[k], ob[k], rtol=rtol, atol=atol, msg=k)
    torch.testing.assert_close(b["forward_inputs"]["states"], ob["forward_inputs"]["states"], rtol=rtol, atol=atol)
    torch.testing.assert_close(b["forward_inputs"]["action"], ob["forward_inputs"]["action"], rtol=rtol, atol=atol)


@pytest.mark.parametrize("B,obs,act,bootstrap_type", [
    (40, 32, 3, "standard"),     # second CTA owns 8 of its 32 environment slots
    (96, 128, 8, "always"),      # config-2 network shapes, bootstrap on every done
    (1000, 128, 8, "standard"),  # 32 CTAs, last one partial
    (64, 64, 1, "standard"),
])
def test_rollout_tc_vs_oracle(B, obs, act, bootstrap_type):
    from rlinf_b200.config import synthetic_ppo_config
    from rlinf_b200.runner import EmbodiedRunner

    T = 20
    cfg = synthetic_ppo_config(B=B, T=T, obs_dim=obs, action_dim=act, **{"rollout.fused_kernel": "tc",
                                                                          "env.train.p_term": 0.03,
                                                                          "env.train.max_episode_steps": 5,
                                                                          "algorithm.bootstrap_type": bootstrap_type})
    run = EmbodiedRunner(cfg)
    assert run.rollout._tc
    orc = RunnerOracle(cfg, params={n: p.detach().cpu().clone() for n, p in run.actor.model.named_parameters()})
    g = torch.Generator().manual_seed(B + obs)
    pn = torch.randn(T + 1, B, act, generator=g)
    en = torch.cat([torch.randn(T, B, obs + 1, generator=g), torch.rand(T, B, 1, generator=g),
                    torch.randn(T, B, obs, generator=g)], -1)
    s0 = torch.randn(B, obs, generator=g)
    orc.env.state = s0.clone()
    orc.obs = {"states": orc.env.state}
    ob = orc.rollout(policy_noise=pn, env_noise=en)
    run.rollout.started = True
    run.buffer.states[0].copy_(s0)
    run.rollout._one_rollout(policy_noise=pn[:T].cuda(), env_noise=en.cuda())
    torch.cuda.synchronize()
    b = _cpu_batch(run.buffer.as_batch())
    assert bool(ob["truncations"].any()) and bool(ob["terminations"].any())
    _compare(b, ob)
    assert (run.env.elapsed.cpu() == orc.env.elapsed).all()


def test_rollout_tc_matches_per_kernel_path_on_device_rng():
    """Same Philox streams and draw order as the per-kernel loop: flags and elapsed counters agree exactly over three
    consecutive rollouts (auto-reset draws included), floats up to the fp32 summation order of the two GEMM paths."""
    from rlinf_b200.config import synthetic_ppo_config
    from rlinf_b200.runner import EmbodiedRunner

    B, T, obs, act = 300, 8, 32, 3
    bufs = []
    for mode in ("tc", False):
        cfg = synthetic_ppo_config(B=B, T=T, obs_dim=obs, action_dim=act, **{"rollout.fused_kernel": mode,
                                                                              "env.train.p_term": 0.03,
                                                                              "env.train.max_episode_steps": 5,
                                                                              "algorithm.bootstrap_type": "always"})
        run = EmbodiedRunner(cfg)
        assert run.rollout._tc == (mode == "tc")
        out = []
        for _ in range(3):
            run.rollout_phase()
            torch.cuda.synchronize()
            out.append(_cpu_batch(run.buffer.as_batch()))
            out[-1]["elapsed"] = run.env.elapsed.cpu().clone()
        bufs.append(out)
    for r in range(3):
        a, b = bufs[0][r], bufs[1][r]
        assert torch.equal(a["elapsed"], b["elapsed"]), r
        _compare(a, b, rtol=2e-3, atol=2e-4)
    assert bool(bufs[0][2]["dones"].any())


def test_rollout_tc_default_and_full_length_episode_statistics():
    """T = 512 at the headline network shapes on the device RNG: finite outputs, the truncation period shows up in the
    flags, and every flagged step carries its bootstrap (rewards differ from the raw reward by gamma * final value)."""
    from rlinf_b200.config import synthetic_ppo_config
    from rlinf_b200.runner import EmbodiedRunner

    B, T = 1024, 512
    cfg = synthetic_ppo_config(B=B, T=T, obs_dim=128, action_dim=8)
    run = EmbodiedRunner(cfg)
    assert run.rollout._tc  # the default from 640 environments per rank on (RolloutWorker.TC_AUTO_MIN_ENVS)
    small = EmbodiedRunner(synthetic_ppo_config(B=256, T=8, obs_dim=128, action_dim=8))
    assert not small.rollout._tc and small.rollout._fused
    run.rollout_phase()
    torch.cuda.synchronize()
    b = _cpu_batch(run.buffer.as_batch())
    for k in ("rewards", "prev_values", "prev_logprobs"):
        assert torch.isfinite(b[k]).all(), k
    assert torch.isfinite(b["forward_inputs"]["states"]).all()
    mes = int(cfg.env.train.max_episode_steps)
    assert bool(b["truncations"][mes].any()) and not bool(b["truncations"][1:mes].any())
    assert float(b["forward_inputs"]["states"].abs().max()) < 8.0
