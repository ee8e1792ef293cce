"""Host-side arithmetic of the optimiser rebuild at the end of critic warm-up (FlatAdamW.reset_state(carry_grads=True))
against torch.optim.AdamW primed the way the reference primes it (oracle.prime_optimizer_state = warmup_optimizer_state,
rlinf/utils/utils.py:594-663; pinned end to end by golden_r5 "warmup").  CPU only: the method is plain tensor code, run
here on a stand-in object that carries the same attributes as the device optimiser."""
import types

import torch

from oracle import rl_oracle as O


def _fake(flat_grads, coef, skipped=0.0, grad_scale=1.0, betas=(0.9, 0.999)):
    n = flat_grads.numel()
    return types.SimpleNamespace(policy=types.SimpleNamespace(flat_grads=flat_grads), betas=betas,
                                 exp_avg=torch.full((n,), 7.0), exp_avg_sq=torch.full((n,), 7.0),
                                 state=torch.tensor([5.0, 1.0, coef, skipped], dtype=torch.float64),
                                 _last_grad_scale=grad_scale)


def test_rebuild_moments_match_primed_torch_adamw():
    from rlinf_b200.policy import FlatAdamW

    g = torch.Generator().manual_seed(0)
    raw = torch.randn(257, generator=g) * 1e-2          # gradient SUM over 2 ranks as it sits in the flat buffer
    world = 2
    avg = raw / world
    norm = avg.norm()
    max_norm = 0.05
    coef = min(1.0, float(max_norm / (norm + 1e-6)))
    fake = _fake(raw.clone(), coef, grad_scale=1.0 / world)
    FlatAdamW.reset_state(fake, carry_grads=True)
    # the reference: the parameter's .grad holds the averaged, clipped gradient when the optimiser is rebuilt
    p = torch.nn.Parameter(torch.randn(257, generator=g))
    p0 = p.detach().clone()
    p.grad = avg * coef
    opt = torch.optim.AdamW([{"params": [p], "lr": 1e-3, "betas": (0.9, 0.999)}], eps=1e-8, weight_decay=1e-2)
    O.prime_optimizer_state(opt)
    st = opt.state[p]
    assert torch.equal(p.detach(), p0) and float(st["step"]) == 0.0 and opt.param_groups[0]["lr"] == 1e-3
    torch.testing.assert_close(fake.exp_avg, st["exp_avg"], rtol=1e-6, atol=1e-12)
    torch.testing.assert_close(fake.exp_avg_sq, st["exp_avg_sq"], rtol=1e-6, atol=1e-16)
    assert float(fake.state[0]) == 0.0 and float(fake.exp_avg.abs().max()) > 0


def test_rebuild_without_carry_and_after_skipped_step_is_zero():
    from rlinf_b200.policy import FlatAdamW

    raw = torch.ones(16)
    fake = _fake(raw, 1.0)
    FlatAdamW.reset_state(fake)
    assert float(fake.exp_avg.abs().max()) == 0.0 and float(fake.exp_avg_sq.abs().max()) == 0.0
    fake = _fake(torch.full((16,), float("inf")), 0.0, skipped=1.0)  # non-finite norm: the step was skipped
    FlatAdamW.reset_state(fake, carry_grads=True)
    assert float(fake.state[0]) == 0.0
    assert float(fake.exp_avg.abs().max()) == 0.0 and float(fake.exp_avg_sq.abs().max()) == 0.0
