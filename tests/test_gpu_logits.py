"""rb200_logits_logprob_entropy_fwd / _bwd against the reference's compute_logprobs_from_logits /
compute_entropy_from_logits (rlinf/utils/utils.py:454-512) + autograd (tests/golden/golden_r3.npz, generated from the
unmodified reference) and against the oracle at vocabulary sizes."""
import numpy as np
import pytest
import torch

from oracle import rl_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _t(a):
    a = np.asarray(a)
    return torch.from_numpy(np.ascontiguousarray(a)).reshape(a.shape)


@pytest.mark.parametrize("case", ["lm", "vla", "wide"])
def test_logits_logprob_entropy_golden(golden, case):
    from rlinf_b200 import ops

    g = golden("r3")
    k = lambda n: g[f"lg_{case}_{n}"]  # noqa: E731
    logits = _t(k("logits")).cuda().requires_grad_(True)
    target = _t(k("target")).cuda()
    temp = float(k("temperature"))
    lo, hi = (int(x) for x in k("window"))
    window = None if (lo, hi) == (0, logits.shape[-1]) else (lo, hi)
    lp, ent = ops.logprobs_entropy_from_logits(logits, target, temp, window)
    torch.testing.assert_close(lp.cpu(), _t(k("logprobs")), rtol=RTOL, atol=1e-5)
    # entropies ~1e-5 from logits of magnitude 70 ("wide") carry the reference's own rounding noise (ulp(70) = 7.6e-6)
    torch.testing.assert_close(ent.cpu(), _t(k("entropy")), rtol=RTOL, atol=5e-6)
    g_lp, g_h = _t(k("g_lp")).cuda(), _t(k("g_h")).cuda()
    (lp * g_lp + ent * g_h).sum().backward()
    got = logits.grad.cpu()
    if case == "vla":
        # the reference's autograd yields NaN here (0 * -inf, see oracle docstring): compare with the oracle's closed form
        _, _, want = O.logprobs_entropy_from_logits(_t(k("logits")), _t(k("target")), temp, window, _t(k("g_lp")), _t(k("g_h")))
        assert (got[..., :lo] == 0).all() and (got[..., hi:] == 0).all()
    else:
        want = _t(k("dlogits"))
    torch.testing.assert_close(got, want, rtol=RTOL, atol=2e-6)
    # log-prob only (entropy not requested): the reference gradient is finite in every case
    logits2 = _t(k("logits")).cuda().requires_grad_(True)
    lp2, none = ops.logprobs_entropy_from_logits(logits2, target, temp, window, compute_entropy=False)
    assert none is None
    (lp2 * g_lp).sum().backward()
    torch.testing.assert_close(logits2.grad.cpu(), _t(k("dlogits_lp_only")), rtol=RTOL, atol=2e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_logits_response_slice_in_place_at_vocab_size(dtype):
    """[bsz, S, V] logits, the caller's `[:, -L-1:-1, :]` slice (fsdp_actor_worker.py:488-490) addressed without a copy;
    V = 32003 (not a multiple of the 16-byte vector width), vs the oracle on the same (rounded) inputs."""
    from rlinf_b200 import ops

    bsz, S, Lr, V = 2, 9, 6, 32003
    g = torch.Generator().manual_seed(0)
    full = (torch.randn(bsz, S, V, generator=g) * 2.5).to(dtype)
    target = torch.randint(0, V, (bsz, Lr), generator=g)
    x_dev = full.cuda().requires_grad_(True)
    sl = x_dev[:, -Lr - 1:-1, :]
    assert not sl.is_contiguous()
    lp, ent = ops.logprobs_entropy_from_logits(sl, target.cuda(), temperature=0.9)
    g_lp = torch.randn(bsz, Lr, generator=g)
    g_h = torch.randn(bsz, Lr, generator=g)
    (lp * g_lp.cuda() + ent * g_h.cuda()).sum().backward()
    ref_in = full[:, -Lr - 1:-1, :].float()
    wlp, went, wgrad = O.logprobs_entropy_from_logits(ref_in, target, 0.9, None, g_lp, g_h)
    torch.testing.assert_close(lp.cpu(), wlp, rtol=RTOL, atol=1e-5)
    torch.testing.assert_close(ent.cpu(), went, rtol=RTOL, atol=1e-5)
    got = x_dev.grad.cpu().float()
    assert (got[:, :S - Lr - 1] == 0).all() and (got[:, -1] == 0).all()
    tol = dict(rtol=RTOL, atol=2e-7) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-6)  # bf16 gradient storage
    torch.testing.assert_close(got[:, -Lr - 1:-1, :], wgrad, **tol)


def test_reference_named_wrappers():
    from rlinf_b200 import ops

    g = torch.Generator().manual_seed(1)
    logits = torch.randn(5, 7, 300, generator=g)
    target = torch.randint(0, 300, (5, 7), generator=g)
    wlp, went = O.logprobs_entropy_from_logits(logits, target)
    torch.testing.assert_close(ops.compute_logprobs_from_logits(logits.cuda(), target.cuda()).cpu(), wlp, rtol=RTOL, atol=1e-6)
    torch.testing.assert_close(ops.compute_entropy_from_logits(logits.cuda()).cpu(), went, rtol=RTOL, atol=1e-6)
