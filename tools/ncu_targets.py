"""One launch of every kernel whose roofline the round reports, at the headline shapes (run under ncu --set full):
rollout_tc_kernel (B=4096,T=512), gae_tma_kernel (4096x512), ppo_main_kernel (262144 samples, gathered), logits fwd/bwd
(4096 x 32000 fp32), fp16-split GEMMs through one MLP forward+backward (262144 rows)."""
import sys
import torch
sys.path.insert(0, '.')
from rlinf_b200 import _lib as L, ops
from rlinf_b200.config import synthetic_ppo_config
from rlinf_b200.policy import MLPPolicy
from rlinf_b200.runner import EmbodiedRunner

lib = L.load()
dev = torch.device("cuda")
what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "rollout"):
    run = EmbodiedRunner(synthetic_ppo_config(B=4096, T=512, obs_dim=128, action_dim=8))
    run.rollout_phase(); run.rollout_phase(); torch.cuda.synchronize()
    del run
if what in ("all", "gae"):
    T, B = 512, 4096
    r = torch.randn(T, B, device=dev); v = torch.randn(T + 1, B, device=dev)
    d = (torch.rand(T + 1, B, device=dev) < 0.01).view(torch.uint8)
    for _ in range(2):
        ops.gae(r, v, d, 0.99, 0.95, want_stats=True)
    torch.cuda.synchronize()
if what in ("all", "ppo"):
    n_all, mb, A = 512 * 4096, 262144, 8
    old = torch.randn(n_all, A, device=dev) * 0.3 - 1
    adv, ret, pv = (torch.randn(n_all, 1, device=dev) for _ in range(3))
    idx = torch.randperm(n_all, device=dev)[:mb].contiguous()
    lp, vv = torch.randn(mb, A, device=dev) * 0.3 - 1, torch.randn(mb, 1, device=dev)
    for _ in range(2):
        ops.ppo_loss(logprobs=lp, values=vv, old_logprobs=old, advantages=adv, returns=ret, prev_values=pv, idx=idx,
                     C_chunks=1, A_dim=A, logprob_type="action_level", value_clip=1.0, huber_delta=10.0)
    torch.cuda.synchronize()
if what in ("all", "logits"):
    N, V = 4096, 32000
    x = (torch.randn(N, V, device=dev) * 2).requires_grad_(True)
    tgt = torch.randint(0, V, (N,), device=dev)
    for _ in range(2):
        lp, ent = ops.logprobs_entropy_from_logits(x, tgt, temperature=0.8)
        (lp.sum() + ent.sum()).backward()
        x.grad = None
    torch.cuda.synchronize()
if what in ("all", "mlp"):
    n = 262144
    pol = MLPPolicy(obs_dim=128, action_dim=8, seed=0)
    states = torch.randn(n, 128, device=dev); action = torch.randn(n, 8, device=dev)
    dl = torch.randn(n, 8, device=dev) / n; dv = torch.randn(n, 1, device=dev) / n
    for _ in range(2):
        pol.forward_train(states, action, compute_entropy=False); pol.backward(dl, dv, None)
    torch.cuda.synchronize()
