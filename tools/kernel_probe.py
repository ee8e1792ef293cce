"""Launch the HBM-bound kernels at BASELINE config-2 shapes for ncu (`-k regex:...`).
Inputs rotate over >300 MB so each profiled launch streams from HBM."""
import sys
import torch
sys.path.insert(0, '.')
from rlinf_b200 import ops, _lib as L

T, B, A = 512, 4096, 8
dev = torch.device('cuda')
lib = L.load()
sets = []
for i in range(9):
    r = torch.randn(T, B, device=dev); v = torch.randn(T + 1, B, device=dev)
    d = (torch.rand(T + 1, B, device=dev) < 0.01).view(torch.uint8)
    sets.append((r, v, d))
for rep in range(2):
    for (r, v, d) in sets:
        ops.gae(r, v, d.view(torch.bool), 0.99, 0.95, None, want_stats=True)
n_all = T * B; mb = n_all // 8
old = torch.randn(n_all, A, device=dev) * .3 - 1
adv, ret, pv = (torch.randn(n_all, 1, device=dev) for _ in range(3))
perm = torch.randperm(n_all, device=dev)
for i in range(10):
    lp = torch.randn(mb, A, device=dev) * .3 - 1; vv = torch.randn(mb, 1, device=dev)
    ix = perm[(i % 8) * mb:][:mb].contiguous()
    ops.ppo_loss(logprobs=lp, values=vv, old_logprobs=old, advantages=adv, returns=ret, prev_values=pv, idx=ix,
                 C_chunks=1, A_dim=A, logprob_type="action_level", value_clip=1.0, huber_delta=10.0)
    # contiguous (pre-gathered) variant
    ops.ppo_loss(logprobs=lp, values=vv, old_logprobs=old[:mb], advantages=adv[:mb], returns=ret[:mb], prev_values=pv[:mb],
                 C_chunks=1, A_dim=A, logprob_type="action_level", value_clip=1.0, huber_delta=10.0)
torch.cuda.synchronize()
print("probe done")
