"""MLP forward+backward time vs the TMA L2-prefetch distance of the fp16-split GEMMs (debug flag bits 8-15)."""
import sys, torch
sys.path.insert(0, '.')
from rlinf_b200 import _lib as L
from rlinf_b200.policy import MLPPolicy
lib = L.load()
for n in (262144, 32768):
    pol = MLPPolicy(obs_dim=128, action_dim=8, seed=0)
    states = torch.randn(n, 128, device='cuda'); action = torch.randn(n, 8, device='cuda')
    dl = torch.randn(n, 8, device='cuda') / n; dv = torch.randn(n, 1, device='cuda') / n
    for pf in (3, 6, 12, 24, 48, 255):
        lib.rb200_debug_set_flags(pf << 8)
        pol.mark_params_changed()
        for _ in range(2):
            pol.forward_train(states, action, compute_entropy=False); pol.backward(dl, dv, None)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps = 4
        tf = tb = 0.0
        for _ in range(reps):
            e[0].record(); pol.forward_train(states, action, compute_entropy=False); e[1].record(); pol.backward(dl, dv, None); e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        print(f"n={n} pf={pf}: fwd {tf/reps:.3f} ms  bwd {tb/reps:.3f} ms", flush=True)
    lib.rb200_debug_set_flags(0)
