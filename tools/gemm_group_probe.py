"""MLP forward+backward time with the two towers grouped in one launch vs one launch per tower (debug flags 16 / 32)."""
import sys, torch
sys.path.insert(0, '.')
from rlinf_b200 import _lib as L
from rlinf_b200.policy import MLPPolicy
lib = L.load()
for n in (262144, 131072, 65536, 32768):
    pol = MLPPolicy(obs_dim=128, action_dim=8, seed=0)
    states = torch.randn(n, 128, device='cuda'); action = torch.randn(n, 8, device='cuda')
    dl = torch.randn(n, 8, device='cuda') / n; dv = torch.randn(n, 1, device='cuda') / n
    for name, fl in (("default", 0), ("split", 32), ("xf8", 64)):
        lib.rb200_debug_set_flags(fl)
        pol.mark_params_changed()
        for _ in range(2):
            pol.forward_train(states, action, compute_entropy=False); pol.backward(dl, dv, None)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        reps, tf, tb = 4, 0.0, 0.0
        for _ in range(reps):
            e[0].record(); pol.forward_train(states, action, compute_entropy=False); e[1].record(); pol.backward(dl, dv, None); e[2].record()
            torch.cuda.synchronize()
            tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
        print(f"n={n} {name}: fwd {tf/reps:.3f} ms  bwd {tb/reps:.3f} ms", flush=True)
    lib.rb200_debug_set_flags(0)
    del pol
