"""Where does a step of the persistent tensor-core rollout go?  Ablations (rb200_rollout_tc_debug flags) and CTA 0's
wait-cycle counters.  usage: python tools/rollout_tc_probe.py [B ...]"""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
from rlinf_b200 import _lib as L
from rlinf_b200.config import synthetic_ppo_config
from rlinf_b200.runner import EmbodiedRunner

NAMES = ["prod_wait_empty", "mma_wait_full", "mma_wait_obs", "mma_wait_vhead", "mma_wait_opnd_a", "mma_wait_opnd_v",
         "actor_wait_acc", "value_wait_acc", "env_wait_acc_env", "env_wait_act", "env_wait_vhead", "env_noise_gen",
         "env_finish", "actor_epilogue", "actor_heads_sample", "total"]
lib = L.load()
Bs = [int(x) for x in sys.argv[1:]] or [512, 4096]
T = 128
for B in Bs:
    cfg = synthetic_ppo_config(B=B, T=T, obs_dim=128, action_dim=8, **{"rollout.fused_kernel": "tc"})
    run = EmbodiedRunner(cfg)
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    for flags in (0, 1, 8, 1 | 8, 4, 2, 2 | 4, 1 | 2 | 4 | 8):
        lib.rb200_rollout_tc_debug(flags, C.c_void_p(prof.data_ptr()))
        run.rollout_phase(); torch.cuda.synchronize()
        prof.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run.rollout_phase(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        c = prof.tolist()
        tot = max(c[15], 1)
        print(f"B={B} T={T} flags={flags:2d}: {ms:.2f} ms = {ms*1e3/T:.1f} us/step (instrumented build); cycles/step: " +
              " ".join(f"{n}={v/T:.0f}" for n, v in zip(NAMES, c)), flush=True)
    lib.rb200_rollout_tc_debug(0, None)
    run.rollout_phase(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run.rollout_phase(); e1.record(); torch.cuda.synchronize()
    print(f"B={B} T={T} production kernel: {e0.elapsed_time(e1):.2f} ms", flush=True)
    del run
