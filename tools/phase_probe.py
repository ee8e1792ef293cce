import sys, time, torch
sys.path.insert(0, '.')
from rlinf_b200.config import synthetic_ppo_config
from rlinf_b200.runner import EmbodiedRunner
def sync(): torch.cuda.synchronize()
for (B,T) in [(512,64),(4096,64),(4096,512)]:
    cfg = synthetic_ppo_config(B=B, T=T, obs_dim=128, action_dim=8, update_epoch=1, num_minibatches=8, **{"rollout.enable_cuda_graph": False})
    run = EmbodiedRunner(cfg)
    for it in range(2):
        t0=time.perf_counter(); run.rollout_phase(); sync(); t1=time.perf_counter()
        run.actor.recv_rollout_trajectories(run.buffer.as_batch()); run.actor.compute_advantages_and_returns(); sync(); t2=time.perf_counter()
        m = run.actor.run_training(); sync(); t3=time.perf_counter()
        print(f"B={B} T={T} it={it} rollout {t1-t0:.3f}s adv {t2-t1:.4f}s train(1 epoch) {t3-t2:.3f}s", flush=True)
    del run
    torch.cuda.empty_cache()
