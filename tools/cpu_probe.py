import os, sys, time, torch
sys.path.insert(0, '.')
from rlinf_b200.config import synthetic_ppo_config
from oracle.runner_oracle import RunnerOracle
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), flush=True)
try:
    print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip(), flush=True)
except Exception as e:
    print("cgroup:", e, flush=True)
for thr in (8, 32, 64, 128):
    torch.set_num_threads(thr)
    x = torch.randn(8192, 256); w = torch.randn(256, 256)
    t = time.perf_counter()
    for _ in range(50): y = torch.tanh(x @ w)
    dt = (time.perf_counter() - t) / 50
    print(f"threads={thr}: [8192x256]x[256x256]+tanh {dt*1e3:.2f} ms -> {2*8192*256*256/dt/1e9:.0f} GFLOP/s", flush=True)
    x = torch.randn(262144, 256)
    t = time.perf_counter()
    for _ in range(5): y = torch.tanh(x @ w)
    dt = (time.perf_counter() - t) / 5
    print(f"threads={thr}: [262144x256]x[256x256]+tanh {dt*1e3:.2f} ms -> {2*262144*256*256/dt/1e9:.0f} GFLOP/s", flush=True)
for thr, B in ((32, 256), (128, 256)):
    torch.set_num_threads(thr)
    cfg = synthetic_ppo_config(B=B, T=64, obs_dim=128, action_dim=8, update_epoch=2, num_minibatches=4)
    r = RunnerOracle(cfg)
    t = time.perf_counter(); r.run_iteration(); dt = time.perf_counter() - t
    print(f"threads={thr} B={B} T=64 2 epochs: {dt:.2f}s {r.timers}", flush=True)
