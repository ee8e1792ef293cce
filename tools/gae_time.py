import sys, torch
sys.path.insert(0, '.')
from rlinf_b200 import _lib as L
lib = L.load(); dev = torch.device('cuda')
T = 512
for B, flags in ((4096, 0),):
  if True:
    sets = []
    for i in range(9):
        sets.append((torch.randn(T, B, device=dev), torch.randn(T + 1, B, device=dev), (torch.rand(T + 1, B, device=dev) < 0.01).view(torch.uint8),
                     torch.empty(T, B, device=dev), torch.empty(T, B, device=dev), torch.empty(6, dtype=torch.float64, device=dev)))
    def run():
        for (r, v, d, a, rt, st) in sets:
            L.check(lib.rb200_gae(L.ptr(r), L.ptr(v), L.ptr(d), None, L.ptr(a), L.ptr(rt), L.ptr(st), T, B, 0.99, 0.95, L.stream_ptr()), "gae")
    run(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): run()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(7):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3 / len(sets))
    ts.sort(); t = ts[len(ts) // 2]
    print(f"B={B} gae per launch {t:.2f} us -> {(17*T*B+5*B) / t / 1e3:.0f} GB/s ({(17*T*B+5*B) / t / 1e3 / 6571.2:.3f} of measured HBM)")
