import sys, torch
sys.path.insert(0, '.')
from rlinf_b200 import ops
from oracle import rl_oracle as O
for (T, B) in ((128, 160), (64, 32), (77, 160), (77, 4096), (512, 4096)):
    torch.manual_seed(T + B)
    r = torch.randn(T, B); v = torch.randn(T + 1, B); d = torch.rand(T + 1, B) < 0.02
    try:
        adv, ret, st = ops.gae(r, v, d, 0.99, 0.95, None, want_stats=True)
        torch.cuda.synchronize()
        oa, orr = O.gae(r, v, d, 0.99, 0.95, normalize_advantages=False)
        print(T, B, "ret", torch.equal(ret.cpu(), orr), "adv", torch.equal(adv.cpu(), oa), flush=True)
    except Exception as e:
        print(T, B, "ERR", str(e)[:300], flush=True)
        break
