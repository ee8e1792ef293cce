"""Time the tcgen05 GEMM / wgrad kernels under the experiment switches (rb200_debug_set_flags) and check that
skipping the in-place hi masking does not change a single bit (tensor core ignores the low 13 mantissa bits)."""
import sys, torch
sys.path.insert(0, '.')
from rlinf_b200 import _lib as L
lib = L.load(); dev = torch.device('cuda')
M, K = 262144, 256
A = torch.randn(M, K, device=dev); B = torch.randn(256, K, device=dev) / 16
C = torch.empty(M, 256, device=dev); work = torch.empty(512 * K, device=dev)
Z = torch.randn(M, 256, device=dev) / 64; H = torch.randn(M, K, device=dev)
dW = torch.zeros(256, K, device=dev)
st = L.stream_ptr()

def t_gemm():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2): L.check(lib.rb200_tc_gemm(L.ptr(A), L.ptr(B), L.ptr(C), M, K, L.ptr(work), st), "g")
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): L.check(lib.rb200_tc_gemm(L.ptr(A), L.ptr(B), L.ptr(C), M, K, L.ptr(work), st), "g")
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 100  # us per call (incl. the tiny weight split)

def t_wgrad():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2): L.check(lib.rb200_tc_wgrad(L.ptr(Z), L.ptr(H), L.ptr(dW), M, K, None, st), "w")
    torch.cuda.synchronize(); e0.record()
    for _ in range(10): L.check(lib.rb200_tc_wgrad(L.ptr(Z), L.ptr(H), L.ptr(dW), M, K, None, st), "w")
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 100

ref_c = ref_w = None
for name, flags in (("default(pf3,nomask)", 0), ("pf off", 255 << 8), ("pf 2", 2 << 8), ("pf 6", 6 << 8), ("pf 12", 12 << 8),
                    ("mask pf3", 1), ("mask pf6", 1 | (6 << 8))):
    lib.rb200_debug_set_flags(flags)
    tg, tw = t_gemm(), t_wgrad()
    dW.zero_(); L.check(lib.rb200_tc_wgrad(L.ptr(Z), L.ptr(H), L.ptr(dW), M, K, None, st), "w"); torch.cuda.synchronize()
    c = C.clone(); w = dW.clone()
    if ref_c is None: ref_c, ref_w = c, w
    print(f"{name:20s} gemm {tg:7.1f} us  wgrad {tw:7.1f} us  C bit-equal {torch.equal(c, ref_c)}  dW maxdiff {(w - ref_w).abs().max().item():.3e}", flush=True)
lib.rb200_debug_set_flags(0)
