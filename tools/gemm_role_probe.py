"""Per-role wait / work cycles of the fp16-split forward / dgrad GEMM (CTA 0), at the headline mini-batch shape."""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
from rlinf_b200 import _lib as L
lib = L.load()
NAMES = ["prod_wait_empty", "mma_wait_tmem_empty", "mma_wait_full", "mma_wait_xf", "xf_wait_full", "xf_work",
         "epi_wait_tmem_full", "epi_work", "total"]
prof = torch.zeros(16, dtype=torch.int64, device="cuda")
for M in (262144, 32768):
    K = 256
    A = torch.randn(M, K, device="cuda")
    B = torch.randn(256, K, device="cuda") / 16
    Cm = torch.empty(M, 256, device="cuda")
    work = torch.empty(512 * K, device="cuda")
    for mode in (0, 1):
        for use_prof in (True, False):
            lib.rb200_tc_h_debug(C.c_void_p(prof.data_ptr()) if use_prof else None)
            for _ in range(2):
                L.check(lib.rb200_tc_gemm_h(L.ptr(A), L.ptr(B), L.ptr(Cm), M, K, mode, None, L.ptr(work), L.stream_ptr()), "g")
            torch.cuda.synchronize()
            prof.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            L.check(lib.rb200_tc_gemm_h(L.ptr(A), L.ptr(B), L.ptr(Cm), M, K, mode, None, L.ptr(work), L.stream_ptr()), "g")
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3
            if use_prof:
                c = prof.tolist()
                tiles = (M // 128 + 147) // 148
                print(f"M={M} mode={mode} PROF {us:.1f} us (incl. split kernel); CTA0 {tiles} tiles; cycles: " +
                      " ".join(f"{n}={v}" for n, v in zip(NAMES, c)), flush=True)
            else:
                print(f"M={M} mode={mode} production {us:.1f} us (incl. split kernel)", flush=True)
lib.rb200_tc_h_debug(None)

# ---- wgrad: slots 9..15 = producer-wait-empty, mma-wait-xf, xf-wait-full, xf-wait-op_free, xf-work, total, k-blocks ----
WNAMES = ["prod_wait_empty", "mma_wait_xf", "xf_wait_full", "xf_wait_op_free", "xf_work", "total", "k_blocks"]
for n in (262144, 32768):
    Z = torch.randn(n, 256, device="cuda") / 64
    H = torch.tanh(torch.randn(n, 256, device="cuda"))
    dW = torch.zeros(256, 256, device="cuda")
    for use_prof in (True, False):
        lib.rb200_tc_h_debug(C.c_void_p(prof.data_ptr()) if use_prof else None)
        for _ in range(2):
            L.check(lib.rb200_tc_wgrad_h(L.ptr(Z), L.ptr(H), L.ptr(dW), n, 256, None, L.stream_ptr()), "w")
        torch.cuda.synchronize()
        prof.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(lib.rb200_tc_wgrad_h(L.ptr(Z), L.ptr(H), L.ptr(dW), n, 256, None, L.stream_ptr()), "w")
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3
        if use_prof:
            c = prof.tolist()[9:16]
            print(f"wgrad n={n} PROF {us:.1f} us; CTA0 cycles: " + " ".join(f"{a}={b}" for a, b in zip(WNAMES, c)), flush=True)
        else:
            print(f"wgrad n={n} production {us:.1f} us", flush=True)
lib.rb200_tc_h_debug(None)
