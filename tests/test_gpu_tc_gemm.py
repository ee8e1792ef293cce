"""tcgen05 3xTF32 GEMM building block vs an fp64 matmul (GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K", [(128, 32), (128, 256), (1000, 128), (5000, 256), (262144, 256)])
def test_tc_gemm_matches_fp64(M, K):
    from rlinf_b200 import _lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(256, K, device="cuda", generator=g) / K ** 0.5
    C = torch.empty(M, 256, device="cuda")
    work = torch.empty(2 * M * K + 512 * K, device="cuda")
    L.check(lib.rb200_tc_gemm(L.ptr(A), L.ptr(B), L.ptr(C), M, K, L.ptr(work), L.stream_ptr()), "tc_gemm")
    torch.cuda.synchronize()
    n = min(M, 4096)
    ref = (A[:n].double() @ B.double().T)
    err = (C[:n].double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # fp32-level accuracy (plain TF32 would be ~1e-3 relative)
    assert err <= 5e-6 * scale + 1e-6, (err, scale)
    if M > n:  # spot-check the tail rows too
        ref2 = (A[-256:].double() @ B.double().T)
        assert (C[-256:].double() - ref2).abs().max().item() <= 5e-6 * scale + 1e-6


@pytest.mark.parametrize("n,IN", [(32, 32), (1000, 128), (4096, 256), (70000, 256), (262144, 128)])
def test_tc_wgrad_matches_fp64(n, IN):
    from rlinf_b200 import _lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(n + IN)
    Z = torch.randn(n, 256, device="cuda", generator=g) / n ** 0.5
    H = torch.randn(n, IN, device="cuda", generator=g)
    dW = torch.ones(256, IN, device="cuda")  # accumulates (+=)
    work = torch.empty(2 * n * (256 + IN), device="cuda")
    L.check(lib.rb200_tc_wgrad(L.ptr(Z), L.ptr(H), L.ptr(dW), n, IN, L.ptr(work), L.stream_ptr()), "tc_wgrad")
    torch.cuda.synchronize()
    ref = 1.0 + Z.double().T @ H.double()
    err = (dW.double() - ref).abs().max().item()
    scale = (ref - 1.0).abs().max().item()
    # fp32 accumulation over n samples (TMEM accumulators are fp32, like any fp32 GEMM): error grows ~sqrt(n)
    assert err <= (5e-6 + 2e-7 * n ** 0.5) * scale + 2e-6, (err, scale)


@pytest.mark.skipif(__import__("os").environ.get("RB200_EXPERIMENTAL", "0") != "1",
                    reason="experimental kernel variants (rb200_debug_set_flags): RB200_EXPERIMENTAL=1 to run")
@pytest.mark.parametrize("M,K", [(256, 32), (128, 256), (1000, 128), (5000, 256), (262144, 256)])
def test_experimental_cta_pair_gemm_bit_identical(M, K):
    """cta_group::2 variant (debug flag bit 2, tc_gemm2.cu): same products, same k order -> same bits."""
    from rlinf_b200 import _lib as L

    lib = L.load()
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    B = torch.randn(256, K, device="cuda", generator=g) / K ** 0.5
    work = torch.empty(512 * K, device="cuda")
    outs = []
    try:
        for flags in (0, 4):
            lib.rb200_debug_set_flags(flags)
            C = torch.full((M, 256), float("nan"), device="cuda")
            L.check(lib.rb200_tc_gemm(L.ptr(A), L.ptr(B), L.ptr(C), M, K, L.ptr(work), L.stream_ptr()), "tc_gemm")
            torch.cuda.synchronize()
            outs.append(C)
    finally:
        lib.rb200_debug_set_flags(0)
    assert torch.equal(outs[0], outs[1])
