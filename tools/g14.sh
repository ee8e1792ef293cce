python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 300 python tools/gemm_group_probe.py > $O/t14_gemm_group_probe.log 2>&1
timeout 400 python -m pytest tests/test_gpu_chunked.py -m gpu -q --timeout 300 2>&1 | tail -30 > $O/t14_chunked.log
