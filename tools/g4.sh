python -m rlinf_b200.build > /dev/null 2>&1
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q --timeout 600 -s -k "_h_ or timing" 2>&1 | tail -80 > gpurun_out/r02/t4_tests.log
