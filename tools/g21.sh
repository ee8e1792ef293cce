python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q --timeout 200 -s -k "_h_ or timing" 2>&1 | grep -E "passed|failed|TIMING|Error|error|assert" | tail -8 > $O/t21_tc_h.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runner.py tests/test_gpu_parity2.py -m gpu -q --timeout 300 -s 2>&1 | grep -E "max \||passed|failed|FAILED|Error" | tail -20 > $O/t21_tests.log
timeout 300 python tools/gemm_group_probe.py > $O/t21_gemm_probe.log 2>&1
timeout 300 python tools/gemm_role_probe.py > $O/t21_gemm_role_probe.log 2>&1
timeout 300 python bench.py --steps 5 --no-cpu-baseline > $O/bench_t21.json 2> $O/bench_t21.err
