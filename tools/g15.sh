python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q --timeout 200 -s -k "_h_ or timing" 2>&1 | grep -E "passed|failed|TIMING|Error|error" | tail -8 > $O/t15_tc_h.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_runner.py tests/test_gpu_parity2.py tests/test_gpu_rollout_tc.py -m gpu -q --timeout 300 2>&1 | tail -12 > $O/t15_tests.log
timeout 300 python tools/gemm_group_probe.py > $O/t15_gemm_group_probe.log 2>&1
timeout 300 python tools/rollout_tc_probe.py 512 4096 > $O/t15_rollout_tc_probe.log 2>&1
timeout 300 python bench.py --steps 5 --no-cpu-baseline > $O/bench_t15.json 2> $O/bench_t15.err
