python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_rollout_tc.py -m gpu -q --timeout 300 -x 2>&1 | tail -30 > $O/t10_rollout_tc.log
timeout 300 python tools/rollout_tc_probe.py 512 4096 > $O/t10_rollout_tc_probe.log 2>&1
