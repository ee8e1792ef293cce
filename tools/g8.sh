python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 300 python tools/rollout_tc_probe.py 512 4096 > $O/t8_rollout_tc_probe.log 2>&1
timeout 300 python tools/gemm_pf_probe.py > $O/t8_gemm_pf_probe.log 2>&1
