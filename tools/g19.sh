python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 300 python tools/gemm_role_probe.py > $O/t19_gemm_role_probe.log 2>&1
