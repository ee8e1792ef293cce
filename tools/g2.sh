python -m rlinf_b200.build > /dev/null 2>&1
mkdir -p gpurun_out/r02
RB200_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -60 > gpurun_out/r02/t2_tests.log
