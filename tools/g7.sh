# tc rollout + logits kernels validation, rollout timing, ncu full capture of the fp16-split GEMMs, default bench
python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_rollout_tc.py -m gpu -q --timeout 300 -x 2>&1 | tail -40 > $O/t7_rollout_tc.log
timeout 300 python -m pytest tests/test_gpu_logits.py -m gpu -q --timeout 200 2>&1 | tail -40 > $O/t7_logits.log
timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -m gpu -q --timeout 200 -s -k "_h_ or timing" 2>&1 | grep -E "tc_|passed|failed|us" | tail -50 > $O/t7_tc_h.log
RB200_PROBE_MODES=tc,graph timeout 300 python tools/rollout_probe.py 512 4096 > $O/t7_rollout_probe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_h_ -c 6 -f -o $O/ncu_tch python tools/mlp_step_probe.py 262144 tc > $O/ncu_tch.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_tc.json 2> $O/bench_1gpu_tc.err
timeout 300 python bench.py --B 512 --steps 5 --no-cpu-baseline --no-kernel-bench 2>>$O/t7_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B512', d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> $O/t7_bench.log 2>&1
