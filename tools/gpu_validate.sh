# GPU validation of a build (run under gpurun): full GPU suite, smoke(), default bench, B=512/1024 benches, ncu launch list of
# one bench step and an ncu --set full table of the reported kernels; writes small text outputs under gpurun_out/r02f/
python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02f; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q --timeout 300 --durations=8 2>&1 | tail -30 > $O/tests.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 500 python bench.py --steps 5 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err
timeout 200 python bench.py --B 512 --steps 5 --no-cpu-baseline --no-kernel-bench 2>>$O/bench_small.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B512', d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> $O/bench_small.log 2>&1
timeout 200 python bench.py --B 1024 --steps 5 --no-cpu-baseline --no-kernel-bench 2>>$O/bench_small.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B1024', d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> $O/bench_small.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-bench > $O/launches.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1
gzip -f $O/launches.csv
timeout 400 ncu --set full --clock-control none -k regex:'rollout_tc_kernel|gae_tma_kernel|ppo_main_kernel|fwd_block_kernel|bwd_block_kernel|tc_h_gemm_kernel|tc_h_wgrad_kernel|head_fwd_kernel|head_bwd_kernel' -c 40 -f -o /tmp/ncu_final python tools/ncu_targets.py all > $O/ncu.log 2>&1
python tools/ncu_table.py /tmp/ncu_final.ncu-rep > $O/ncu_table.txt 2>&1
tail -5 $O/ncu.log > $O/ncu_tail.log; rm -f $O/ncu.log
