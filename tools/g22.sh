# round-2 final validation: full GPU suite, default bench (with cpu baseline), B=512 bench, launch list, ncu --set full
python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --durations=8 2>&1 | tail -30 > $O/t22_tests.log
timeout 500 python bench.py --steps 5 --warmup 3 > $O/bench_1gpu_final2.json 2> $O/bench_1gpu_final2.err
timeout 200 python bench.py --B 512 --steps 5 --no-cpu-baseline --no-kernel-bench 2>>$O/t22_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B512', d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> $O/t22_bench.log 2>&1
timeout 200 python bench.py --B 1024 --steps 5 --no-cpu-baseline --no-kernel-bench 2>>$O/t22_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B1024', d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> $O/t22_bench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_final2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-bench > $O/launches_final2.log 2>&1
python tools/summarize_launches.py $O/launches_final2.csv > $O/launches_final2_summary.txt 2>&1
gzip -f $O/launches_final2.csv
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'rollout_tc_kernel|gae_tma_kernel|ppo_main_kernel|fwd_block_kernel|bwd_block_kernel|tc_h_gemm_kernel|tc_h_wgrad_kernel|head_fwd_kernel|head_bwd_kernel' -c 40 -f -o $O/ncu_final2 python tools/ncu_targets.py all > $O/ncu_final2.log 2>&1
