# round-2 validation: full GPU suite, default bench, B=512 bench, launch list, ncu --set full of the reported kernels
python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=10 2>&1 | tail -40 > $O/t13_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_1gpu_final.json 2> $O/bench_1gpu_final.err
timeout 300 python bench.py --B 512 --steps 5 --no-cpu-baseline --no-kernel-bench 2>>$O/t13_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B512', d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> $O/t13_bench.log 2>&1
timeout 300 python bench.py --B 1024 --steps 5 --no-cpu-baseline --no-kernel-bench 2>>$O/t13_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('B1024', d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> $O/t13_bench.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_final.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-bench > $O/launches_final.log 2>&1
python tools/summarize_launches.py $O/launches_final.csv > $O/launches_final_summary.txt 2>&1
gzip -f $O/launches_final.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'rollout_tc_kernel|gae_tma_kernel|ppo_main_kernel|fwd_block_kernel|bwd_block_kernel|tc_h_gemm_kernel|tc_h_wgrad_kernel|head_fwd_kernel|head_bwd_kernel' --launch-skip 0 -c 60 -f -o $O/ncu_final python tools/ncu_targets.py all > $O/ncu_final.log 2>&1
