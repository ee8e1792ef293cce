# round-2 re-entry: full GPU suite + default bench + variants + launch list
python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=15 2>&1 | tail -70 > $O/t5_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_1gpu_default.json 2> $O/bench_1gpu_default.err
for v in "--debug-flags 8" "--rollout fused" ; do
  echo "== B4096 $v" >> $O/t5_bench.log
  timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-kernel-bench $v 2>>$O/t5_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['value'], d['e2e']['value'])" >> $O/t5_bench.log 2>&1
done
for v in "" "--graph-update"; do
  echo "== B512 $v" >> $O/t5_bench.log
  timeout 300 python bench.py --B 512 --steps 5 --no-cpu-baseline --no-kernel-bench $v 2>>$O/t5_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> $O/t5_bench.log 2>&1
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/launches_b4096.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-bench > $O/launches_b4096.log 2>&1
python tools/summarize_launches.py $O/launches_b4096.csv > $O/launches_b4096_summary.txt 2>&1
gzip -f $O/launches_b4096.csv
