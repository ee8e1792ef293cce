mkdir -p gpurun_out/r02
RB200_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02/t0_tests.log
for v in "" "--graph-update" "--debug-flags 2" "--graph-update --debug-flags 2"; do
  echo "== B512 $v" >> gpurun_out/r02/t0_bench.log
  timeout 300 python bench.py --B 512 --steps 5 --no-cpu-baseline --no-kernel-bench $v 2>>gpurun_out/r02/t0_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['wall_ms_per_step'])" >> gpurun_out/r02/t0_bench.log 2>&1
done
for v in "" "--debug-flags 4"; do
  echo "== B4096 $v" >> gpurun_out/r02/t0_bench.log
  timeout 300 python bench.py --steps 5 --no-cpu-baseline $v 2>>gpurun_out/r02/t0_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['roofline'].get('us_per_launch'), {k:(v['us_per_launch'],v['frac']) for k,v in d.get('roofline_hbm_kernels',{}).items()})" >> gpurun_out/r02/t0_bench.log 2>&1
done
