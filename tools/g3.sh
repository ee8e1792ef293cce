python -m rlinf_b200.build > /dev/null 2>&1
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity2.py -m gpu -q --timeout 600 -k "two_rank or critic_warmup" 2>&1 | tail -40 > gpurun_out/r02/t3_tests.log
for v in "" "--graph-update"; do
echo "== 2gpu $v" >> gpurun_out/r02/t3_bench.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-kernel-bench $v 2>>gpurun_out/r02/t3_bench.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['value'])" >> gpurun_out/r02/t3_bench.log 2>&1
done
