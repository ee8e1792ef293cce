# 2 GPUs (gpurun --gpus 2): 2-rank NCCL parity test against the oracle, 2-rank bench
python -m rlinf_b200.build > /dev/null 2>&1
O=gpurun_out/r02; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity2.py -m gpu -q --timeout 500 -k "two_rank" 2>&1 | tail -15 > $O/t20_two_rank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-kernel-bench > $O/bench_2gpu.json 2> $O/bench_2gpu.err
