python -m rlinf_b200.build > /dev/null 2>&1
mkdir -p gpurun_out/r02
for B in 512 4096; do
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02/launches_b$B.csv python bench.py --B $B --steps 1 --warmup 3 --no-cpu-baseline --no-kernel-bench --rollout fused > gpurun_out/r02/launches_b$B.log 2>&1
python tools/summarize_launches.py gpurun_out/r02/launches_b$B.csv > gpurun_out/r02/launches_b${B}_summary.txt 2>&1
gzip -f gpurun_out/r02/launches_b$B.csv
done
