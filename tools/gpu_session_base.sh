#!/usr/bin/env bash
# Baseline gpurun session: the whole GPU suite (timed), the default bench, the sampling bench and the ncu launch
# list of one training step (durations + DRAM traffic).  usage: tools/gpu_session_base.sh <tag>
set -u
TAG=${1:-base}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/${TAG}_gpu.txt 2>&1
t0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 --durations=15 > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$? $(( $(date +%s) - t0 ))s"
t0=$(date +%s)
timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$? $(( $(date +%s) - t0 ))s"
timeout 600 python bench.py --mode sample --batch 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-eager > $OUT/${TAG}_sample.json 2> $OUT/${TAG}_sample.err; echo "sample rc=$?"
t0=$(date +%s)
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 7200 -c 2600 --csv \
  --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prof --no-gpu-eager > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$? $(( $(date +%s) - t0 ))s"
tail -25 $OUT/${TAG}_tests.log
