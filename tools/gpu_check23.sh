#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1100 -c 800 --csv \
   --log-file gpurun_out/launches_warm.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launches exit=$?"
