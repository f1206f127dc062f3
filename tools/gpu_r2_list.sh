#!/bin/bash
# launch list of the default bench with caches NOT flushed between kernels (closer to the in-step durations)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
ARGS="--steps 2 --warmup 3 --sustained 0 --no-cpu-baseline --no-gpu-reference"
timeout 200 python bench.py $ARGS > gpurun_out/r02g_plain_bench.log 2>&1 &&
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 1300 -c 800 --csv \
   --log-file gpurun_out/r02g_launches_warm.csv python bench.py $ARGS > gpurun_out/r02g_ncu_launches.log 2>&1
echo "launches rc=$?"
