#!/bin/bash
# launch list of the bench command + --set full of the conv kernels (1 GPU); each ncu run follows a plain run
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TAG=${1:-r02}
ARGS="--steps 2 --warmup 3 --sustained 0 --no-cpu-baseline --no-gpu-reference"
timeout 200 python bench.py $ARGS > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 800 --csv \
   --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launches exit=$?"
timeout 100 python tools/kernels_once.py 1 > gpurun_out/${TAG}_plain_once.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2h|wgrad_tc2h|conv_tc2_" -c 10 \
   -o gpurun_out/${TAG}_prof python tools/kernels_once.py 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full exit=$?"
ls -la gpurun_out/${TAG}_*
