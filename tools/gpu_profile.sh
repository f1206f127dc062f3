#!/bin/bash
# Round profile: (1) launch list of the default bench command, (2) ncu --set full of the top kernels.
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 1000 --csv \
   --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "launches exit=$?"
python tools/kernels_once.py > gpurun_out/plain_kernels_once.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc|wgrad_tc|bn_" -s 16 -c 16 \
   -o gpurun_out/prof_r01_final python tools/kernels_once.py > gpurun_out/ncu_final.log 2>&1
echo "full exit=$?"
