#!/bin/bash
mkdir -p gpurun_out
python tools/bn_once.py > gpurun_out/plain_bn_once.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_" -s 12 -c 6 \
   -o gpurun_out/prof_r1_bn python tools/bn_once.py > gpurun_out/ncu_bn.log 2>&1
echo "ncu exit=$?"
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-260
